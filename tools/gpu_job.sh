cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2_b_n1_s2.json 2>gpurun_out/r2_b_n1_s2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b_n1_s2.json').read()); print(d['value'], d['ms_per_step'], d['config']['state_hash'], d['e2e']['value'], d['roofline']['traffic'])"
