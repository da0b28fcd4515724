cd /root/repo
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -q -m gpu -x -k "gauss or f1 or F1 or golden or F2" 2>&1 | tail -2
timeout 300 python bench.py --workload c1 --steps 40 --warmup 5 --no-cpu-baseline --no-hash --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c1 value %.3e ms/step %.2f kernel_ms %s its %s' % (d['value'], d['ms_per_step'], d['kernel_ms'], d['config']['solver_iterations_per_step'][:5]))"
