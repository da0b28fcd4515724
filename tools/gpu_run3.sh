cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/trace_fused.py tools/_variants/lib_tw.so 16384 2>&1 | tail -8 | tee gpurun_out/r2_trace3.log
for w in c1 c2 c3 c4; do timeout 400 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-hash 2>gpurun_out/r2_b_$w.err | tee gpurun_out/r2_b_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'value %.3e ms/step %.1f its %s kernel_ms %s phase %s launches %d' % (d['value'], d['ms_per_step'], d['config']['solver_iterations_per_step'], d['kernel_ms'], d['phase_ms_last_step'], d['gpu_launches']))"; tail -2 gpurun_out/r2_b_$w.err; done
