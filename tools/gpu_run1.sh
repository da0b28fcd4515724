set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
tail -5 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log; tail -15 gpurun_out/r2_t1.log
timeout 1200 python -m pytest tests/test_parity_at_size_gpu.py -x -q -m gpu -s > gpurun_out/r2_t2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2.log; tail -25 gpurun_out/r2_t2.log
timeout 900 python -m pytest tests/test_parity_solids_gpu.py tests/test_parity_flip_gpu.py -x -q -m gpu > gpurun_out/r2_t3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3.log; tail -15 gpurun_out/r2_t3.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e --no-hash > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "bench rc=$?"; cat gpurun_out/r2_b1.json; tail -5 gpurun_out/r2_b1.err
