cd /root/repo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/trace_fused.py - 4096 2>&1 | grep "==" | tee gpurun_out/r2_trace14.log
for sz in 16384 4096 1024; do timeout 200 python tools/tune_pcg.py - $sz 40 2>&1 | tail -1; done | tee gpurun_out/r2_tune14.log
timeout 600 python -m pytest tests/test_parity_at_size_gpu.py -q -m gpu -x 2>&1 | tail -3
