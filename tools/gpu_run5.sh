cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -15
for sz in 16384 4096 1024; do timeout 200 python tools/tune_pcg.py - $sz 40 2>&1 | tail -1; done | tee gpurun_out/r2_tune5.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t5_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5_all.log; grep -v "^$" gpurun_out/r2_t5_all.log | tail -30
