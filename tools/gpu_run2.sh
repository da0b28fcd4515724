set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in - tools/_variants/lib_lag3.so tools/_variants/lib_lag1.so; do timeout 200 python tools/tune_pcg.py $v 16384 64 2>&1 | tail -1; done | tee gpurun_out/r2_tune2.log
timeout 200 python tools/tune_pcg.py - 4096 64 2>&1 | tail -1 | tee -a gpurun_out/r2_tune2.log
timeout 200 python tools/tune_pcg.py - 1024 64 2>&1 | tail -1 | tee -a gpurun_out/r2_tune2.log
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/r2_t_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_all.log; grep -v "^$" gpurun_out/r2_t_all.log | tail -40
