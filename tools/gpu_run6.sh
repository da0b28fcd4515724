cd /root/repo
for sz in 4096 16384; do timeout 300 python tools/trace_fused.py - $sz 2>&1 | grep "==" ; done | tee gpurun_out/r2_trace6.log
