cd /root/repo
for v in - tools/_variants/lib_nodst.so tools/_variants/lib_noedgest.so tools/_variants/lib_prededge.so tools/_variants/lib_noload.so; do echo "## $v"; timeout 300 python tools/trace_fused.py $v 4096 2>&1 | grep "==" ; done | tee gpurun_out/r2_trace7.log
