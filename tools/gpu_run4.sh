cd /root/repo
mkdir -p gpurun_out
for v in - tools/_variants/lib_nostg.so tools/_variants/lib_nomax.so tools/_variants/lib_noproc.so; do timeout 200 python tools/tune_pcg.py $v 16384 40 2>&1 | tail -1; done | tee gpurun_out/r2_tune4.log
