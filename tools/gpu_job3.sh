cd /root/repo
mkdir -p gpurun_out
for lib in - tools/_variants/lib_xchg.so tools/_variants/lib_cl4.so tools/_variants/lib_xchgcl4.so; do
timeout 300 python tools/tune_pcg.py $lib 16384 12 2>&1 | tail -1
timeout 300 python tools/tune_pcg.py $lib 4096 40 2>&1 | tail -1
timeout 300 python tools/tune_pcg.py $lib 1024 100 2>&1 | tail -1
done
timeout 300 python tools/trace_fused.py tools/_variants/lib_xchg.so 16384 2>&1 | grep "==" 
cp tools/_variants/lib_xchg.so incremental-fluids-java_b200/libincfluid_b200.so
timeout 600 python -m pytest tests/test_parity_at_size_gpu.py -q -x -k "f3 or f4" 2>&1 | tail -2
