cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_at_size_gpu.py -q -x -k "f3 or f4" 2>&1 | tail -3
timeout 300 python tools/tune_pcg.py - 16384 12 2>&1 | tail -4
timeout 300 python tools/trace_fused.py - 16384 2>&1 | grep -v "^   strip warp\|loader warp" | tee gpurun_out/r2_trace_mover.log | cut -c1-330
timeout 300 python tools/tune_pcg.py - 4096 40 2>&1 | tail -2
timeout 300 python tools/tune_pcg.py - 1024 100 2>&1 | tail -2
