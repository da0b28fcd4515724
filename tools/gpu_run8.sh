cd /root/repo
timeout 200 python tools/tune_pcg.py - 4096 12 > gpurun_out/plain8.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_precon_sweep -s 12 -c 2 -o gpurun_out/prof_sweep1_r02 python tools/tune_pcg.py - 4096 12 > gpurun_out/ncu8.log 2>&1
tail -3 gpurun_out/ncu8.log; ls -la gpurun_out/*.ncu-rep
