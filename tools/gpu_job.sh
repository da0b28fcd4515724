# N=1 measurement recipe of profiles/r02_summary.md section A (run on a GPU box from the repo root): default bench line,
# ncu launch list of one timed step, ncu --set full capture of the three PCG kernels (each only after the plain command exited 0).
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real
head -c 300 gpurun_out/r2_bench_n1.json; echo; tail -3 gpurun_out/r2_bench_n1.err
NCU="ncu --clock-control none"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --no-hash"
timeout 300 $CMD > gpurun_out/plain_c5.log 2>&1 && timeout 900 $NCU --metrics gpu__time_duration.sum -s 5600 -c 300 --csv --log-file gpurun_out/launches_c5_r02.csv $CMD > gpurun_out/ncu_c5.log 2>&1
timeout 300 python tools/prof_target.py c5 1 12 > gpurun_out/plain_c5b.log 2>&1 && timeout 900 $NCU --set full --import-source on -k regex:"k_step_a|k_precon_sweep" -s 12 -c 3 -o gpurun_out/prof_pcg_c5_r02 python tools/prof_target.py c5 1 12 > gpurun_out/ncu_c5b.log 2>&1
ls -la gpurun_out/prof_pcg_c5_r02.ncu-rep gpurun_out/launches_c5_r02.csv
