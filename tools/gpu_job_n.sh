# bash tools/gpu_job_n.sh N : slab check at 4096^2 and the default bench line on N GPUs (profiles/bench_c5_nN_r02.json, slab_check_nN_r02.log).
cd /root/repo
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $T --master-port 29533 tools/slab_check.py 4096 4096 2 30 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -$((2*N+1)) | tee gpurun_out/r2_slab_n$N.log
timeout 900 $T --master-port 29534 bench.py --gpus $N --steps 2 --warmup 3 2> gpurun_out/r2_b_n$N.err > gpurun_out/r2_b_n$N.json; head -c 250 gpurun_out/r2_b_n$N.json; echo; tail -2 gpurun_out/r2_b_n$N.err
