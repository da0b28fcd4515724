cd /root/repo
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29512 tools/slab_check.py 1000 1100 2 40 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -6 | tee gpurun_out/r2_slab_n2.log
timeout 600 $T --master-port 29513 tools/slab_check.py 8192 4096 2 30 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -6 | tee -a gpurun_out/r2_slab_n2.log
timeout 900 $T --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 2> gpurun_out/r2_b_n2.err > gpurun_out/r2_b_n2.json; head -c 300 gpurun_out/r2_b_n2.json; echo; tail -3 gpurun_out/r2_b_n2.err
