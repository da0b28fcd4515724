cd /root/repo
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for v in 4 5 6 7; do
timeout 200 $T --master-port 2952$v tools/slab_check.py 1024 1024 2 40 $v 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5 | cut -c1-260 | tee -a gpurun_out/r2_slab_masked_n2.log
done
timeout 400 $T --master-port 29534 bench.py --workload c3 --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_b_c3_n2.err > gpurun_out/r2_b_c3_n2.json; head -c 300 gpurun_out/r2_b_c3_n2.json; echo; tail -3 gpurun_out/r2_b_c3_n2.err
