# N=2 slab checks of profiles/slab_check_n2_r02.log: ragged 1000x1100 and 8192x4096 against a single-GPU solver, bit for bit.
cd /root/repo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29512 tools/slab_check.py 1000 1100 2 40 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5
timeout 600 $T --master-port 29513 tools/slab_check.py 8192 4096 2 30 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5
