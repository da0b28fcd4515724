"""Per-strip timeline of one applyPreconditioner on a slab-decomposed solver (run under torchrun; diagnostic).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/trace_slab.py 16384"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import incfluid_b200 as ifl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [ifl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, device=local, rank=rank, world_size=world,
                    comm_id=ids[0], solver_limit=3)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
for rep in range(3):
    dist.barrier()
    tr = g.wavefront_trace()
out = []
for sw, name in ((0, "forward"), (1, "backward")):
    t = tr[sw]
    mine = np.nonzero(t[:, 3])[0]
    t0 = t[mine, 0].min()
    out.append(f"== rank {rank} {name}: strips {mine.min()}..{mine.max()}  kernel-enter(abs ns) {t0}  last end {(t[mine,3].max()-t0)/1e3:.1f} us")
    order = mine[np.argsort(t[mine, 6])]
    if len(order) > 24: order = np.concatenate([order[:8], order[len(order)//2-2:len(order)//2+2], order[-8:]])
    for k in order:
        e = t[k]
        out.append(f"  strip {k:4d} slot {e[6]:4d} sm {e[7]:3d} enter {(e[0]-t0)/1e3:8.2f} after-block-1 {(e[1]-t0)/1e3:8.2f} "
                   f"end {(e[3]-t0)/1e3:8.2f} dur {(e[3]-e[1])/1e3:8.2f} us")
gathered = [None] * world
dist.all_gather_object(gathered, "\n".join(out))
if rank == 0:
    print("\n".join(gathered), flush=True)
dist.barrier()
g.close()
dist.destroy_process_group()
