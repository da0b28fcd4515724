"""Slab-decomposed (multi-GPU) check, run under torchrun: every rank steps its slab of the decomposed solver AND a
single-GPU solver of the whole grid on the same device, and compares the rows it owns bit for bit after every step
(p, d, u, v and the iteration count).  The reductions are summed in global strip order and every rank performs the
reference's per-cell operations, so the result must not depend on the number of GPUs.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py W H steps [limit]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import incfluid_b200 as ifl

w = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = int(sys.argv[2]) if len(sys.argv) > 2 else w
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
limit = int(sys.argv[4]) if len(sys.argv) > 4 else 600
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [ifl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
F3 = ifl.abi.F3_CONJUGATE_GRADIENTS
g = ifl.FluidSolver(w, h, 0.1, variant=F3, device=local, rank=rank, world_size=world, comm_id=ids[0], solver_limit=limit)
ref = ifl.FluidSolver(w, h, 0.1, variant=F3, device=local, solver_limit=limit)
ok = True
inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
for step in range(steps):
    g.addInflow(*inflow)
    t0 = time.perf_counter(); st = g.update(0.005); dt = time.perf_counter() - t0
    ref.addInflow(*inflow)
    t0 = time.perf_counter(); sr = ref.update(0.005); dtr = time.perf_counter() - t0
    same = st["iterations_pressure"] == sr["iterations_pressure"] and st["residual_pressure"] == sr["residual_pressure"]
    detail = []
    for name in ("p", "d", "u", "v"):
        r0, nr = g.slab_rows(name)
        mine = g.read_field_rows(name, r0, nr)
        full = ref.read_field(name)
        width = full.size // (h + (1 if name == "v" else 0))
        eq = np.array_equal(mine, full[r0 * width:(r0 + nr) * width])
        detail.append(f"{name}[{r0}:{r0 + nr}]={'ok' if eq else 'DIFF'}")
        same = same and eq
    ok = ok and same
    print(f"rank {rank}/{world} step {step}: slab {dt*1e3:.1f} ms its {st['iterations_pressure']} | single {dtr*1e3:.1f} ms its "
          f"{sr['iterations_pressure']} | {' '.join(detail)} | bit-identical {same}", flush=True)
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
g.close(); ref.close()
if rank == 0:
    print(f"SLAB CHECK {w}x{h} x{world}", "PASSED" if int(flag.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
