"""Slab-decomposed (multi-GPU) check, run under torchrun: every rank steps the slab solver; rank 0 also steps a
single-GPU solver on the same scene and compares bit for bit (the reductions are summed in global strip order, so the
result must not depend on the number of GPUs).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py 256 3"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import incfluid_b200 as ifl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 600
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [ifl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, device=local, rank=rank, world_size=world,
                    comm_id=ids[0], solver_limit=limit)
ref = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, device=local, solver_limit=limit) if rank == 0 else None
ok = True
for step in range(steps):
    g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    t0 = time.perf_counter()
    st = g.update(0.005)
    dt = time.perf_counter() - t0
    if ref is not None:
        ref.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        t0 = time.perf_counter()
        sr = ref.update(0.005)
        dtr = time.perf_counter() - t0
        same = all(np.array_equal(g.read(k), ref.read(k)) for k in ("p", "d", "u", "v"))
        ok = ok and same and st["iterations_pressure"] == sr["iterations_pressure"]
        print(f"step {step}: slab x{world} {dt*1e3:.1f} ms its {st['iterations_pressure']} | single {dtr*1e3:.1f} ms its "
              f"{sr['iterations_pressure']} | bit-identical {same}", flush=True)
flag = torch.tensor([1 if ok else 0])
dist.broadcast(flag, src=0)
dist.barrier()
g.close()
if rank == 0:
    print("SLAB CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
