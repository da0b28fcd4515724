"""Slab-decomposed (multi-GPU) check, run under torchrun: every rank steps its slab of the decomposed solver AND a
single-GPU solver of the whole grid on the same device, and compares the rows it owns bit for bit after every step
(p, d, u, v and the iteration count).  The reductions are summed in global strip order and every rank performs the
reference's per-cell operations, so the result must not depend on the number of GPUs.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py W H steps [limit] [variant 3..7]
Variants 4-7 (solid bodies; every rank holds the whole state, only the solver loops are slab-decomposed) run the scene of
their reference main() and also compare t and the heat solve's iteration count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import incfluid_b200 as ifl

w = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = int(sys.argv[2]) if len(sys.argv) > 2 else w
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
limit = int(sys.argv[4]) if len(sys.argv) > 4 else 600
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [ifl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
abi = ifl.abi
fields = ("p", "d", "u", "v")
if variant == abi.F3_CONJUGATE_GRADIENTS:
    pos, kw = (), {}
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)
else:
    import scenes
    spec, inflow, over = scenes.solid_scene(abi, variant)
    bodies = scenes.make_bodies(ifl, spec)
    if variant >= abi.F6_HEAT:
        pos, kw = (over["density_soot"], 0.01, bodies), {}
        fields = ("p", "d", "t", "u", "v")
    else:
        pos, kw = (bodies,), {}
g = ifl.FluidSolver(w, h, 0.1, *pos, variant=variant, device=local, rank=rank, world_size=world, comm_id=ids[0], solver_limit=limit, **kw)
ref = ifl.FluidSolver(w, h, 0.1, *pos, variant=variant, device=local, solver_limit=limit, **kw)
ok = True
for step in range(steps):
    g.add_inflow(*inflow)
    t0 = time.perf_counter(); st = g.update(0.005); dt = time.perf_counter() - t0
    ref.add_inflow(*inflow)
    t0 = time.perf_counter(); sr = ref.update(0.005); dtr = time.perf_counter() - t0
    same = all(st[k] == sr[k] for k in ("iterations_pressure", "residual_pressure", "iterations_heat", "residual_heat"))
    detail = []
    for name in fields:
        r0, nr = g.slab_rows(name)
        mine = g.read_field_rows(name, r0, nr)
        full = ref.read_field(name)
        width = full.size // (h + (1 if name == "v" else 0))
        eq = np.array_equal(mine, full[r0 * width:(r0 + nr) * width])
        detail.append(f"{name}[{r0}:{r0 + nr}]={'ok' if eq else 'DIFF'}")
        same = same and eq
    ok = ok and same
    print(f"rank {rank}/{world} step {step}: slab {dt*1e3:.1f} ms its {st['iterations_heat']}+{st['iterations_pressure']} | single {dtr*1e3:.1f} ms its "
          f"{sr['iterations_heat']}+{sr['iterations_pressure']} | {' '.join(detail)} | bit-identical {same}", flush=True)
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
g.close(); ref.close()
if rank == 0:
    print(f"SLAB CHECK F{variant} {w}x{h} x{world}", "PASSED" if int(flag.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
