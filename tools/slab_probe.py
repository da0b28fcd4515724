"""Per-kernel timing of the slab-decomposed F3 PCG iteration (capped iteration count), run under torchrun:
  python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 tools/slab_probe.py 16384 24"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import incfluid_b200 as ifl

n = int(sys.argv[1]); limit = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [ifl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, device=local, rank=rank, world_size=world,
                    comm_id=ids[0], solver_limit=limit)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
g.set_profiling(True)
dist.barrier()
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
t0 = time.perf_counter()
st = g.update(0.005)
dt = time.perf_counter() - t0
ph = g.last_timing()
line = f"rank {rank}: its {st['iterations_pressure']} step {dt*1e3:.1f} ms projection {ph['projection_ms']:.1f} ms -> {ph['projection_ms']/max(st['iterations_pressure'],1):.3f} ms/iteration |"
for k in ("matvec_dot", "update_pr", "precon_fwd", "precon_bwd", "update_s"):
    ms, ns = g.kernel_timing(k)
    line += f" {k} {ms:.3f}"
out = [None] * world
dist.all_gather_object(out, line)
if rank == 0:
    print(f"grid {n}^2 world {world} sweep kernel {'two-row' if os.environ.get('IFL_SWEEP2') else 'one-row'}")
    print("\n".join(out), flush=True)
dist.barrier()
g.close()
dist.destroy_process_group()
