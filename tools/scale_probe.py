"""Per-kernel timing of the F3 PCG iteration at large grids (capped iteration count)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incfluid_b200 as ifl
if os.environ.get("IFL_VARIANT_LIB"):          # kernel-variant experiments (tools/ only)
    ifl.package._lib.LIB_PATH = os.environ["IFL_VARIANT_LIB"]

n = int(sys.argv[1]); limit = int(sys.argv[2]) if len(sys.argv) > 2 else 24
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=limit)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
g.set_profiling(True)
t0 = time.perf_counter()
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
st = g.update(0.005)
dt = time.perf_counter() - t0
cells = n * n
print(f"grid {n}^2 its {st['iterations_pressure']} step {dt*1e3:.1f} ms  phases {g.last_timing()}")
for k, bytes_per_cell in (("matvec_dot", 16), ("update_pr", 24), ("precon_fwd", 40), ("precon_bwd", 48), ("update_s", 40)):
    ms, ns = g.kernel_timing(k)
    if ns:
        print(f"  {k:12s} {ms:9.3f} ms  {bytes_per_cell*cells/ms/1e6:8.1f} GB/s  ({ns} samples)")
