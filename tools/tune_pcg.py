"""Per-kernel timing of the PCG iteration for a (variant) build of the library:
   python tools/tune_pcg.py [path/to/lib.so] [grid] [limit]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incfluid_b200 as ifl
from importlib import import_module
lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else None
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 96
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if lib:
    import_module(ifl.FluidSolver.__module__.rsplit(".", 1)[0] + "._lib").LIB_PATH = os.path.abspath(lib)
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=limit, flags=flags)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0); g.update(0.005)
g.set_profiling(True)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0); st = g.update(0.005)
kt = {k: g.kernel_timing(k) for k in ("step_a", "precon_fwd", "precon_bwd")}
tot = sum(v[0] for v in kt.values())
print(f"{lib or 'default'} flags {flags} grid {n} its {st['iterations_pressure']}: " + "  ".join(f"{k} {v[0]:.3f} ms" for k, v in kt.items()) + f"  | iteration {tot:.3f} ms", flush=True)
