"""A single strip (h = 32): the pure in-strip cost of the sweep kernel, for ncu source-level sampling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incfluid_b200 as ifl
w = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = ifl.FluidSolver(w, 32, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=4)
g.addInflow(0.0, 0.0, 100.0, 1.0, 1.0, 0.0, 3.0)
st = g.update(0.005)
g.set_profiling(True)
g.addInflow(0.0, 0.0, 100.0, 1.0, 1.0, 0.0, 3.0)
st = g.update(0.005)
print(st, {k: g.kernel_timing(k) for k in ("precon_fwd", "precon_bwd")})
