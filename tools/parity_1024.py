"""configs[1] size: one F3 timestep at 1024^2, CUDA fast path against the oracle (iteration counts and relative field
differences; the serial-order reference needs 564 iterations, its residual creeps under the 1e-5 threshold slowly)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import incfluid_b200 as ifl
import scenes
from oracle_harness import OracleSolver
abi = ifl.abi
n = 1024
g = ifl.FluidSolver(n, n, scenes.DENSITY, variant=abi.F3_CONJUGATE_GRADIENTS)
o = OracleSolver(abi, abi.default_config(abi.F3_CONJUGATE_GRADIENTS, n, n))
for s in (g, o):
    s.add_inflow(*scenes.F3_INFLOW)
sg, so = g.update(scenes.DT), o.update(scenes.DT)
print("iterations gpu", sg["iterations_pressure"], "oracle", so["iterations_pressure"], "residual gpu", sg["residual_pressure"], "oracle", so["residual_pressure"])
for k in ("p", "d", "u", "v"):
    a, b = g.read(k), o.read(k)
    print(k, "max|diff| / max|ref| =", float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)))
