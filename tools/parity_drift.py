"""Per-step fast-path drift between the CUDA path and the oracle (diagnostic for the tolerance tests)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import incfluid_b200 as ifl
from oracle_harness import OracleSolver
import scenes
from test_parity_solids_gpu import pair

abi = ifl.abi
variant = getattr(abi, sys.argv[1] if len(sys.argv) > 1 else "F6_HEAT")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spec, rect, kw = scenes.solid_scene(abi, variant)
g, o = pair(ifl, OracleSolver, variant, 128, 128, spec, **kw)
for step in range(steps):
    for s in (g, o):
        s.add_inflow(*rect)
    a, b = g.update(scenes.DT), o.update(scenes.DT)
    row = [f"step {step:2d} its {a['iterations_heat']}/{a['iterations_pressure']} vs {b['iterations_heat']}/{b['iterations_pressure']}"]
    for n in ["d", "t", "u", "v", "p"] if variant >= abi.F6_HEAT else ["d", "u", "v", "p"]:
        x, y = g.read(n), o.read(n)
        row.append(f"{n}: abs {np.max(np.abs(x-y)):.2e} max {np.max(np.abs(y)):.2e}")
    print(" | ".join(row))
