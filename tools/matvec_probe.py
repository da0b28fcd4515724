"""Mat-vec alone (matrixVectorProduct F3:544 without the fused dot product): wall time per launch over a batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incfluid_b200 as ifl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=2)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
for mode in ("position-derived matrix", "matrix arrays"):
    g.run_stage("build_matrix", 0.005)
    if mode == "matrix arrays":
        os.environ["X"] = "1"
        g.write("adiag", g.read("adiag"))
    for rep in range(3):
        g.run_stage("matvec"); g.synchronize()
    t0 = time.perf_counter()
    for rep in range(40):
        g.run_stage("matvec")
    g.synchronize()
    dt = (time.perf_counter() - t0) / 40
    print(f"{mode}: {dt*1e3:.3f} ms per launch without the dot epilogue")
