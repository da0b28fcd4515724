"""Small driver for ncu: a few PCG iterations at n x n (default 1024) so that -k regex:k_precon_sweep -s N picks a steady-state sweep."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incfluid_b200 as ifl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=8)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
print(g.update(0.005))
