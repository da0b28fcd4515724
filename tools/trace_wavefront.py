"""Print the per-strip timeline of one applyPreconditioner (diagnostic, see ifl_wavefront_trace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incfluid_b200 as ifl
if os.environ.get("IFL_VARIANT_LIB"):          # kernel-variant experiments (tools/ only)
    ifl.package._lib.LIB_PATH = os.environ["IFL_VARIANT_LIB"]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hh = int(os.environ.get("IFL_TRACE_H", n))
g = ifl.FluidSolver(n, hh, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
if len(sys.argv) > 2 and sys.argv[2] == 'random':
    g.write_field('r', np.random.default_rng(0).normal(size=n*hh))
for rep in range(2):
    tr = g.wavefront_trace()
for sw, name in ((0, "forward"), (1, "backward")):
    t = tr[sw]
    t0 = t[:, 0].min()
    print(f"== {name}: total {(t[:,3].max()-t0)/1e3:.1f} us")
    order = np.argsort(t[:, 6])
    if len(order) > 40: order = np.concatenate([order[:12], order[-6:]])
    for k in order:
        e = t[k]
        print(f"strip {k:4d} slot {e[6]:4d} sm {e[7]:3d} enter {(e[0]-t0)/1e3:8.2f} after-block-1 {(e[1]-t0)/1e3:8.2f} "
              f"end {(e[3]-t0)/1e3:8.2f} dur {(e[3]-e[1])/1e3:8.2f} us  [IFL_SW_TRACE_WAITS builds] starved steps {e[2]//1000000} re-reads {e[2]%1000000} blocks1-3 {e[4]/1e3:.2f} us credit+edge clocks {e[5]}")
