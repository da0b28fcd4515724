"""Where does the wavefront of one applyPreconditioner stall?  Prints, per sweep, the strip-to-strip start lags that
exceed twice the median, and a summary (diagnostic; single GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incfluid_b200 as ifl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=2)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
for rep in range(3):
    tr = g.wavefront_trace()
for sw, name in ((0, "forward"), (1, "backward")):
    t = tr[sw]
    order = np.argsort(t[:, 6])
    t = t[order]
    t0 = t[:, 0].min()
    start = (t[:, 1] - t0) / 1e3
    end = (t[:, 3] - t0) / 1e3
    enter = (t[:, 0] - t0) / 1e3
    lag = np.diff(start)
    dur = end - start
    med = np.median(lag)
    print(f"== {name}: total {end.max():.1f} us; median lag {med:.2f} us; sum of lags {lag.sum():.1f}; strip duration first {dur[0]:.1f} "
          f"median {np.median(dur):.1f} last {dur[-1]:.1f} us")
    for i in np.nonzero(lag > 2 * med)[0]:
        print(f"   slot {i+1:4d}: lag {lag[i]:8.2f} us  (entered {enter[i+1]:8.2f}, predecessor started block 1 at {start[i]:8.2f}, sm {t[i+1,7]})")
    late = np.nonzero(enter[1:] > start[:-1])[0]
    print(f"   strips that entered after their predecessor had started: {len(late)}")
