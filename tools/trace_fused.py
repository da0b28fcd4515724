"""Where does the time of the two sweep kernels of a PCG iteration go?  Needs a -DIFL_SW_TRACE_WAITS build:
   tools/build_variant.sh tw -DIFL_SW_TRACE_WAITS ; python tools/trace_fused.py tools/_variants/lib_tw.so [grid]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import incfluid_b200 as ifl
from importlib import import_module
if len(sys.argv) > 1 and sys.argv[1] != "-":
    import_module(ifl.FluidSolver.__module__.rsplit(".", 1)[0] + "._lib").LIB_PATH = os.path.abspath(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = ifl.FluidSolver(n, n, 0.1, variant=ifl.abi.F3_CONJUGATE_GRADIENTS, solver_limit=3, flags=flags)
g.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
g.update(0.005)
for rep in range(3):
    tr = g.wavefront_trace()
CLK = 1.965  # GHz
for sw, name in ((0, "forward (fused r update)"), (1, "backward (fused dot)")):
    t = tr[sw]
    t = t[np.argsort(t[:, 6])]
    t0 = t[:, 0].min()
    start, end = (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    lag = np.diff(start)
    dur = end - start
    print(f"== {name}: total {end.max():.1f} us; strip-to-strip lag median {np.median(lag):.2f} us (sum {lag.sum():.0f}); strip duration "
          f"first {dur[0]:.0f} median {np.median(dur):.0f} last {dur[-1]:.0f} us")
    print("   durations of strips 0..39 (us):", " ".join(f"{d:.0f}" for d in dur[:40]))
    print("   start lags of strips 1..40 (us):", " ".join(f"{d:.1f}" for d in lag[:40]))
    print("   end lags of strips 1..40 (us):", " ".join(f"{d:.1f}" for d in np.diff(end)[:40]))
    print(f"   strip warp, median clocks->us: waiting operand tiles {np.median(t[:,8])/CLK/1e3:.1f}, credit+edge {np.median(t[:,5])/CLK/1e3:.1f}; "
          f"starved macro-steps {np.median(t[:,2]//1000000):.0f}")
    if t[:, 12].max() > 0:
        nt = np.maximum(t[:, 13], 1)
        print(f"   loader warp per tile, median ns: wait bulk copy {np.median(t[:,9]/nt)/CLK:.0f}, r update {np.median(t[:,10]/nt)/CLK:.0f}, "
              f"issue+wait free stage {np.median(t[:,11]/nt)/CLK:.0f}, total {np.median(t[:,12]/nt)/CLK:.0f}")
