"""Scene constants of the reference's main() programs, shared by oracle and parity tests."""
# (x, y, w, h, d, t, u, v) as passed to ifl_add_inflow / FluidSolver.addInflow
F1_INFLOW = (0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 0.0, 3.0)     # ExplicitFluid1matrixless.java:76
F2_INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)    # ExplicitFluid2betterAdvection.java:77
F3_INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)    # ExplicitFluid3conjugateGradients.java:76
DT = 0.005                                                 # F1:63, F3:66
DENSITY = 0.1                                              # F1:62


def inflow_for(abi, variant):
    return F1_INFLOW if variant == abi.F1_MATRIXLESS else F3_INFLOW


def analytic_velocity(w, h, amp=3.0):
    """Smooth divergence-free test field on the MAC grid (SURVEY.md 8d micro-inputs), in grid units."""
    import numpy as np
    hx = 1.0 / min(w, h)
    yu, xu = np.meshgrid((np.arange(h) + 0.5) * hx, np.arange(w + 1) * hx, indexing="ij")
    yv, xv = np.meshgrid(np.arange(h + 1) * hx, (np.arange(w) + 0.5) * hx, indexing="ij")
    u = amp * hx * np.sin(2 * np.pi * xu) * np.cos(2 * np.pi * yu)
    v = -amp * hx * np.cos(2 * np.pi * xv) * np.sin(2 * np.pi * yv)
    yd, xd = np.meshgrid((np.arange(h) + 0.5) * hx, (np.arange(w) + 0.5) * hx, indexing="ij")
    d = np.exp(-((xd - 0.5) ** 2 + (yd - 0.4) ** 2) / 0.02)
    return d.ravel(), u.ravel(), v.ravel()


# ---- F4-F7 scenes (main() of ExplicitFluid4..7): bodies as (kind, args) so both sides build identical ifl_body records
import math
F4_BODIES = [("box", (0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0))]          # F4:80
F5_BODIES = [("sphere", (0.5, 0.6, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0))]            # F5:80
F6_BODIES = [("box", (0.3, 0.6, 0.1, 0.5, -math.pi * 0.05, 0.0, 0.0, 0.0))]         # F6:83
F7_BODIES = [("box", (0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0))]          # F7:83
F4_INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)                             # F4:86
F6_INFLOW = (0.35, 0.9, 0.1, 0.05, 1.0, 294.0 + 300.0, 0.0, 0.0)                    # F6:92
F7_INFLOW = (0.45, 0.2, 0.1, 0.05, 1.0, 294.0, 0.0, 0.0)                            # F7:92
# moving / rotating bodies and two bodies at once: exercises velocityX/Y, the body index and sphere code paths
MIXED_BODIES = [("box", (0.35, 0.55, 0.3, 0.12, 0.3, 0.2, -0.1, 0.7)), ("sphere", (0.7, 0.4, 0.18, 0.1, -0.3, 0.15, -0.4))]


def make_bodies(ifl, spec):
    out = []
    for kind, args in spec:
        cls = ifl.SolidBox if kind == "box" else ifl.SolidSphere
        out.append(cls(*args).as_abi())
    return out


def solid_scene(abi, variant):
    """(bodies, inflow rect, config overrides) of the reference main() for a solid-boundary variant"""
    if variant == abi.F4_SOLID_BOUNDARIES:
        return F4_BODIES, F4_INFLOW, {}
    if variant == abi.F5_CURVED_BOUNDARIES:
        return F5_BODIES, F4_INFLOW, {}
    if variant == abi.F6_HEAT:
        return F6_BODIES, F6_INFLOW, {"density_soot": 0.1}       # F6:77
    return F7_BODIES, F7_INFLOW, {"density_soot": 1.0}           # F7:76
