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
