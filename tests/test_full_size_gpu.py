"""BASELINE.json configs at their FULL sizes, checked through size-independent properties (the oracle cannot run
these sizes in seconds): configs[2] F7 4096x4096, configs[3] F8 2048x2048 with 8 particles/cell, and a 16384-wide F3
projection with a capped iteration count.  Iteration caps keep the whole file around a minute on a B200."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def test_config3_f7_4096_solids_heat_variable_density(ifl):
    abi = ifl.abi
    n = 4096
    g = ifl.FluidSolver(n, n, scenes.DENSITY, 1.0, 0.01, scenes.make_bodies(ifl, scenes.F7_BODIES),
                        variant=abi.F7_VARIABLE_DENSITY, solver_limit=40)
    for _ in range(2):
        g.add_inflow(*scenes.F7_INFLOW)
        st = g.update(scenes.DT)
    assert st["nan_residual"] == 0 and st["iterations_pressure"] == 40 and st["exceeded_budget"] & 1
    cell = g.read("cell_d")
    vol = g.read("volume_d")
    # the rotated 0.7 x 0.1 box: solid fraction of the unit square (marching squares integrates the SDF)
    assert abs(np.sum(1.0 - vol) / (n * n) - 0.07) < 1e-4
    assert np.all((cell == 1) == (vol == 0.0))
    d, t = g.read("d"), g.read("t")
    assert np.all(np.isfinite(d)) and np.all(np.isfinite(t)) and np.all(np.isfinite(g.read("u")))
    assert d.min() >= 0.0 and d.max() <= 1.0 + 1e-12          # Catmull-Rom is clamped to its samples (INT:151)
    # inflow at ambient temperature: nothing heats up (the heat solve is capped at 40 iterations here, so T is
    # only converged to a fraction of a kelvin)
    assert abs(t[cell == 0].mean() - 294.0) < 0.05 and np.max(np.abs(t[cell == 0] - 294.0)) < 30.0
    # domain-edge faces are zeroed by setBoundaryCondition (F7:1201-1208) and stay zero through advection of the edge rows
    g.run_stage("set_boundary")
    u = g.read("u").reshape(n, n + 1)
    assert np.all(u[:, 0] == 0.0) and np.all(u[:, -1] == 0.0)


def test_config4_f8_2048_flip_8_particles_per_cell(ifl):
    abi = ifl.abi
    n = 2048
    g = ifl.FluidSolver(n, n, scenes.DENSITY, 0.25, 0.01, scenes.make_bodies(ifl, scenes.F7_BODIES),
                        variant=abi.F8_FLIP, solver_limit=30, particles_avg_per_cell=8)
    n0 = g.particle_count()
    # 8 tries per cell, the ones inside the box (7 % of the area) are dropped (F8:1575-1585)
    assert abs(n0 / (8.0 * n * n) - 0.93) < 2e-3
    st = None
    for _ in range(2):
        st = g.update(0.0025)
    assert st["nan_residual"] == 0
    assert st["particle_count"] >= n0                          # nothing pruned in two steps, holes are re-seeded
    p = g.read_particles()
    assert np.all((p["x"] >= 0) & (p["x"] <= n - 0.001) & (p["y"] >= 0) & (p["y"] <= n - 0.001))   # F8:1786-1787
    # P2G of a constant property returns the constant on every non-solid cell
    p["t"][:] = 300.0
    g.write_particles(p)
    g.run_stage("fill_solid_fields")
    g.run_stage("particles_to_grid")
    t = g.read("t"); cell = g.read("cell_t")
    assert np.max(np.abs(t[cell == 0] - 300.0)) < 1e-9


def test_config5_width_16384_projection_reduces_divergence(ifl):
    """16384 columns (the slab benchmark's width) on a 2048-row grid: 64 strips, 2 clusters; a capped solve must
    still monotonically reduce the residual and never produce NaN."""
    abi = ifl.abi
    w, h = 16384, 2048
    res = []
    for lim in (5, 40):
        g = ifl.FluidSolver(w, h, scenes.DENSITY, variant=abi.F3_CONJUGATE_GRADIENTS, solver_limit=lim)
        g.add_inflow(*scenes.F3_INFLOW)
        st = g.update(scenes.DT)
        assert st["nan_residual"] == 0 and st["iterations_pressure"] == lim
        res.append(st["residual_pressure"])
        g.close()
    assert res[1] < res[0]
