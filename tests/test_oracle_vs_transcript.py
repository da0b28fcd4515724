"""The C++ oracle against an independent, literal Python transcription of ExplicitFluid1matrixless.java
(tests/java_f1_transcript.py): whole timesteps -- addInflow F1:298, buildRhs F1:364, Gauss-Seidel project F1:380,
applyPressure F1:432, Euler/lerp advect F1:178 -- must agree bit for bit on d, u, v and p, on a ragged grid and on a
square one, and the solver's exit iteration and residual must agree too.  (The reference itself cannot run here.)"""
import numpy as np
import pytest

import java_f1_transcript as jt
import scenes


@pytest.mark.parametrize("w,h,rect", [(24, 20, (0.45, 0.2, 0.15, 0.12, 1.0, 0.0, 3.0)),
                                      (16, 16, (0.3, 0.1, 0.4, 0.2, 1.0, 0.5, 3.0)),
                                      (20, 28, (0.2, 0.2, 0.3, 0.1, 0.7, -1.0, 2.0))])
def test_f1_timesteps_match_python_transcription(ifl, oracle_cls, w, h, rect):
    abi = ifl.abi
    o = oracle_cls(abi, abi.default_config(abi.F1_MATRIXLESS, w, h))
    j = jt.FluidSolver(w, h, 0.1)
    x, y, rw, rh, d, u, v = rect
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, 0.0, u, v)
        j.addInflow(x, y, rw, rh, d, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxDelta, (step, st, j.lastMaxDelta)
        # the oracle reports executed iterations; the reference logs the 0-based index of the converging one (F1:420)
        assert st["iterations_pressure"] in (j.lastIterations, j.lastIterations + 1), (st, j.lastIterations)
    assert np.any(np.array(j._d._src) != 0.0) and np.any(np.array(j._p) != 0.0)


import java_f3_transcript as jt3


@pytest.mark.parametrize("w,h,rect", [(24, 20, (0.45, 0.2, 0.15, 0.12, 1.0, 0.0, 3.0)),
                                      (16, 16, (0.3, 0.1, 0.4, 0.2, 1.0, 0.5, 3.0)),
                                      (20, 28, (0.2, 0.2, 0.3, 0.1, 0.7, -1.0, 2.0))])
def test_f3_timesteps_match_python_transcription(ifl, oracle_cls, w, h, rect):
    """The headline path (configs 2 and 5): cubicPulse inflow F3:230, RK3 + Catmull-Rom advection F3:207/255/185,
    buildPressureMatrix F3:435, MIC(0) buildPreconditioner F3:464, PCG project F3:590 with applyPreconditioner F3:495,
    matrixVectorProduct F3:544, serial dotProduct F3:532, scaledAdd, float-rounded infinityNorm F3:579."""
    abi = ifl.abi
    o = oracle_cls(abi, abi.default_config(abi.F3_CONJUGATE_GRADIENTS, w, h))
    j = jt3.FluidSolver(w, h, 0.1)
    x, y, rw, rh, d, u, v = rect
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, 0.0, u, v)
        j.addInflow(x, y, rw, rh, d, u, v)
        st = o.update(0.005)
        j.update(0.005)
        fields = (("d", j._d._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r),
                  ("precon", j._precon), ("adiag", j._aDiag), ("aplusx", j._aPlusX), ("aplusy", j._aPlusY))
        for name, arr in fields:
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError, (step, st, j.lastMaxError)
        assert j.lastIterations is not None and 1 < j.lastIterations < 600
        assert st["iterations_pressure"] in (j.lastIterations, j.lastIterations + 1), (st, j.lastIterations)
    assert np.any(np.array(j._p) != 0.0) and max(j._d._src) > 0.1 and j.lastIterations > 5


class F2Solver(jt.FluidSolver):
    """ExplicitFluid2betterAdvection.java: its FluidQuantity is F3's and its FluidSolver is F1's minus maxTimestep
    (checked mechanically against the reference: the comment-stripped class bodies are identical), so the transcription
    is the composition of the two."""
    def __init__(self, w, h, density):
        super().__init__(w, h, density)
        self._d = jt3.FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = jt3.FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = jt3.FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)


@pytest.mark.parametrize("w,h,rect", [(24, 20, (0.45, 0.2, 0.15, 0.12, 1.0, 0.0, 3.0)), (16, 22, (0.3, 0.1, 0.4, 0.2, 1.0, 0.5, 3.0))])
def test_f2_timesteps_match_python_transcription(ifl, oracle_cls, w, h, rect):
    abi = ifl.abi
    o = oracle_cls(abi, abi.default_config(abi.F2_BETTER_ADVECTION, w, h))
    j = F2Solver(w, h, 0.1)
    x, y, rw, rh, d, u, v = rect
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, 0.0, u, v)
        j.addInflow(x, y, rw, rh, d, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxDelta


def test_max_timestep_matches_python_transcription(ifl, oracle_cls):
    """maxTimestep F1:309 (lerpGrid at cell centres, sqrt, 2 hx / max, capped at 1)."""
    import math
    abi = ifl.abi
    w, h = 20, 14
    o = oracle_cls(abi, abi.default_config(abi.F1_MATRIXLESS, w, h))
    j = jt.FluidSolver(w, h, 0.1)
    for s in range(2):
        o.add_inflow(0.3, 0.1, 0.4, 0.3, 1.0, 0.0, 0.7, 3.0)
        j.addInflow(0.3, 0.1, 0.4, 0.3, 1.0, 0.7, 3.0)
        o.update(0.005); j.update(0.005)
    maxVelocity = 0.0
    for y in range(h):
        for x in range(w):
            uu = j._u.lerpGrid(x + 0.5, y + 0.5)
            vv = j._v.lerpGrid(x + 0.5, y + 0.5)
            maxVelocity = jt.jmax(maxVelocity, math.sqrt(uu * uu + vv * vv))
    assert o.max_timestep() == jt.jmin(2.0 * j._hx / maxVelocity, 1.0)


def test_solid_bodies_and_fill_solid_fields_match_python_transcription(ifl, oracle_cls):
    """SolidBox / SolidSphere distance and distanceNormal (BOX:81,117; SPH:62,84) through fillSolidFields F4:329 on the
    three staggered quantities, two bodies (a rotated, moving box and a sphere)."""
    import java_solid_transcript as js
    abi = ifl.abi
    w, h = 40, 28
    spec = scenes.MIXED_BODIES
    recs = scenes.make_bodies(ifl, spec)
    bodies = []
    for (kind, args), rec in zip(spec, recs):
        if kind == "box":
            px, py, sx, sy, th, vx, vy, vt = args
            bodies.append(js.SolidBox(px, py, sx, sy, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
        else:
            px, py, s, th, vx, vy, vt = args
            bodies.append(js.SolidSphere(px, py, s, s, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
    o = oracle_cls(abi, abi.default_config(abi.F4_SOLID_BOUNDARIES, w, h))
    o.set_bodies(recs)
    o.run_stage("fill_solid_fields")
    hx = 1.0 / min(w, h)
    for q, (qw, qh, ox, oy) in {"d": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}.items():
        nx, ny = [0.0] * (qw * qh), [0.0] * (qw * qh)
        cell, body = js.fillSolidFields(qw, qh, ox, oy, hx, bodies, nx, ny)
        assert np.array_equal(o.read(f"cell_{q}"), np.array(cell, dtype=np.float64)), q
        assert np.array_equal(o.read(f"body_{q}"), np.array(body, dtype=np.float64)), q
        assert np.array_equal(o.read(f"normalx_{q}"), np.array(nx)), q
        assert np.array_equal(o.read(f"normaly_{q}"), np.array(ny)), q
        assert 0 < sum(cell) < qw * qh and 0 < sum(body) < qw * qh


def test_curved_boundary_volumes_match_python_transcription(ifl, oracle_cls):
    """occupancy / triangleOccupancy / trapezoidOccupancy (F5:135-232) through fillSolidFields F5:463: level set at the
    cell corners, marching-squares volumes (all 16 cases occur with two bodies), the 0.01 cut-off, cell types."""
    import java_solid_transcript as js
    abi = ifl.abi
    w, h = 48, 36
    spec = scenes.MIXED_BODIES
    recs = scenes.make_bodies(ifl, spec)
    bodies = []
    for (kind, args), rec in zip(spec, recs):
        if kind == "box":
            px, py, sx, sy, th, vx, vy, vt = args
            bodies.append(js.SolidBox(px, py, sx, sy, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
        else:
            px, py, s, th, vx, vy, vt = args
            bodies.append(js.SolidSphere(px, py, s, s, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
    o = oracle_cls(abi, abi.default_config(abi.F5_CURVED_BOUNDARIES, w, h))
    o.set_bodies(recs)
    o.run_stage("fill_solid_fields")
    hx = 1.0 / min(w, h)
    seen = set()
    for q, (qw, qh, ox, oy) in {"d": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}.items():
        nx, ny = [0.0] * (qw * qh), [0.0] * (qw * qh)
        phi, volume, cell, body = js.fillSolidFields5(qw, qh, ox, oy, hx, bodies, nx, ny)
        assert np.array_equal(o.read(f"phi_{q}"), np.array(phi)), q
        assert np.array_equal(o.read(f"volume_{q}"), np.array(volume)), q
        assert np.array_equal(o.read(f"cell_{q}"), np.array(cell, dtype=np.float64)), q
        assert np.array_equal(o.read(f"body_{q}"), np.array(body, dtype=np.float64)), q
        assert np.array_equal(o.read(f"normalx_{q}"), np.array(nx)), q
        assert np.array_equal(o.read(f"normaly_{q}"), np.array(ny)), q
        for iy in range(qh):
            for ix in range(qw):
                p = ix + iy * (qw + 1)
                ds = [phi[p], phi[p + 1], phi[p + qw + 2], phi[p + qw + 1]]
                seen.add(sum((1 if ds[i] < 0.0 else 0) << i for i in range(4)))
    assert len(seen) >= 12, sorted(seen)          # the marching-squares cases this scene exercises


def _transcript_bodies(ifl, spec):
    import java_solid_transcript as js
    recs = scenes.make_bodies(ifl, spec)
    bodies = []
    for (kind, args), rec in zip(spec, recs):
        if kind == "box":
            px, py, sx, sy, th, vx, vy, vt = args
            bodies.append(js.SolidBox(px, py, sx, sy, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
        else:
            px, py, s, th, vx, vy, vt = args
            bodies.append(js.SolidSphere(px, py, s, s, th, vx, vy, vt, rec.cos_theta, rec.sin_theta))
    return recs, bodies


@pytest.mark.parametrize("w,h,spec", [(32, 32, scenes.F4_BODIES), (36, 28, scenes.MIXED_BODIES)])
def test_f4_timesteps_match_python_transcription(ifl, oracle_cls, w, h, spec):
    """Whole ExplicitFluid4 timesteps (F4:617): fillSolidFields, setBoundaryCondition (velocityX quirk), masked
    buildRhs / matrix / MIC(0) / PCG, applyPressure, stack-driven extrapolate F4:452, back-projected advect F4:268 --
    with the reference scene's static box and with a moving box + sphere."""
    import java_f4_transcript as j4
    abi = ifl.abi
    recs, bodies = _transcript_bodies(ifl, spec)
    o = oracle_cls(abi, abi.default_config(abi.F4_SOLID_BOUNDARIES, w, h))
    o.set_bodies(recs)
    j = j4.FluidSolver(w, h, 0.1, bodies)
    x, y, rw, rh, d, t, u, v = scenes.F4_INFLOW
    rh = 0.12                                        # tall enough to hit a v face at this resolution
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, t, u, v)
        j.addInflow(x, y, rw, rh, d, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r),
                          ("precon", j._precon), ("adiag", j._aDiag), ("aplusx", j._aPlusX), ("aplusy", j._aPlusY)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError, (step, st, j.lastMaxError)
        assert st["iterations_pressure"] in (j.lastIterations, j.lastIterations + 1), (st, j.lastIterations)
    assert np.any(np.array(j._p) != 0.0) and j._d._cell.count(j4.CELL_SOLID) > 0


@pytest.mark.parametrize("w,h,spec", [(32, 32, scenes.F5_BODIES), (36, 28, scenes.MIXED_BODIES)])
def test_f5_timesteps_match_python_transcription(ifl, oracle_cls, w, h, spec):
    """Whole ExplicitFluid5 timesteps (F5:1089): marching-squares volumes, the volume-weighted right-hand side with the
    body velocities F5:785, the volume-weighted matrix F5:825, everything else as in F4."""
    import java_f5_transcript as j5
    abi = ifl.abi
    recs, bodies = _transcript_bodies(ifl, spec)
    o = oracle_cls(abi, abi.default_config(abi.F5_CURVED_BOUNDARIES, w, h))
    o.set_bodies(recs)
    j = j5.FluidSolver(w, h, 0.1, bodies)
    x, y, rw, rh, d, t, u, v = scenes.F4_INFLOW
    rh = 0.12
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, t, u, v)
        j.addInflow(x, y, rw, rh, d, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r),
                          ("precon", j._precon), ("adiag", j._aDiag), ("aplusx", j._aPlusX), ("aplusy", j._aPlusY),
                          ("volume_u", j._u._volume), ("volume_v", j._v._volume)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError, (step, st, j.lastMaxError)
        assert st["iterations_pressure"] in (j.lastIterations, j.lastIterations + 1), (st, j.lastIterations)
    assert np.any(np.array(j._p) != 0.0) and any(0.0 < vv < 1.0 for vv in j._d._volume)


@pytest.mark.parametrize("w,h,spec,soot", [(32, 32, scenes.F6_BODIES, 0.1), (36, 28, scenes.MIXED_BODIES, 0.3)])
def test_f6_timesteps_match_python_transcription(ifl, oracle_cls, w, h, spec, soot):
    """Whole ExplicitFluid6heat timesteps (F6:1181): implicit heat diffusion solve, buoyancy, fluid-only vector
    operations, then the F5 projection and four advected quantities."""
    import java_f6_transcript as j6
    abi = ifl.abi
    recs, bodies = _transcript_bodies(ifl, spec)
    o = oracle_cls(abi, abi.default_config(abi.F6_HEAT, w, h, density_soot=soot))
    o.set_bodies(recs)
    j = j6.FluidSolver(w, h, 0.1, soot, 0.01, bodies)
    assert o.cfg.t_ambient == j._tAmb and o.cfg.diffusion == 0.01 and o.cfg.gravity == j._g
    x, y, rw, rh, d, t, u, v = scenes.F6_INFLOW
    y, rh = 0.8, 0.12
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, t, u, v)
        j.addInflow(x, y, rw, rh, d, t, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("t", j._t._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError and st["residual_heat"] == j.heatMaxError, (step, st)
        assert st["iterations_heat"] in (j.heatIterations, j.heatIterations + 1), (st, j.heatIterations)
    assert np.any(np.array(j._t._src) != j._tAmb) and np.any(np.array(j._p) != 0.0)


@pytest.mark.parametrize("w,h,spec", [(32, 32, scenes.F7_BODIES), (36, 28, scenes.MIXED_BODIES)])
def test_f7_timesteps_match_python_transcription(ifl, oracle_cls, w, h, spec):
    """Whole ExplicitFluid7variableDensity timesteps (F7:1212): face densities from temperature and soot, the
    density-weighted matrix (index quirk included) and pressure application -- BASELINE config 3's variant."""
    import java_f7_transcript as j7
    abi = ifl.abi
    recs, bodies = _transcript_bodies(ifl, spec)
    o = oracle_cls(abi, abi.default_config(abi.F7_VARIABLE_DENSITY, w, h, density_soot=1.0))
    o.set_bodies(recs)
    j = j7.FluidSolver(w, h, 0.1, 1.0, 0.01, bodies)
    x, y, rw, rh, d, t, u, v = scenes.F7_INFLOW
    rh = 0.12
    for step in range(3):
        o.add_inflow(x, y, rw, rh, d, t + 40.0 * step, u, v)
        j.addInflow(x, y, rw, rh, d, t + 40.0 * step, u, v)
        st = o.update(0.005)
        j.update(0.005)
        for name, arr in (("d", j._d._src), ("t", j._t._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r),
                          ("udensity", j._uDensity), ("vdensity", j._vDensity)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError and st["residual_heat"] == j.heatMaxError, (step, st)
    assert np.any(np.array(j._p) != 0.0) and np.any(np.array(j._uDensity) != j._uDensity[0])


def test_f8_particle_stages_match_python_transcription(ifl, oracle_cls):
    """gridToParticles F8:1748 and the particle advect F8:1778 (rungeKutta3 F8:1708, backProject F8:1673, clamping) in
    isolation, on particles and grid fields written through the ABI (no random stream involved)."""
    import java_f8_transcript as j8
    import java_f3_transcript as j3
    abi = ifl.abi
    w, h = 28, 24
    hx = 1.0 / min(w, h)
    recs, bodies = _transcript_bodies(ifl, scenes.F7_BODIES)
    o = oracle_cls(abi, abi.default_config(abi.F8_FLIP, w, h))
    o.set_bodies(recs)
    rng = np.random.default_rng(5)
    n = 400
    parts = {"x": rng.uniform(-0.5, w + 0.5, n), "y": rng.uniform(-0.5, h + 0.5, n)}
    for k in ("d", "t", "u", "v"):
        parts[k] = rng.normal(size=n)
    grids = {"d": (w, h, 0.5, 0.5), "t": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}
    qs = {}
    for k, (qw, qh, ox, oy) in grids.items():
        q = j3.FluidQuantity(qw, qh, ox, oy, hx)
        q._src = list(rng.normal(size=qw * qh) * (0.05 if k in "uv" else 1.0))
        o.write(k, np.array(q._src))
        qs[k] = q
    o.write_particles(parts)
    # grid -> particles (the stage also applies diff / undiff with _old = 0, which leaves the grids unchanged)
    o.run_stage("grid_to_particles")
    props = [list(parts[k]) for k in ("d", "t", "u", "v")]
    px, py = list(parts["x"]), list(parts["y"])
    j8.gridToParticles([qs[k] for k in ("d", "t", "u", "v")], props, px, py, n, o.cfg.flip_alpha)
    got = o.read_particles()
    for k, p in zip(("d", "t", "u", "v"), props):
        assert np.array_equal(got[k], np.array(p)), k
    # particle advection
    o.run_stage("advect_particles", 0.0025)
    j8.advect(px, py, n, 0.0025, qs["u"], qs["v"], bodies, w, h, hx)
    got = o.read_particles()
    assert np.array_equal(got["x"], np.array(px)) and np.array_equal(got["y"], np.array(py))
    assert np.any(np.array(px) != parts["x"])


def test_f8_particles_to_grid_matches_python_transcription(ifl, oracle_cls):
    """fromParticles F8:722 / addSample F8:312: the bilinear splat in particle-index order and the division by the
    accumulated weights.  Particles cover every node and there are no bodies, so the extrapolation that follows the
    transfer inside particlesToGrid F8:1761 has nothing to fill and the stage's grids are the transfer's output."""
    import java_f8_transcript as j8
    abi = ifl.abi
    w, h = 20, 16
    o = oracle_cls(abi, abi.default_config(abi.F8_FLIP, w, h))
    rng = np.random.default_rng(11)
    n = w * h * 6
    parts = {"x": rng.uniform(0.0, w, n), "y": rng.uniform(0.0, h, n)}
    for k in ("d", "t", "u", "v"):
        parts[k] = rng.normal(size=n)
    o.write_particles(parts)
    o.run_stage("particles_to_grid")
    grids = {"d": (w, h, 0.5, 0.5), "t": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}
    for k, (qw, qh, ox, oy) in grids.items():
        cell = [j8.CELL_FLUID] * (qw * qh)
        want = j8.fromParticles(qw, qh, ox, oy, cell, n, list(parts["x"]), list(parts["y"]), list(parts[k]))
        assert cell.count(j8.CELL_EMPTY) == 0, k            # the premise of this test
        assert np.array_equal(o.read(k), np.array(want)), k


def test_f8_empty_cell_extrapolation_matches_python_transcription(ifl, oracle_cls):
    """Sparse particles leave EMPTY cells (fromParticles F8:745): the transfer followed by FluidQuantity.extrapolate
    F8:651 -- averages of FLUID neighbours in stack order, chains of EMPTY cells, domain borders and corners."""
    import java_f8_transcript as j8
    abi = ifl.abi
    w, h = 24, 20
    o = oracle_cls(abi, abi.default_config(abi.F8_FLIP, w, h))
    rng = np.random.default_rng(3)
    # a blob that covers the middle of the domain only, with holes
    n = 500
    parts = {"x": rng.uniform(3.0, w - 2.0, n), "y": rng.uniform(0.0, h - 5.0, n)}
    for k in ("d", "t", "u", "v"):
        parts[k] = rng.normal(size=n)
    o.write_particles(parts)
    o.run_stage("particles_to_grid")
    grids = {"d": (w, h, 0.5, 0.5), "t": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}
    for k, (qw, qh, ox, oy) in grids.items():
        cell = [j8.CELL_FLUID] * (qw * qh)
        src = j8.fromParticles(qw, qh, ox, oy, cell, n, list(parts["x"]), list(parts["y"]), list(parts[k]))
        empties = cell.count(j8.CELL_EMPTY)
        assert 20 < empties < qw * qh - 20, (k, empties)
        j8.extrapolate8(src, cell, [0.0] * (qw * qh), [0.0] * (qw * qh), qw, qh)
        a, b = o.read(k), np.array(src)
        bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
        assert bad.size == 0, (k, bad[:8], a[bad[:4]], b[bad[:4]])


def test_f8_particle_bookkeeping_matches_python_transcription(ifl, oracle_cls):
    """countParticles / pruneParticles / seedParticles (F8:1595-1660) after the transfer, on the injected random stream:
    over-full cells lose particles (swap with the last), under-full cells are re-seeded from the grids."""
    import java_f8_transcript as j8
    import java_f3_transcript as j3
    abi = ifl.abi
    w, h = 14, 12
    cfg = abi.default_config(abi.F8_FLIP, w, h)
    o = oracle_cls(abi, cfg)
    rng = np.random.default_rng(9)
    # 40 particles in each of a few cells (pruned to 12), 0-2 in most others (seeded up to 3)
    xs, ys = [], []
    for cx, cy in ((3, 4), (7, 7), (10, 2)):
        xs += list(cx + rng.uniform(0.0, 1.0, 40)); ys += list(cy + rng.uniform(0.0, 1.0, 40))
    xs += list(rng.uniform(0.0, w, 150)); ys += list(rng.uniform(0.0, h, 150))
    n = len(xs)
    parts = {"x": np.array(xs), "y": np.array(ys)}
    for k in ("d", "t", "u", "v"):
        parts[k] = rng.normal(size=n)
    o.write_particles(parts)
    o.run_stage("particles_to_grid")
    # transcription: transfer + extrapolation per quantity, then the bookkeeping
    hx = 1.0 / min(w, h)
    grids = {"d": (w, h, 0.5, 0.5), "t": (w, h, 0.5, 0.5), "u": (w + 1, h, 0.0, 0.5), "v": (w, h + 1, 0.5, 0.0)}
    maxp = w * h * cfg.particles_max_per_cell
    px, py = xs + [0.0] * (maxp - n), ys + [0.0] * (maxp - n)
    props, quantities = [], []
    for k, (qw, qh, ox, oy) in grids.items():
        cell = [j8.CELL_FLUID] * (qw * qh)
        src = j8.fromParticles(qw, qh, ox, oy, cell, n, px, py, list(parts[k]))
        j8.extrapolate8(src, cell, [0.0] * (qw * qh), [0.0] * (qw * qh), qw, qh)
        q = j3.FluidQuantity(qw, qh, ox, oy, hx)
        q._src = src
        quantities.append(q)
        props.append(list(parts[k]) + [0.0] * (maxp - n))
        assert np.array_equal(o.read(k), np.array(src)), k
    counts = j8.countParticles(w, h, n, px, py)
    assert max(counts) > cfg.particles_max_per_cell and min(counts) < cfg.particles_min_per_cell
    count = j8.pruneParticles(w, h, n, px, py, props, counts, cfg.particles_max_per_cell)
    assert count < n
    rnd = j8.InjectedRandom(cfg.rng_seed)
    count2 = j8.seedParticles(w, h, count, maxp, px, py, props, quantities, counts, cfg.particles_min_per_cell, rnd,
                              lambda x, y: False)
    assert count2 > count
    got = o.read_particles()
    assert got["x"].size == count2
    assert np.array_equal(got["x"], np.array(px[:count2])) and np.array_equal(got["y"], np.array(py[:count2]))
    for k, p in zip(("d", "t", "u", "v"), props):
        assert np.array_equal(got[k], np.array(p[:count2])), k


def test_f8_timesteps_match_python_transcription(ifl, oracle_cls):
    """Whole ExplicitFluid8flip timesteps (F8:1306) on the injected random stream: particle init, transfer, EMPTY-cell
    extrapolation, prune / seed, inflow inside update(), heat solve, variable-density projection, PIC/FLIP blend and
    particle advection -- BASELINE config 4's program."""
    import java_f8_transcript as j8
    abi = ifl.abi
    w, h = 44, 40                                    # fine enough for the inflow rectangle of F8:1330 to hit cell centres
    cfg = abi.default_config(abi.F8_FLIP, w, h)
    recs, bodies = _transcript_bodies(ifl, scenes.F7_BODIES)
    o = oracle_cls(abi, cfg)
    o.set_bodies(recs)
    rnd = j8.InjectedRandom(cfg.rng_seed)
    j = j8.FluidSolver(w, h, cfg.density, cfg.density_soot, cfg.diffusion, bodies, rnd, cfg.particles_avg_per_cell,
                       cfg.particles_max_per_cell, cfg.particles_min_per_cell)
    assert cfg.flip_alpha == j._flipAlpha
    for step in range(6):
        st = o.update(0.0025)
        j.update(0.0025)
        got = o.read_particles()
        cnt = j._qs._particleCount
        assert got["x"].size == cnt, (step, got["x"].size, cnt)
        assert np.array_equal(got["x"], np.array(j._qs._posX[:cnt])), step
        assert np.array_equal(got["y"], np.array(j._qs._posY[:cnt])), step
        for k, p in zip(("d", "t", "u", "v"), j._qs._properties):
            assert np.array_equal(got[k], np.array(p[:cnt])), (step, k)
        for name, arr in (("d", j._d._src), ("t", j._t._src), ("u", j._u._src), ("v", j._v._src), ("p", j._p), ("r", j._r)):
            a, b = o.read(name), np.array(arr)
            bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
            assert bad.size == 0, (step, name, bad[:5], a[bad[:3]], b[bad[:3]])
        assert st["residual_pressure"] == j.lastMaxError and st["residual_heat"] == j.heatMaxError, (step, st)
    assert np.any(np.array(j._p) != 0.0) and max(j._d._src) > 0.1 and j.lastIterations > 5
