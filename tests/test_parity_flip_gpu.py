"""GPU parity for F8 (FLIP/PIC, BASELINE config 4 family) against the CPU oracle.

Math.random() is replaced on both sides by the same injected counter-based stream (cfg.rng_seed), particles are
compared in storage order (the library keeps the reference's particle order, including prune's swap-with-last).
Strict mode (sequential dot products) is bit-exact for whole timesteps; the fast path is compared per step from a
re-synchronised state like the other solid variants."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def pair(ifl, oracle_cls, w, h, spec, **kw):
    abi = ifl.abi
    bodies = scenes.make_bodies(ifl, spec)
    cfg = dict(kw)
    g = ifl.FluidSolver(w, h, scenes.DENSITY, cfg.pop("density_soot", 0.25), cfg.pop("diffusion", 0.01), bodies,
                        variant=abi.F8_FLIP, **cfg)
    o = oracle_cls(abi, abi.default_config(abi.F8_FLIP, w, h, **kw))
    o.set_bodies(bodies)
    return g, o


def assert_particles_equal(g, o, what=""):
    pg, po = g.read_particles(), o.read_particles()
    assert pg["x"].size == po["x"].size, (what, pg["x"].size, po["x"].size)
    for k in ("x", "y", "d", "t", "u", "v"):
        bad = np.flatnonzero(pg[k] != po[k])
        assert bad.size == 0, f"{what} particle {k}: {bad.size} differ, first {bad[:5]} gpu {pg[k][bad[:3]]} ref {po[k][bad[:3]]}"


def assert_fields_equal(g, o, names, what=""):
    for n in names:
        a, b = g.read(n), o.read(n)
        bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
        assert bad.size == 0, f"{what} {n}: {bad.size} cells differ, first {bad[:5]} gpu {a[bad[:3]]} ref {b[bad[:3]]}"


@pytest.mark.parametrize("w,h", [(64, 64), (96, 40)])
def test_init_particles_and_first_g2p(ifl, oracle_cls, w, h):
    g, o = pair(ifl, oracle_cls, w, h, scenes.F7_BODIES)
    assert g.particle_count() == o.particle_count()
    assert_particles_equal(g, o, "init")


@pytest.mark.parametrize("flags", [0, 2])
def test_stages_bit_exact(ifl, oracle_cls, flags):
    """particlesToGrid (P2G + extrapolate + count/prune/seed), gridToParticles, particle advect in isolation, with
    holes punched into the particle set so that EMPTY cells (isolated and adjacent), pruning and seeding all occur."""
    abi = ifl.abi
    w, h = 64, 48
    g, o = pair(ifl, oracle_cls, w, h, scenes.MIXED_BODIES, flags=flags)
    rng = np.random.default_rng(11)
    assert g.particle_count() == o.particle_count()     # both sides have consumed the init draws of the random stream
    parts = o.read_particles()
    n = parts["x"].size
    keep = np.ones(n, dtype=bool)
    cx, cy = parts["x"].astype(int), parts["y"].astype(int)
    keep &= ~((cx >= 10) & (cx < 14) & (cy >= 8) & (cy < 11))          # a 4x3 hole: adjacent EMPTY cells
    keep &= ~((cx == 40) & (cy == 30))                                   # an isolated EMPTY cell
    parts = {k: v[keep] for k, v in parts.items()}
    # crowd one cell so that pruneParticles has work: 20 extra particles in cell (30, 20)
    extra = 20
    parts["x"] = np.concatenate([parts["x"], 30 + rng.uniform(0, 1, extra)])
    parts["y"] = np.concatenate([parts["y"], 20 + rng.uniform(0, 1, extra)])
    m = parts["x"].size
    for k in ("d", "t", "u", "v"):
        base = {"d": 0.5, "t": 300.0, "u": 0.01, "v": -0.02}[k]
        parts[k] = base * (1.0 + 0.1 * rng.normal(size=m))
    perm = rng.permutation(m)                                           # storage order is arbitrary in general
    parts = {k: v[perm] for k, v in parts.items()}
    for s in (g, o):
        s.write_particles(parts)
        s.run_stage("fill_solid_fields")
        s.run_stage("particles_to_grid")
    assert_fields_equal(g, o, ["d", "t", "u", "v", "cell_d", "cell_u", "cell_v"], "particlesToGrid")
    assert_particles_equal(g, o, "after prune/seed")
    u = rng.normal(size=(w + 1) * h) * 1e-2
    v = rng.normal(size=w * (h + 1)) * 1e-2
    for s in (g, o):
        s.write("u", u); s.write("v", v)
        s.run_stage("grid_to_particles")
    assert_particles_equal(g, o, "gridToParticles")
    assert_fields_equal(g, o, ["d", "t", "u", "v"], "diff/undiff")
    for s in (g, o):
        s.run_stage("advect_particles", 0.0025)
    assert_particles_equal(g, o, "advect")


@pytest.mark.parametrize("w,h", [(64, 64), (96, 40)])
def test_timesteps_strict_bit_exact(ifl, oracle_cls, w, h):
    abi = ifl.abi
    g, o = pair(ifl, oracle_cls, w, h, scenes.F7_BODIES, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS)
    for step in range(4):
        sg, so = g.update(0.0025), o.update(0.0025)
        assert (sg["iterations_heat"], sg["iterations_pressure"], sg["particle_count"]) == \
               (so["iterations_heat"], so["iterations_pressure"], so["particle_count"]), (step, sg, so)
        assert_fields_equal(g, o, ["d", "t", "u", "v", "p"], f"step {step}")
        assert_particles_equal(g, o, f"step {step}")


def test_reference_scene_fast_path(ifl, oracle_cls):
    """F8 main() scene at 128x128 (F8:72-86), 12 steps, fast path, per-step re-sync of particles (the whole state of
    a FLIP solver): iteration counts within one, particles within 1e-5 of the field scales (see
    test_parity_solids_gpu.py for why the fast path is held to the solver tolerance, not to 1e-12)."""
    g, o = pair(ifl, oracle_cls, 128, 128, scenes.F7_BODIES)
    assert g.particle_count() == o.particle_count()     # both sides have consumed the init draws of the random stream
    for step in range(12):
        g.write_particles(o.read_particles())
        for n in ("d", "t", "u", "v"):            # _src of step n-1 only matters through `_old`; keep both sides identical
            g.write(n, o.read(n))
        a, b = g.update(0.0025), o.update(0.0025)
        assert abs(a["iterations_heat"] - b["iterations_heat"]) <= 1 and abs(a["iterations_pressure"] - b["iterations_pressure"]) <= 1
        assert a["particle_count"] == b["particle_count"]
        pg, po = g.read_particles(), o.read_particles()
        vel = max(np.max(np.abs(po["u"])), np.max(np.abs(po["v"])), 1e-300)
        assert np.max(np.abs(pg["x"] - po["x"])) <= 1e-7 and np.max(np.abs(pg["y"] - po["y"])) <= 1e-7
        assert np.max(np.abs(pg["u"] - po["u"])) / vel <= 1e-5 and np.max(np.abs(pg["v"] - po["v"])) / vel <= 1e-5
        assert np.max(np.abs(pg["d"] - po["d"])) <= 1e-7 and np.max(np.abs(pg["t"] - po["t"])) / 294.0 <= 1e-7
