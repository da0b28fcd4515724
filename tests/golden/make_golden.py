"""Generates tests/golden/*.npz: fields and solver iteration counts produced by the CPU oracle on the reference's
main() scenes (64x64, a few steps).  The reference itself ships no golden vectors and cannot run here (no JVM), so
these pin the ORACLE against regressions and give the GPU tests fixed known answers (strict mode must reproduce them
bit for bit).  Re-run only when the oracle is deliberately changed:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import incfluid_b200 as ifl
import scenes
from oracle_harness import OracleSolver

W = H = 64
STEPS = 6


def cases(abi):
    yield "f1", abi.F1_MATRIXLESS, None, scenes.F1_INFLOW, {}
    yield "f2", abi.F2_BETTER_ADVECTION, None, scenes.F2_INFLOW, {}
    yield "f3", abi.F3_CONJUGATE_GRADIENTS, None, scenes.F3_INFLOW, {}
    for name, v in (("f4", abi.F4_SOLID_BOUNDARIES), ("f5", abi.F5_CURVED_BOUNDARIES), ("f6", abi.F6_HEAT), ("f7", abi.F7_VARIABLE_DENSITY)):
        spec, rect, kw = scenes.solid_scene(abi, v)
        yield name, v, spec, rect, kw
    # F8 main() (F8:72-86): inflow lives inside update(), dt = 0.0025, soot 0.25
    yield "f8", abi.F8_FLIP, scenes.F7_BODIES, None, {}


def run_case(abi, variant, spec, rect, kw, solver_factory):
    s = solver_factory(variant, spec, kw)
    its = []
    flip = variant == abi.F8_FLIP
    for _ in range(STEPS):
        if rect is not None:
            s.add_inflow(*rect)
        st = s.update(0.0025 if flip else scenes.DT)
        its.append((st["iterations_heat"], st["iterations_pressure"]))
    out = {"iterations": np.array(its, dtype=np.int32)}
    if flip:
        for k, v in s.read_particles().items():
            out["particle_" + k] = v
    for n in ["d", "u", "v", "p"] + (["t"] if variant >= abi.F6_HEAT else []):
        out[n] = s.read(n)
    # the frame toImage would write after the last step (F1:341 ... F8:1417), with the heat panel where it exists
    out["image"] = s.image(False)
    if variant >= abi.F6_HEAT:
        out["image_heat"] = s.image(True)
    return out


def main():
    abi = ifl.abi

    def oracle_factory(variant, spec, kw):
        o = OracleSolver(abi, abi.default_config(variant, W, H, **kw))
        if spec is not None:
            o.set_bodies(scenes.make_bodies(ifl, spec))
        return o

    for name, variant, spec, rect, kw in cases(abi):
        out = run_case(abi, variant, spec, rect, kw, oracle_factory)
        np.savez_compressed(os.path.join(HERE, f"{name}_64x64_{STEPS}steps.npz"), **out)
        print(name, out["iterations"].tolist())


if __name__ == "__main__":
    main()
