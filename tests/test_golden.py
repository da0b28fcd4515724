"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle on the reference's main()
scenes): the oracle must still reproduce them bit for bit (CPU), and the CUDA path in strict mode
(IFL_FLAG_SEQUENTIAL_REDUCTIONS) must reproduce them bit for bit too (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden
import scenes


def load(name):
    return np.load(os.path.join(HERE, "golden", f"{name}_64x64_{make_golden.STEPS}steps.npz"))


def case_list(abi):
    return list(make_golden.cases(abi))


def check(got, want, name):
    assert got["iterations"].tolist() == want["iterations"].tolist(), name
    for k in want.files:
        if k == "iterations":
            continue
        assert np.array_equal(got[k], want[k]), (name, k)


@pytest.mark.parametrize("idx", range(8))
def test_oracle_reproduces_golden(abi, oracle_cls, ifl, idx):
    name, variant, spec, rect, kw = case_list(abi)[idx]

    def factory(variant, spec, kw):
        o = oracle_cls(abi, abi.default_config(variant, make_golden.W, make_golden.H, **kw))
        if spec is not None:
            o.set_bodies(scenes.make_bodies(ifl, spec))
        return o
    check(make_golden.run_case(abi, variant, spec, rect, kw, factory), load(name), name)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(8))
def test_cuda_strict_mode_reproduces_golden(abi, ifl, idx):
    name, variant, spec, rect, kw = case_list(abi)[idx]

    def factory(variant, spec, kw):
        kw = dict(kw, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS)
        if variant >= abi.F6_HEAT:
            return ifl.FluidSolver(make_golden.W, make_golden.H, scenes.DENSITY, kw.pop("density_soot", 0.25), 0.01,
                                   scenes.make_bodies(ifl, spec), variant=variant, **kw)
        if variant >= abi.F4_SOLID_BOUNDARIES:
            return ifl.FluidSolver(make_golden.W, make_golden.H, scenes.DENSITY, scenes.make_bodies(ifl, spec), variant=variant, **kw)
        return ifl.FluidSolver(make_golden.W, make_golden.H, scenes.DENSITY, variant=variant, **kw)
    check(make_golden.run_case(abi, variant, spec, rect, kw, factory), load(name), name)
