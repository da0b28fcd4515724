"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/incfluid_b200.h declares, agrees with the ctypes struct layout, and fails loudly (no CPU
fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "incfluid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ifl_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(ifl):
    lib, fn = ifl.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/incfluid_b200.h but not exported"
        assert s in ifl.abi.PROTOTYPES, f"{s} has no ctypes prototype in _abi.py"


def test_default_config_matches_python_mirror(ifl):
    lib, fn = ifl.load()
    abi = ifl.abi
    for variant in range(1, 9):
        c = abi.Config()
        assert fn["ifl_default_config"](C.byref(c), variant, 128, 96) == abi.OK
        p = abi.default_config(variant, 128, 96)
        for name, _ in abi.Config._fields_:
            if name == "comm_id":
                continue
            assert getattr(c, name) == getattr(p, name), (variant, name)
    assert fn["ifl_default_config"](C.byref(abi.Config()), 99, 8, 8) == abi.ERR_INVALID


def test_abi_version_and_errors(ifl):
    lib, fn = ifl.load()
    assert fn["ifl_abi_version"]() == ifl.abi.ABI_VERSION
    h = C.c_void_p()
    bad = ifl.abi.default_config(ifl.abi.F1_MATRIXLESS, 1, 1)
    assert fn["ifl_create"](C.byref(bad), C.byref(h)) == ifl.abi.ERR_INVALID
    assert b"2x2" in fn["ifl_last_error"]()
    assert fn["ifl_update"](None, 0.005, None) == ifl.abi.ERR_INVALID


def test_no_cpu_fallback_without_gpu(ifl):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device error path is exercised on the CPU box")
    with pytest.raises(ifl.IflError) as e:
        ifl.FluidSolver(16, 16, 0.1)
    assert e.value.code == ifl.abi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_loads_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "incremental-fluids-java_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
