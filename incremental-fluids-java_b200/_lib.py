"""Loader of the product library libincfluid_b200.so (the C ABI of include/incfluid_b200.h).

The library is built in-tree by csrc/Makefile (nvcc, sm_100a).  There is no fallback of any
kind: if the library is missing or a CUDA device is not usable, calls raise.
"""
import ctypes as C
import os
import subprocess

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libincfluid_b200.so")

_lib = None
_fn = None


class IflError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[ifl status {code}] {message}")
        self.code = code


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libincfluid_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "incfluid_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or \
        any(os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in srcs)
    if stale:
        cmd = ["make", "-C", CSRC, "-j4"] + ([] if verbose else ["-s"])
        if force:
            subprocess.check_call(["make", "-C", CSRC, "-s", "clean"])
        subprocess.check_call(cmd)
    return LIB_PATH


def load():
    """dlopen the product library and bind every prototype of include/incfluid_b200.h."""
    global _lib, _fn
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IflError(_abi.ERR_CUDA, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                          "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _fn = _abi.bind(_lib)
    return _lib, _fn


def check(rc, allow=()):
    if rc != _abi.OK and rc not in allow:
        _, fn = load()
        raise IflError(rc, fn["ifl_last_error"]().decode("utf-8", "replace"))
    return rc
