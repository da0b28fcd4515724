"""Test-side driver of the CPU oracle (oracle/libifl_oracle.so).

The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs load it.  It exposes the same create / add_inflow /
update / run_stage / read_field / write_field shape as the product's C ABI (prefix ifo_),
so a parity test drives both sides with identical calls.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libifl_oracle.so")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "ifl_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


def _load(abi):
    lib = C.CDLL(build_oracle())
    dp = C.POINTER(C.c_double)
    protos = {
        "ifo_create": (C.c_int, [C.POINTER(abi.Config), C.POINTER(C.c_void_p)]),
        "ifo_destroy": (None, [C.c_void_p]),
        "ifo_set_bodies": (C.c_int, [C.c_void_p, C.POINTER(abi.Body), C.c_int32]),
        "ifo_add_inflow": (C.c_int, [C.c_void_p] + [C.c_double] * 8),
        "ifo_update": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(abi.StepStatus)]),
        "ifo_run_steps": (C.c_int, [C.c_void_p, C.c_int32, C.c_double] + [C.c_double] * 8 + [C.POINTER(abi.StepStatus)]),
        "ifo_run_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.POINTER(abi.StepStatus)]),
        "ifo_max_timestep": (C.c_int, [C.c_void_p, dp]),
        "ifo_image_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]),
        "ifo_to_image": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_size_t]),
        "ifo_write_ppm": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
        "ifo_field_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]),
        "ifo_read_field": (C.c_int, [C.c_void_p, C.c_int32, dp, C.c_size_t]),
        "ifo_write_field": (C.c_int, [C.c_void_p, C.c_int32, dp, C.c_size_t]),
        "ifo_particle_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
        "ifo_read_particles": (C.c_int, [C.c_void_p] + [dp] * 6 + [C.c_size_t]),
        "ifo_write_particles": (C.c_int, [C.c_void_p] + [dp] * 6 + [C.c_size_t]),
        "ifo_lerp": (C.c_double, [C.c_double] * 3),
        "ifo_cerp": (C.c_double, [C.c_double] * 5),
        "ifo_cubic_pulse": (C.c_double, [C.c_double]),
        "ifo_dot": (C.c_double, [C.c_void_p, C.c_int, C.c_int]),
        "ifo_inf_norm": (C.c_double, [C.c_void_p, C.c_int]),
    }
    for name, (res, args) in protos.items():
        if not hasattr(lib, name):
            continue  # optional symbols (added as the oracle widens)
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


class OracleSolver:
    """Oracle-side twin of the product's FluidSolver (same method names)."""

    def __init__(self, abi, cfg):
        self.abi = abi
        self.lib = _load(abi)
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.lib.ifo_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ifo_create failed rc={rc}")
        self.h = h
        self.status = abi.StepStatus()

    def close(self):
        if self.h:
            self.lib.ifo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bodies(self, bodies):
        arr = (self.abi.Body * max(len(bodies), 1))(*bodies)
        rc = self.lib.ifo_set_bodies(self.h, arr, len(bodies))
        assert rc == 0, rc

    def add_inflow(self, x, y, w, h, d, t, u, v):
        rc = self.lib.ifo_add_inflow(self.h, x, y, w, h, d, t, u, v)
        assert rc == 0, rc

    def update(self, dt):
        rc = self.lib.ifo_update(self.h, dt, C.byref(self.status))
        assert rc in (0, self.abi.ERR_NAN), rc
        return self.status.as_dict()

    def run_steps(self, n, dt, rect):
        rc = self.lib.ifo_run_steps(self.h, n, dt, *rect, C.byref(self.status))
        assert rc in (0, self.abi.ERR_NAN), rc
        return self.status.as_dict()

    def run_stage(self, stage, dt=0.0):
        rc = self.lib.ifo_run_stage(self.h, self.abi.STAGES[stage], dt, C.byref(self.status))
        assert rc == 0, (stage, rc)
        return self.status.as_dict()

    def max_timestep(self):
        out = C.c_double()
        assert self.lib.ifo_max_timestep(self.h, C.byref(out)) == 0
        return out.value

    def to_image(self, destination, render_heat=False):
        rc = self.lib.ifo_write_ppm(self.h, os.fsencode(destination), 1 if render_heat else 0)
        assert rc == 0, rc

    def image(self, render_heat=False):
        n = C.c_size_t()
        assert self.lib.ifo_image_size(self.h, 1 if render_heat else 0, C.byref(n)) == 0
        out = np.empty(3 * n.value, dtype=np.int32)
        rc = self.lib.ifo_to_image(self.h, 1 if render_heat else 0, out.ctypes.data_as(C.POINTER(C.c_int32)), n.value)
        assert rc == 0, rc
        return out.reshape(self.cfg.h, -1, 3)

    def field_size(self, name):
        n = C.c_size_t()
        rc = self.lib.ifo_field_size(self.h, self.abi.FIELDS[name], C.byref(n))
        assert rc == 0, (name, rc)
        return n.value

    def read(self, name):
        n = self.field_size(name)
        out = np.empty(n, dtype=np.float64)
        rc = self.lib.ifo_read_field(self.h, self.abi.FIELDS[name], out.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert rc == 0, (name, rc)
        return out

    def write(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        rc = self.lib.ifo_write_field(self.h, self.abi.FIELDS[name], arr.ctypes.data_as(C.POINTER(C.c_double)), arr.size)
        assert rc == 0, (name, rc)

    def particle_count(self):
        n = C.c_int32()
        assert self.lib.ifo_particle_count(self.h, C.byref(n)) == 0
        return n.value

    def read_particles(self):
        n = self.particle_count()
        arrs = [np.empty(n, dtype=np.float64) for _ in range(6)]
        dp = C.POINTER(C.c_double)
        rc = self.lib.ifo_read_particles(self.h, *[a.ctypes.data_as(dp) for a in arrs], n)
        assert rc == 0, rc
        return dict(zip(("x", "y", "d", "t", "u", "v"), arrs))

    def write_particles(self, parts):
        dp = C.POINTER(C.c_double)
        arrs = [np.ascontiguousarray(parts[k], dtype=np.float64) for k in ("x", "y", "d", "t", "u", "v")]
        rc = self.lib.ifo_write_particles(self.h, *[a.ctypes.data_as(dp) for a in arrs], arrs[0].size)
        assert rc == 0, rc

    def dot(self, a, b):
        return self.lib.ifo_dot(self.h, self.abi.FIELDS[a], self.abi.FIELDS[b])

    def inf_norm(self, a):
        return self.lib.ifo_inf_norm(self.h, self.abi.FIELDS[a])
