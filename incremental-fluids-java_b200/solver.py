"""Host-side mirror of the reference's Java API for the timestep hot path.

Class and method names follow the reference (physics.EulerianFluids.ExplicitFluid<N>*.FluidSolver /
FluidQuantity, physics.solidBodies.SolidBox / SolidSphere): `FluidSolver(w, h, density)`,
`addInflow(...)`, `update(timestep)`, `maxTimestep()`; field state is read like the Java code reads
`solver._d._src[i]`.  Every method is a thin call into the C ABI (include/incfluid_b200.h) -- the same
entry points a Java FluidSolver binds with Panama FFM.  No arithmetic happens here.
"""
import ctypes as C
import os
import math

import numpy as np

from . import _abi
from ._lib import load, check, IflError  # noqa: F401

_dp = C.POINTER(C.c_double)


class SolidBody:
    """A_SolidBody (solidBodies/A_SolidBody.java:73-87) as a plain record."""
    TYPE = None

    def __init__(self, posX, posY, scaleX, scaleY, theta, velX, velY, velTheta):
        self._posX, self._posY = posX, posY
        self._scaleX, self._scaleY = scaleX, scaleY
        self._theta = theta
        self._velX, self._velY, self._velTheta = velX, velY, velTheta

    def update(self, timestep):
        """A_SolidBody.update (A_SolidBody.java:215): explicit Euler pose update."""
        self._posX += self._velX * timestep
        self._posY += self._velY * timestep
        self._theta += self._velTheta * timestep

    def as_abi(self):
        b = _abi.Body()
        b.type = self.TYPE
        b.pos_x, b.pos_y = self._posX, self._posY
        b.scale_x, b.scale_y = self._scaleX, self._scaleY
        b.theta = self._theta
        b.cos_theta, b.sin_theta = math.cos(self._theta), math.sin(self._theta)
        b.vel_x, b.vel_y, b.vel_theta = self._velX, self._velY, self._velTheta
        return b


class SolidBox(SolidBody):
    """SolidBox(x, y, sx, sy, theta, vx, vy, vt) -- solidBodies/SolidBox.java:75"""
    TYPE = _abi.BODY_BOX


class SolidSphere(SolidBody):
    """SolidSphere(x, y, s, theta, vx, vy, vt) -- solidBodies/SolidSphere.java:57"""
    TYPE = _abi.BODY_SPHERE

    def __init__(self, x, y, s, theta, velX, velY, velTheta):
        super().__init__(x, y, s, s, theta, velX, velY, velTheta)


def comm_unique_id():
    """128-byte id for a slab-decomposed solver group (rank 0 creates it, the host broadcasts it)."""
    _, fn = load()
    buf = (C.c_uint8 * 128)()
    check(fn["ifl_comm_unique_id"](buf))
    return bytes(buf)


class FluidQuantity:
    """View of one device-resident FluidQuantity (F1:105): `_src` / `_dst` copy the buffers out."""

    def __init__(self, solver, name, w, h, ox, oy):
        self._solver, self._name = solver, name
        self._w, self._h, self._ox, self._oy = w, h, ox, oy
        self._hx = solver._hx

    @property
    def _src(self):
        return self._solver.read_field(self._name)

    @_src.setter
    def _src(self, values):
        self._solver.write_field(self._name, values)

    @property
    def _dst(self):
        return self._solver.read_field(self._name + "_dst")


class FluidSolver:
    """FluidSolver of ExplicitFluid<variant>: same constructor arguments and methods as the reference.

    FluidSolver(w, h, density)                         F1:262, F2:305, F3:339
    FluidSolver(w, h, density, bodies)                 F4:594, F5:707
    FluidSolver(w, h, rhoAir, rhoSoot, diffusion, bodies)  F6:763, F7:769, F8:858
    """

    def __init__(self, w, h, density, *rest, variant=_abi.F3_CONJUGATE_GRADIENTS, device=0, **overrides):
        _, self._fn = load()
        bodies = None
        cfg_kw = dict(density=density, device=device)
        if variant in (_abi.F4_SOLID_BOUNDARIES, _abi.F5_CURVED_BOUNDARIES):
            (bodies,) = rest
        elif variant >= _abi.F6_HEAT:
            rho_soot, diffusion, bodies = rest
            cfg_kw.update(density_soot=rho_soot, diffusion=diffusion)
        elif rest:
            raise TypeError("FluidSolver(w, h, density) takes no further positional arguments for F1-F3")
        comm_id = overrides.pop("comm_id", None)
        cfg_kw.update(overrides)
        self.cfg = _abi.default_config(variant, w, h, **cfg_kw)
        if comm_id is not None:
            C.memmove(self.cfg.comm_id, comm_id, 128)
        self._w, self._h = w, h
        self._hx = 1.0 / min(w, h)
        self._variant = variant
        self._h_ = C.c_void_p()
        check(self._fn["ifl_create"](C.byref(self.cfg), C.byref(self._h_)))
        self.status = _abi.StepStatus()
        self._d = FluidQuantity(self, "d", w, h, 0.5, 0.5)
        self._u = FluidQuantity(self, "u", w + 1, h, 0.0, 0.5)
        self._v = FluidQuantity(self, "v", w, h + 1, 0.5, 0.0)
        if variant >= _abi.F6_HEAT:
            self._t = FluidQuantity(self, "t", w, h, 0.5, 0.5)
        self._bodies = bodies
        if bodies is not None:
            self.set_bodies(bodies)

    def sync_bodies(self):
        """Re-send the shared `bodies` list after main() advanced it (body.update(dt), F4:98)."""
        if self._bodies is not None:
            self.set_bodies(self._bodies)

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h_", None) and self._h_.value:
            self._fn["ifl_destroy"](self._h_)
            self._h_ = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's FluidSolver methods ---------------------------------------------
    def addInflow(self, x, y, w, h, d, *rest):
        """addInflow(x,y,w,h,d,u,v) F1:298 / addInflow(x,y,w,h,d,t,u,v) F6:1233"""
        if len(rest) == 2:
            t, (u, v) = 0.0, rest
        else:
            t, u, v = rest
        check(self._fn["ifl_add_inflow"](self._h_, x, y, w, h, d, t, u, v))

    def update(self, timestep, want_status=True):
        """FluidSolver.update(timestep) F1:276 ... F8:1306; returns what the reference only logs."""
        st = C.byref(self.status) if want_status else None
        check(self._fn["ifl_update"](self._h_, timestep, st), allow=(_abi.ERR_NAN,))
        return self.status.as_dict() if want_status else None

    def maxTimestep(self):
        """FluidSolver.maxTimestep() F1:309"""
        out = C.c_double()
        check(self._fn["ifl_max_timestep"](self._h_, C.byref(out)))
        return out.value

    def toImage(self, destination, renderHeat=False):
        """FluidSolver.toImage(destination) F1:341 ... F5:1133 / toImage(destination, renderHeat) F6:1250, F8:1417:
        the reference's plain-text PPM frame, byte for byte (pixels computed on the device)."""
        check(self._fn["ifl_write_ppm"](self._h_, os.fsencode(destination), 1 if renderHeat else 0))

    def image(self, renderHeat=False):
        """The integers of one frame as an array [h][(2w if renderHeat else w)][3] (ifl_to_image)."""
        n = C.c_size_t()
        check(self._fn["ifl_image_size"](self._h_, 1 if renderHeat else 0, C.byref(n)))
        out = np.empty(3 * n.value, dtype=np.int32)
        check(self._fn["ifl_to_image"](self._h_, 1 if renderHeat else 0, out.ctypes.data_as(C.POINTER(C.c_int32)), n.value))
        return out.reshape(self._h, -1, 3)

    def ambientT(self):
        """FluidSolver.ambientT() F6:1240"""
        return self.cfg.t_ambient

    # -- beyond the reference: batched stepping, stages, raw field access -------------------
    def set_bodies(self, bodies):
        arr = (_abi.Body * max(len(bodies), 1))(*[b.as_abi() if isinstance(b, SolidBody) else b for b in bodies])
        check(self._fn["ifl_set_bodies"](self._h_, arr, len(bodies)))

    def run_steps(self, n, timestep, rect, want_status=True):
        """n x { addInflow(*rect); update(timestep) } -- the body of the reference's main() loop."""
        st = C.byref(self.status) if want_status else None
        check(self._fn["ifl_run_steps"](self._h_, n, timestep, *rect, st), allow=(_abi.ERR_NAN,))
        return self.status.as_dict() if want_status else None

    def run_adaptive(self, t_end, rect, dt_max, cfl_fraction=1.0, max_steps=1000000):
        """Adaptive-timestep driver on top of maxTimestep() (F1:309-333, which the reference defines and never calls):
        until `t_end`, every step does addInflow(*rect), takes dt = min(dt_max, cfl_fraction * maxTimestep(), time left)
        from the state it is about to advance, and update(dt) -- the body of main() with a CFL-limited timestep.
        Returns the list of timesteps taken.  On a slab decomposition the same
        CFL reduction sizes the advection halo inside update()."""
        t, taken = 0.0, []
        while t < t_end and len(taken) < max_steps:
            self.addInflow(*rect)
            dt = min(dt_max, cfl_fraction * self.maxTimestep(), t_end - t)
            self.update(dt, want_status=False)
            taken.append(dt)
            t += dt
        return taken

    def run_stage(self, stage, timestep=0.0, want_status=False):
        st = C.byref(self.status) if want_status else None
        check(self._fn["ifl_run_stage"](self._h_, _abi.STAGES[stage], timestep, st), allow=(_abi.ERR_NAN,))
        return self.status.as_dict() if want_status else None

    def field_size(self, name):
        n = C.c_size_t()
        check(self._fn["ifl_field_size"](self._h_, _abi.FIELDS[name], C.byref(n)))
        return n.value

    def read_field(self, name, out=None):
        n = self.field_size(name)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        assert out.dtype == np.float64 and out.size == n and out.flags.c_contiguous
        check(self._fn["ifl_read_field"](self._h_, _abi.FIELDS[name], out.ctypes.data_as(_dp), n))
        return out

    def write_field(self, name, values):
        arr = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(self._fn["ifl_write_field"](self._h_, _abi.FIELDS[name], arr.ctypes.data_as(_dp), arr.size))

    def slab_rows(self, name):
        """(row0, nrows) of field `name` this rank owns in a slab decomposition (everything on a single GPU)."""
        r0, nr = C.c_int32(), C.c_int32()
        check(self._fn["ifl_slab_rows"](self._h_, _abi.FIELDS[name], C.byref(r0), C.byref(nr)))
        return r0.value, nr.value

    def read_field_rows(self, name, row0, nrows, out=None):
        width = self.field_size(name) // (self._h + (1 if name.startswith(("v", "phi")) else 0))
        if out is None:
            out = np.empty(nrows * width, dtype=np.float64)
        assert out.dtype == np.float64 and out.size == nrows * width and out.flags.c_contiguous
        check(self._fn["ifl_read_field_rows"](self._h_, _abi.FIELDS[name], row0, nrows, out.ctypes.data_as(_dp), out.size))
        return out

    def write_field_rows(self, name, row0, nrows, values):
        arr = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(self._fn["ifl_write_field_rows"](self._h_, _abi.FIELDS[name], row0, nrows, arr.ctypes.data_as(_dp), arr.size))

    # snake_case aliases (used by the parity tests)
    read = read_field
    write = write_field
    add_inflow = addInflow
    to_image = toImage
    max_timestep = maxTimestep

    # -- F8 particles (ParticleQuantities F8:1479) ------------------------------------------
    def particle_count(self):
        n = C.c_int32()
        check(self._fn["ifl_particle_count"](self._h_, C.byref(n)))
        return n.value

    def read_particles(self):
        n = self.particle_count()
        arrs = [np.empty(n, dtype=np.float64) for _ in range(6)]
        check(self._fn["ifl_read_particles"](self._h_, *[a.ctypes.data_as(_dp) for a in arrs], n))
        return dict(zip(("x", "y", "d", "t", "u", "v"), arrs))

    def write_particles(self, parts):
        arrs = [np.ascontiguousarray(parts[k], dtype=np.float64) for k in ("x", "y", "d", "t", "u", "v")]
        check(self._fn["ifl_write_particles"](self._h_, *[a.ctypes.data_as(_dp) for a in arrs], arrs[0].size))

    def synchronize(self):
        check(self._fn["ifl_synchronize"](self._h_))

    def launch_count(self):
        n = C.c_uint64()
        check(self._fn["ifl_launch_count"](self._h_, C.byref(n)))
        return n.value

    def set_profiling(self, enable=True):
        check(self._fn["ifl_set_profiling"](self._h_, 1 if enable else 0))

    def kernel_timing(self, kernel):
        """(average ms per launch, samples) of one solver kernel, from live CUDA-event samples."""
        ms, n = C.c_double(), C.c_int32()
        check(self._fn["ifl_kernel_timing"](self._h_, _abi.KERNELS[kernel], C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def wavefront_trace(self):
        """Per-strip timestamps of the two sweep kernels of a PCG iteration: array [2 sweeps][nstrips][16] (see the header)."""
        nstrips = (self._h + 31) // 32
        out = np.zeros(2 * nstrips * 16, dtype=np.int64)
        check(self._fn["ifl_wavefront_trace"](self._h_, out.ctypes.data_as(C.POINTER(C.c_int64)), out.size))
        return out.reshape(2, nstrips, 16)

    def last_timing(self):
        out = (C.c_double * 3)()
        check(self._fn["ifl_last_timing"](self._h_, out))
        return {"projection_ms": out[0], "advection_ms": out[1], "total_ms": out[2]}


def slab_partition(h, world_size):
    """Strip ranges of the slab decomposition, identical to ifl_create() (slab_range in csrc/ifl_internal.h): the
    grid's 32-row strips are dealt out in contiguous, balanced ranges -- base = nstrips // world, the first
    nstrips % world ranks own one strip more -- so no rank is ever empty.  Rank r owns strips [s0, s1), i.e. rows
    [32*s0, min(32*s1, h)).  Returns a list of (s0, s1, row0, row1) per rank."""
    nstrips = (h + 31) // 32
    if world_size < 1 or nstrips < world_size:
        raise ValueError("fewer 32-row strips than ranks")
    base, rem = divmod(nstrips, world_size)
    out = []
    for r in range(world_size):
        s0 = r * base + min(r, rem)
        s1 = s0 + base + (1 if r < rem else 0)
        out.append((s0, s1, min(32 * s0, h), min(32 * s1, h)))
    return out
