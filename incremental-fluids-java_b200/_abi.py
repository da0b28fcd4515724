"""ctypes view of include/incfluid_b200.h (the C-ABI drop-in boundary).

Only plain structs, enums and prototypes live here; nothing in this module computes.
The same struct definitions are reused by the test harness so that both
sides of a parity test are driven with byte-identical configuration.
"""
import ctypes as C

ABI_VERSION = 1

# ifl_status_code
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NAN, ERR_COMM = range(6)

# ifl_variant (reference program reproduced)
F1_MATRIXLESS, F2_BETTER_ADVECTION, F3_CONJUGATE_GRADIENTS, F4_SOLID_BOUNDARIES, \
    F5_CURVED_BOUNDARIES, F6_HEAT, F7_VARIABLE_DENSITY, F8_FLIP = range(1, 9)

BODY_BOX, BODY_SPHERE = 0, 1

# ifl_flags
FLAG_SEQUENTIAL_REDUCTIONS = 1
FLAG_SEQUENTIAL_EXTRAPOLATE = 2
FLAG_SEPARATE_R_UPDATE = 4


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("variant", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
        ("density", C.c_double), ("density_soot", C.c_double), ("diffusion", C.c_double),
        ("t_ambient", C.c_double), ("gravity", C.c_double), ("flip_alpha", C.c_double),
        ("tolerance", C.c_double), ("solver_limit", C.c_int32),
        ("mic_tau", C.c_double), ("mic_sigma", C.c_double),
        ("particles_avg_per_cell", C.c_int32), ("particles_max_per_cell", C.c_int32),
        ("particles_min_per_cell", C.c_int32),
        ("rng_seed", C.c_uint64), ("device", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("flags", C.c_int32),
        ("comm_id", C.c_uint8 * 128),
    ]


class Body(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("reserved", C.c_int32),
        ("pos_x", C.c_double), ("pos_y", C.c_double),
        ("scale_x", C.c_double), ("scale_y", C.c_double),
        ("theta", C.c_double), ("cos_theta", C.c_double), ("sin_theta", C.c_double),
        ("vel_x", C.c_double), ("vel_y", C.c_double), ("vel_theta", C.c_double),
    ]


class StepStatus(C.Structure):
    _fields_ = [
        ("iterations_pressure", C.c_int32), ("iterations_heat", C.c_int32),
        ("residual_pressure", C.c_double), ("residual_heat", C.c_double),
        ("exceeded_budget", C.c_int32), ("nan_residual", C.c_int32),
        ("particle_count", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


FIELDS = {
    "d": 0, "t": 1, "u": 2, "v": 3, "d_dst": 4, "t_dst": 5, "u_dst": 6, "v_dst": 7,
    "p": 8, "r": 9, "z": 10, "s": 11, "precon": 12, "adiag": 13, "aplusx": 14, "aplusy": 15,
    "udensity": 16, "vdensity": 17,
    "volume_d": 20, "volume_t": 21, "volume_u": 22, "volume_v": 23,
    "normalx_d": 24, "normalx_t": 25, "normalx_u": 26, "normalx_v": 27,
    "normaly_d": 28, "normaly_t": 29, "normaly_u": 30, "normaly_v": 31,
    "cell_d": 32, "cell_t": 33, "cell_u": 34, "cell_v": 35,
    "body_d": 36, "body_t": 37, "body_u": 38, "body_v": 39,
    "phi_d": 40, "phi_t": 41, "phi_u": 42, "phi_v": 43,
}

STAGES = {
    "build_rhs": 1, "build_matrix": 2, "build_precon": 3, "project": 4, "apply_pressure": 5,
    "advect_d": 6, "advect_t": 7, "advect_u": 8, "advect_v": 9, "swap": 10,
    "apply_precon": 11, "matvec": 12, "fill_solid_fields": 13, "set_boundary": 14,
    "extrapolate_d": 15, "extrapolate_t": 16, "extrapolate_u": 17, "extrapolate_v": 18,
    "heat_solve": 19, "add_buoyancy": 20, "compute_densities": 21,
    "particles_to_grid": 22, "grid_to_particles": 23, "advect_particles": 24,
}

KERNELS = {"step_a": 0, "precon_fwd": 1, "precon_bwd": 2, "gs_sweep": 3}

_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes) for every symbol include/incfluid_b200.h declares
PROTOTYPES = {
    "ifl_default_config": (C.c_int, [C.POINTER(Config), C.c_int32, C.c_int32, C.c_int32]),
    "ifl_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "ifl_destroy": (None, [C.c_void_p]),
    "ifl_set_bodies": (C.c_int, [C.c_void_p, C.POINTER(Body), C.c_int32]),
    "ifl_add_inflow": (C.c_int, [C.c_void_p] + [C.c_double] * 8),
    "ifl_update": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(StepStatus)]),
    "ifl_run_steps": (C.c_int, [C.c_void_p, C.c_int32, C.c_double] + [C.c_double] * 8 + [C.POINTER(StepStatus)]),
    "ifl_run_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.POINTER(StepStatus)]),
    "ifl_max_timestep": (C.c_int, [C.c_void_p, _dp]),
    "ifl_field_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]),
    "ifl_read_field": (C.c_int, [C.c_void_p, C.c_int32, _dp, C.c_size_t]),
    "ifl_write_field": (C.c_int, [C.c_void_p, C.c_int32, _dp, C.c_size_t]),
    "ifl_read_field_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, C.c_size_t]),
    "ifl_write_field_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, C.c_size_t]),
    "ifl_slab_rows": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ifl_particle_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "ifl_read_particles": (C.c_int, [C.c_void_p] + [_dp] * 6 + [C.c_size_t]),
    "ifl_write_particles": (C.c_int, [C.c_void_p] + [_dp] * 6 + [C.c_size_t]),
    "ifl_image_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]),
    "ifl_to_image": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_size_t]),
    "ifl_write_ppm": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "ifl_synchronize": (C.c_int, [C.c_void_p]),
    "ifl_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ifl_last_timing": (C.c_int, [C.c_void_p, _dp]),
    "ifl_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "ifl_kernel_timing": (C.c_int, [C.c_void_p, C.c_int32, _dp, C.POINTER(C.c_int32)]),
    "ifl_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "ifl_wavefront_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_size_t]),
    "ifl_last_error": (C.c_char_p, []),
    "ifl_abi_version": (C.c_int, []),
}


def bind(lib, prototypes=PROTOTYPES, prefix_from="ifl_", prefix_to="ifl_"):
    """Attach restype/argtypes to `lib` for every prototype; raises AttributeError if a symbol is missing."""
    out = {}
    for name, (res, args) in prototypes.items():
        sym = prefix_to + name[len(prefix_from):]
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
        out[name] = fn
    return out


def default_config(variant, w, h, **overrides):
    """Python mirror of ifl_default_config(): the literals the reference hard-codes.

    F1-F3 run 600 solver iterations at most (F1:278, F3:365), F4-F8 2000 (F4:627 ...);
    density 0.1 everywhere, soot 1.0 (F6/F7 main) or 0.25 (F8:76); dt lives in main().
    """
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.variant = variant
    cfg.w, cfg.h = w, h
    cfg.density = 0.1
    cfg.density_soot = 0.25 if variant == F8_FLIP else 1.0
    cfg.diffusion = 0.01
    cfg.t_ambient = 294.0
    cfg.gravity = 9.81
    cfg.flip_alpha = 0.001
    cfg.tolerance = 1e-5
    cfg.solver_limit = 600 if variant <= F3_CONJUGATE_GRADIENTS else 2000
    cfg.mic_tau = 0.97
    cfg.mic_sigma = 0.25
    cfg.particles_avg_per_cell = 4
    cfg.particles_max_per_cell = 12
    cfg.particles_min_per_cell = 3
    cfg.rng_seed = 0x1234
    cfg.device = 0
    cfg.rank = 0
    cfg.world_size = 1
    cfg.flags = 0
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg
