"""incremental-fluids on B200: the timestep hot path of g-amador/incremental-fluids-java as
hand-written sm_100a kernels behind a C ABI (include/incfluid_b200.h), plus this host-side
mirror of the reference's FluidSolver / FluidQuantity / SolidBody API.

The directory name contains '-', so import it with
    importlib.import_module("incremental-fluids-java_b200")
or through the repo-root shim `incfluid_b200`.
"""
from . import _abi as abi  # noqa: F401
from ._lib import build, load, IflError, LIB_PATH  # noqa: F401
from .solver import FluidSolver, FluidQuantity, SolidBody, SolidBox, SolidSphere, comm_unique_id, slab_partition  # noqa: F401

F1_MATRIXLESS = abi.F1_MATRIXLESS
F2_BETTER_ADVECTION = abi.F2_BETTER_ADVECTION
F3_CONJUGATE_GRADIENTS = abi.F3_CONJUGATE_GRADIENTS
F4_SOLID_BOUNDARIES = abi.F4_SOLID_BOUNDARIES
F5_CURVED_BOUNDARIES = abi.F5_CURVED_BOUNDARIES
F6_HEAT = abi.F6_HEAT
F7_VARIABLE_DENSITY = abi.F7_VARIABLE_DENSITY
F8_FLIP = abi.F8_FLIP
