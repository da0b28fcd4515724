/*
 * incfluid_b200.h -- C ABI of the B200-native incremental-fluids timestep hot path.
 *
 * This header is the drop-in boundary.  The reference
 * (g-amador/incremental-fluids-java) has no FFI of its own: the hot path is
 * reached by direct Java calls into the nested classes FluidSolver /
 * FluidQuantity of each ExplicitFluid<N>*.java file.  Every entry point below
 * names the reference method(s) it replaces as file:line (paths relative to
 * src/physics/, F1..F8 as in SURVEY.md).  A Java FluidSolver keeps its public
 * surface and forwards to these symbols through Panama FFM or JNI; the stubs
 * are shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no C++/torch types, all functions return an ifl_status_code;
 *   - the library owns device memory, the caller owns every host buffer;
 *   - one handle = one CUDA device + one stream; a handle is NOT thread safe
 *     (the reference is single threaded; F8 even keeps its state static);
 *   - fields cross the boundary as dense row-major doubles, idx = x + y*width,
 *     exactly the layout of the Java double[] buffers (F1:149-155);
 *   - there is no CPU fallback: without a usable CUDA device every call that
 *     needs one returns IFL_ERR_CUDA.
 */
#ifndef INCFLUID_B200_H
#define INCFLUID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFL_ABI_VERSION 1

typedef enum ifl_status_code {
    IFL_OK = 0,
    IFL_ERR_INVALID = 1,      /* bad argument / bad handle */
    IFL_ERR_CUDA = 2,         /* CUDA runtime error, see ifl_last_error() */
    IFL_ERR_UNSUPPORTED = 3,  /* valid request this build does not implement */
    IFL_ERR_NAN = 4,          /* solver residual became NaN (reference: System.exit(-1), F3:616-619) */
    IFL_ERR_COMM = 5          /* NCCL error on the slab-decomposed path */
} ifl_status_code;

/* Which reference program the handle reproduces. */
typedef enum ifl_variant {
    IFL_F1_MATRIXLESS = 1,         /* EulerianFluids/ExplicitFluid1matrixless.java        */
    IFL_F2_BETTER_ADVECTION = 2,   /* EulerianFluids/ExplicitFluid2betterAdvection.java   */
    IFL_F3_CONJUGATE_GRADIENTS = 3,/* EulerianFluids/ExplicitFluid3conjugateGradients.java*/
    IFL_F4_SOLID_BOUNDARIES = 4,   /* EulerianFluids/ExplicitFluid4solidBoundaries.java   */
    IFL_F5_CURVED_BOUNDARIES = 5,  /* EulerianFluids/ExplicitFluid5curvedBoundaries.java  */
    IFL_F6_HEAT = 6,               /* EulerianFluids/ExplicitFluid6heat.java              */
    IFL_F7_VARIABLE_DENSITY = 7,   /* EulerianFluids/ExplicitFluid7variableDensity.java   */
    IFL_F8_FLIP = 8                /* hybridFluids/ExplicitFluid8flip.java                */
} ifl_variant;

/*
 * Solver configuration = the constructor arguments of FluidSolver plus the
 * literals the reference hard-codes.  ifl_default_config() fills the literals
 * with the reference's values; zero is never a magic "default" here.
 *   FluidSolver(w,h,density)                          F1:262, F2:305, F3:339
 *   FluidSolver(w,h,density,bodies)                   F4:594, F5:707
 *   FluidSolver(w,h,rhoAir,rhoSoot,diffusion,bodies)  F6:763, F7:769, F8:858
 */
typedef struct ifl_config {
    int32_t abi_version;      /* IFL_ABI_VERSION */
    int32_t variant;          /* ifl_variant */
    int32_t w, h;             /* grid size in cells */
    double  density;          /* F1-F5: fluid density; F6-F8: density of air (rhoAir) */
    double  density_soot;     /* F6-F8 rhoSoot */
    double  diffusion;        /* F6-F8 heat diffusion rate */
    double  t_ambient;        /* _tAmb = 294.0   (F6:772, F7:778, F8:867) */
    double  gravity;          /* _g    = 9.81    (F6:773, F7:779, F8:868) */
    double  flip_alpha;       /* _flipAlpha = 0.001 (F8:869) */
    double  tolerance;        /* 1e-5 (F1:420, F3:599,612) */
    int32_t solver_limit;     /* 600 (F1:278, F3:365) or 2000 (F4:627 ...) */
    double  mic_tau;          /* 0.97 (F3:465) */
    double  mic_sigma;        /* 0.25 (F3:466) */
    int32_t particles_avg_per_cell;  /* _AvgPerCell = 4  (F8:1494) */
    int32_t particles_max_per_cell;  /* _MaxPerCell = 12 (F8:1484) */
    int32_t particles_min_per_cell;  /* _MinPerCell = 3  (F8:1489) */
    /* F8: the unseeded Math.random() (F8:1576,1647) is replaced by an injected counter-based stream so that runs are
     * reproducible and a parallel implementation can consume "the k-th draw" in the reference's order:
     *   z = rng_seed + 0x9E3779B97F4A7C15*(k+1); z = (z^(z>>30))*0xBF58476D1CE4E5B9; z = (z^(z>>27))*0x94D049BB133111EB;
     *   z ^= z>>31;  draw_k = (float)(z >> 40) / 2^24            (splitmix64 finaliser, 24-bit float in [0,1))
     * Draws are consumed exactly where the reference calls random(): 2 per init / seed attempt, in loop order. */
    uint64_t rng_seed;
    int32_t device;           /* CUDA device ordinal */
    /* Slab decomposition.  world_size == 1: single GPU.
     * world_size > 1: this process (one process per rank: the strip hand-over maps the neighbours' memory with CUDA IPC)
     * owns a contiguous range of 32-row strips, see ifl_slab_rows(); comm_id holds the 128-byte ncclUniqueId all ranks
     * share. */
    int32_t rank;
    int32_t world_size;
    int32_t flags;            /* ifl_flags, 0 = default */
    uint8_t comm_id[128];
} ifl_config;

/* ifl_config.flags */
typedef enum ifl_flags {
    /* Evaluate dotProduct (F3:532) in the reference's strictly sequential order instead of
     * the fixed-shape parallel tree.  Slow (one thread per reduction); makes the whole PCG
     * solve bit-identical to the Java loops.  Meant for parity tests on small grids. */
    IFL_FLAG_SEQUENTIAL_REDUCTIONS = 1,
    /* F8 only: always replay FluidQuantity.extrapolate's LIFO stack (F8:651) with one device thread instead of the
     * level-parallel sweep.  The library switches to the replay by itself whenever two CELL_EMPTY cells touch (the
     * only case in which the reference's result depends on its stack order); the flag exists to test that path. */
    IFL_FLAG_SEQUENTIAL_EXTRAPOLATE = 2,
    /* PCG: run scaledAdd(_r, _r, _z, -alpha) + infinityNorm (F3:608-609) as a kernel of its own instead of inside the
     * forward triangular sweep (same operations, same bits; exists to test and time both arrangements). */
    IFL_FLAG_SEPARATE_R_UPDATE = 4
} ifl_flags;

/* Rigid body table entry = the fields of A_SolidBody (solidBodies/A_SolidBody.java:73-87).
 * cos_theta / sin_theta are cos(_theta), sin(_theta) evaluated ONCE by the
 * caller: the reference calls Math.cos/sin per distance evaluation
 * (A_SolidBody.java:67-68) and those are not bit-specified across JVM, glibc
 * and CUDA, so both sides of the boundary consume the same two doubles. */
typedef enum ifl_body_type { IFL_BODY_BOX = 0, IFL_BODY_SPHERE = 1 } ifl_body_type;
typedef struct ifl_body {
    int32_t type;             /* ifl_body_type: SolidBox.java:75 / SolidSphere.java:57 */
    int32_t reserved;
    double pos_x, pos_y;
    double scale_x, scale_y;  /* sphere: both = diameter s */
    double theta, cos_theta, sin_theta;
    double vel_x, vel_y, vel_theta;
} ifl_body;

/* What the reference only logs (F1:421,426; F3:613,628; F8:1771). */
typedef struct ifl_step_status {
    int32_t iterations_pressure;   /* iterations executed by the pressure solve */
    int32_t iterations_heat;       /* F6-F8 heat-diffusion solve */
    double  residual_pressure;     /* last maxDelta (GS) / maxError (PCG) */
    double  residual_heat;
    int32_t exceeded_budget;       /* bit0: pressure solve hit solver_limit, bit1: heat solve */
    int32_t nan_residual;          /* 1 if a residual was NaN */
    int32_t particle_count;        /* F8 */
    int32_t reserved;
} ifl_step_status;

/* Device-resident arrays a caller may read or write.  Sizes in doubles:
 *   cell-centred  w*h        : D, T, P, R, Z, S, PRECON, ADIAG, APLUSX, APLUSY, *_D/_T per-field arrays
 *   u faces       (w+1)*h    : U, UDENSITY, *_U
 *   v faces       w*(h+1)    : V, VDENSITY, *_V
 * "SRC" is FluidQuantity._src, "DST" is _dst (F8: _old). */
typedef enum ifl_field {
    IFL_FIELD_D = 0, IFL_FIELD_T = 1, IFL_FIELD_U = 2, IFL_FIELD_V = 3,           /* _src */
    IFL_FIELD_D_DST = 4, IFL_FIELD_T_DST = 5, IFL_FIELD_U_DST = 6, IFL_FIELD_V_DST = 7,
    IFL_FIELD_P = 8, IFL_FIELD_R = 9, IFL_FIELD_Z = 10, IFL_FIELD_S = 11,
    IFL_FIELD_PRECON = 12, IFL_FIELD_ADIAG = 13, IFL_FIELD_APLUSX = 14, IFL_FIELD_APLUSY = 15,
    IFL_FIELD_UDENSITY = 16, IFL_FIELD_VDENSITY = 17,
    /* per-quantity solid data (F4:138-153, F5:241-266); q = D,T,U,V in that order */
    IFL_FIELD_VOLUME_D = 20, IFL_FIELD_VOLUME_T = 21, IFL_FIELD_VOLUME_U = 22, IFL_FIELD_VOLUME_V = 23,
    IFL_FIELD_NORMALX_D = 24, IFL_FIELD_NORMALX_T = 25, IFL_FIELD_NORMALX_U = 26, IFL_FIELD_NORMALX_V = 27,
    IFL_FIELD_NORMALY_D = 28, IFL_FIELD_NORMALY_T = 29, IFL_FIELD_NORMALY_U = 30, IFL_FIELD_NORMALY_V = 31,
    /* cell type / closest body, returned widened to double (0 = FLUID, 1 = SOLID, 2 = EMPTY) */
    IFL_FIELD_CELL_D = 32, IFL_FIELD_CELL_T = 33, IFL_FIELD_CELL_U = 34, IFL_FIELD_CELL_V = 35,
    IFL_FIELD_BODY_D = 36, IFL_FIELD_BODY_T = 37, IFL_FIELD_BODY_U = 38, IFL_FIELD_BODY_V = 39,
    /* phi is (fw+1)*(fh+1) for a quantity of size fw*fh (F5:296) */
    IFL_FIELD_PHI_D = 40, IFL_FIELD_PHI_T = 41, IFL_FIELD_PHI_U = 42, IFL_FIELD_PHI_V = 43
} ifl_field;

/* Single stages of FluidSolver.update(), exposed so that every kernel can be
 * parity-tested in isolation against the oracle (rows a7, a10-a24 of SURVEY 8a).
 * They operate on the handle's device-resident state. */
typedef enum ifl_stage {
    IFL_STAGE_BUILD_RHS = 1,          /* buildRhs            F1:364 F3:420 F4:693 F5:785 F7:813 */
    IFL_STAGE_BUILD_MATRIX = 2,       /* buildPressureMatrix F3:435 F4:713 F5:825 F7:877 */
    IFL_STAGE_BUILD_PRECON = 3,       /* buildPreconditioner F3:464 F4:744 F7:951 */
    IFL_STAGE_PROJECT = 4,            /* project             F1:380 (GS) / F3:590 F7:1102 (PCG) */
    IFL_STAGE_APPLY_PRESSURE = 5,     /* applyPressure       F1:432 F3:634 F4:927 F7:1148 */
    IFL_STAGE_ADVECT_D = 6,           /* _d.advect (writes _dst, no swap)  F1:178 F3:207 F4:268 */
    IFL_STAGE_ADVECT_T = 7,
    IFL_STAGE_ADVECT_U = 8,
    IFL_STAGE_ADVECT_V = 9,
    IFL_STAGE_SWAP = 10,              /* swap() of every advected quantity F1:143 */
    IFL_STAGE_APPLY_PRECON = 11,      /* z = M^-1 r  applyPreconditioner(_z,_r) F3:495 */
    IFL_STAGE_MATVEC = 12,            /* z = A s     matrixVectorProduct(_z,_s) F3:544 */
    IFL_STAGE_FILL_SOLID_FIELDS = 13, /* fillSolidFields on d,(t,)u,v F4:329 F5:463 F7:484 */
    IFL_STAGE_SET_BOUNDARY = 14,      /* setBoundaryCondition F4:948 F7:1185 */
    IFL_STAGE_EXTRAPOLATE_D = 15,     /* FluidQuantity.extrapolate F4:452 F7:632 */
    IFL_STAGE_EXTRAPOLATE_T = 16,
    IFL_STAGE_EXTRAPOLATE_U = 17,
    IFL_STAGE_EXTRAPOLATE_V = 18,
    IFL_STAGE_HEAT_SOLVE = 19,        /* r<-T; heat matrix; MIC; project; T<-p  F6:1190-1201 F7:1221-1230 */
    IFL_STAGE_ADD_BUOYANCY = 20,      /* addBuoyancy F6:1135 F7:1169 */
    IFL_STAGE_COMPUTE_DENSITIES = 21, /* computeDensities F7:854 */
    IFL_STAGE_PARTICLES_TO_GRID = 22, /* ParticleQuantities.particlesToGrid F8:1761 */
    IFL_STAGE_GRID_TO_PARTICLES = 23, /* diff; gridToParticles(alpha); undiff F8:1367-1384 */
    IFL_STAGE_ADVECT_PARTICLES = 24   /* ParticleQuantities.advect F8:1778 */
} ifl_stage;

typedef struct ifl_solver ifl_solver;   /* opaque */

/* Fill *cfg with the reference literals for `variant` at grid w x h. */
int ifl_default_config(ifl_config* cfg, int32_t variant, int32_t w, int32_t h);

/* Slab decomposition (the reference has no distributed path).  Rank 0 obtains a 128-byte id
 * (an ncclUniqueId), the host program broadcasts it to the other ranks by any means, and every rank passes it in
 * ifl_config.comm_id together with its rank / world_size.  One process per GPU.  Every rank computes only the rows it
 * owns (ifl_slab_rows): RHS, PCG, applyPressure and the semi-Lagrangian advection, whose halo of
 * ceil(dt * max|v| / hx) + 4 rows is sized every step from a CFL reduction over all ranks (cf. maxTimestep F1:309) and
 * exchanged with ncclSend/ncclRecv.  That is variant 3.
 * Variants 4-7 (solid bodies, heat, variable density) on several GPUs: every rank holds the whole state and repeats
 * every stage of update() except the solver loops (heat and pressure MIC(0)-PCG, >= 98 % of such a step), which run on
 * the rank's strips exactly like variant 3's; the strips of p are gathered on every rank after a solve.  Whole-field
 * reads and writes stay valid on every rank (a host that changes a field writes it on every rank); z and s hold only
 * the rank's strips.  Variants 1, 2 (Gauss-Seidel) and 8 (particles) are single-GPU. */
int ifl_comm_unique_id(uint8_t out[128]);

/* new FluidSolver(...)  -- F1:262 F3:339 F4:594 F6:763 F7:769 F8:858 */
int ifl_create(const ifl_config* cfg, ifl_solver** out);
void ifl_destroy(ifl_solver* s);

/* The `bodies` ArrayList the solver shares with main() (F4:594, F7:767).  The
 * reference mutates it between frames (body.update(dt), F4:98); the caller
 * re-sends the table after doing the same. */
int ifl_set_bodies(ifl_solver* s, const ifl_body* bodies, int32_t count);

/* FluidSolver.addInflow(x,y,w,h,d,u,v) F1:298 F3:385 / (x,y,w,h,d,t,u,v) F6:1233 F7:1265.
 * `t` is ignored by variants without a temperature field.  (F8 calls its own
 * addInflow inside update(), F8:1330; for F8 this sets that rectangle.) */
int ifl_add_inflow(ifl_solver* s, double x, double y, double w, double h,
                   double d, double t, double u, double v);

/* FluidSolver.update(timestep) F1:276 F2:322 F3:361 F4:617 F5:1089 F6:1181 F7:1212 F8:1306.
 * `status` may be NULL; passing it costs one device synchronisation. */
int ifl_update(ifl_solver* s, double timestep, ifl_step_status* status);

/* Run n_steps x { addInflow(rect...) ; update(timestep) } without returning to
 * the host in between -- the body of the reference's main() loop (F3:75-79). */
int ifl_run_steps(ifl_solver* s, int32_t n_steps, double timestep,
                  double x, double y, double w, double h,
                  double d, double t, double u, double v,
                  ifl_step_status* last_status);

/* One stage of update() on the resident state (see ifl_stage). */
int ifl_run_stage(ifl_solver* s, int32_t stage, double timestep, ifl_step_status* status);

/* FluidSolver.maxTimestep() F1:309 */
int ifl_max_timestep(ifl_solver* s, double* out);

/* Element count of a field (in doubles) for this handle. */
int ifl_field_size(const ifl_solver* s, int32_t field, size_t* count);
/* Copy a field device -> host / host -> device.  `count` must equal ifl_field_size. */
int ifl_read_field(ifl_solver* s, int32_t field, double* host, size_t count);
int ifl_write_field(ifl_solver* s, int32_t field, const double* host, size_t count);
/* Rows [row0, row0 + nrows) of a field (count = nrows * row length).  In the slab decomposition of variant 3 a rank holds
 * -- and a host reads and writes -- only the rows it owns, ifl_slab_rows(); the whole-field calls above return
 * IFL_ERR_UNSUPPORTED there.  (Variants 4-7 on several GPUs keep whole fields on every rank: both forms work.) */
int ifl_read_field_rows(ifl_solver* s, int32_t field, int32_t row0, int32_t nrows, double* host, size_t count);
int ifl_write_field_rows(ifl_solver* s, int32_t field, int32_t row0, int32_t nrows, const double* host, size_t count);
/* The rows of `field` this rank owns: the rows of its 32-row strips (balanced contiguous ranges: nstrips / world strips
 * each, the first nstrips % world ranks one more); the extra last row of v / phi belongs to the last rank. */
int ifl_slab_rows(const ifl_solver* s, int32_t field, int32_t* row0, int32_t* nrows);

/* F8 particles: positions and the four per-particle properties (d,t,u,v order, F8:883-886). */
int ifl_particle_count(ifl_solver* s, int32_t* count);
int ifl_read_particles(ifl_solver* s, double* pos_x, double* pos_y,
                       double* prop_d, double* prop_t, double* prop_u, double* prop_v,
                       size_t capacity);
int ifl_write_particles(ifl_solver* s, const double* pos_x, const double* pos_y,
                        const double* prop_d, const double* prop_t,
                        const double* prop_u, const double* prop_v, size_t count);

/* FluidSolver.toImage: the reference's only observable output (a plain-text PPM frame per call from main()).
 *   F1:341 (=F2:356, F3:397): shade = clamp((int)((1 - d) * 255), 0, 255);  F4:664: solid cells black;
 *   F5:1133: (1 - d) * volume * 255;  toImage(destination, renderHeat) F6:1250 (=F7:1282): the density panel
 *   (int)(clamp((1 - d) * volume, 0, 1) * 255) and, with renderHeat, a heat panel to its LEFT (t = (T - Tamb) / 700);
 *   F8:1417: EMPTY cells red, t = |(float) T - Tamb| / 70.
 * ifl_image_size: pixels of a frame = (render_heat ? 2w : w) * h (render_heat needs variant >= F6).
 * ifl_to_image:   the three integers per pixel the reference writes, rgb[3 * pixels], computed on the device.
 * ifl_write_ppm:  the reference's text file, byte for byte: "P3\n<W> <h>\n255\n", then "r g b " per pixel with a
 *                 newline after every pixel whose index is a multiple of the SOLVER width (the `i % _w == 0` of F1:354,
 *                 kept as is -- also for the double-width frames of renderHeat). */
int ifl_image_size(const ifl_solver* s, int32_t render_heat, size_t* pixels);
int ifl_to_image(ifl_solver* s, int32_t render_heat, int32_t* rgb, size_t pixels);
int ifl_write_ppm(ifl_solver* s, const char* path, int32_t render_heat);

/* Block until every launched kernel of this handle has finished. */
int ifl_synchronize(ifl_solver* s);

/* Count of kernels this handle has launched since creation (bench.py gpu_launches). */
int ifl_launch_count(const ifl_solver* s, uint64_t* count);

/* Device-side duration of the most recent ifl_update / ifl_run_steps per phase, in ms
 * (CUDA events on the handle's stream): [0] projection, [1] advection, [2] whole call. */
int ifl_last_timing(ifl_solver* s, double out_ms[3]);

/* Live kernel timing for the roofline report (bench.py).  With profiling enabled the solver
 * brackets the kernels of ONE iteration per enqueued chunk with CUDA events on the handle's
 * stream (a sample; iterations that ran after convergence are discarded).  ifl_kernel_timing
 * returns the average device time per launch of kernel `which` over the samples collected
 * since the last ifl_set_profiling(s, 1). */
typedef enum ifl_kernel {
    /* the three kernels of one MIC(0)-PCG iteration (F3:603-628) */
    IFL_KERNEL_STEP_A = 0,        /* scaledAdd(p) F3:607 + scaledAdd(s,z,s,beta) F3:626 of the previous iteration,
                                   * matrixVectorProduct F3:544 + dotProduct F3:532 */
    IFL_KERNEL_PRECON_FWD = 1,    /* scaledAdd(r) F3:608 + infinityNorm F3:579 + applyPreconditioner forward sweep F3:496-507 */
    IFL_KERNEL_PRECON_BWD = 2,    /* applyPreconditioner backward sweep F3:509-521 + dotProduct F3:625 */
    IFL_KERNEL_GS_SWEEP = 3,      /* one Gauss-Seidel sweep                  F1:383-417 */
    IFL_KERNEL_COUNT = 4
} ifl_kernel;
int ifl_set_profiling(ifl_solver* s, int32_t enable);
int ifl_kernel_timing(ifl_solver* s, int32_t which, double* avg_ms, int32_t* samples);

/* Diagnostic for profiles/: per-strip timestamps of the two sweep kernels of one PCG iteration (forward sweep with the
 * fused r update, backward sweep with the fused dot product) on the resident state; r and z are modified.
 * out[(sweep*nstrips + k)*16 + i], sweep 0 = forward, 1 = backward; i: 0 enter, 1 after the first block, 3 end (device
 * ns), 6 ticket slot, 7 SM id; builds with -DIFL_SW_TRACE_WAITS also fill 2, 4, 5, 8 (strip warp: starved steps, time of
 * blocks 1-3, clocks waiting for credit / edge / operand tiles) and 9-13 (loader warp: clocks waiting for bulk copies,
 * updating r, issuing copies, total, tiles). */
int ifl_wavefront_trace(ifl_solver* s, int64_t* out, size_t capacity);

/* Human readable text for the last error on this thread. */
const char* ifl_last_error(void);
int ifl_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* INCFLUID_B200_H */
