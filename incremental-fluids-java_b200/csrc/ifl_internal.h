/*
 * ifl_internal.h -- host-side declarations shared by the translation units of
 * libincfluid_b200.so.  Nothing here crosses the C ABI (include/incfluid_b200.h).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/incfluid_b200.h"
#include "ifl_device.cuh"

namespace ifl {

/* Geometry of one FluidQuantity (F1:105-140): size, sample offset, cell size. */
struct QGeom {
    int w, h;
    double ox, oy;
};

/* Per-quantity solid data of F4+ (F4:138-153, F5:241-266), device pointers. */
struct SolidQ {
    int w, h;
    double ox, oy;
    double* phi;            /* (w+1)*(h+1) */
    double* volume;
    double* normalX;
    double* normalY;
    unsigned char* cell;    /* IFL_CELL_* */
    int* body;
    int* mask;
};
struct SolidView {
    const unsigned char* cell;
    const int* body;
    const double* volume;
};

/* Device scalars of one pressure solve.  Every solver kernel reads its coefficients
 * from here, so a whole solve is enqueued without returning to the host. */
struct SolveScalars {
    double sigma;        /* z.r */
    double alpha;        /* sigma / (z.s) */
    double neg_alpha;
    double beta;         /* sigmaNew / sigma */
    double zs;
    double max_error;    /* PCG: infinityNorm(r) (F3:579)  GS: maxDelta (F1:414) */
    unsigned long long max_bits;   /* running max, bit pattern of a non-negative float/double */
    int iterations;      /* iterations executed so far */
    int done;            /* 1: converged (or NaN) -- remaining kernels of the solve return at once */
    int nan;
    int exceeded;
    unsigned int ctr_a;  /* "last block" counters */
    unsigned int ctr_b;
    unsigned int ticket_f;  /* strip tickets of the forward / backward wavefront */
    unsigned int ticket_b;
    int p_pending;       /* 1: r has been updated by the last iteration but its p += alpha*s has not been applied yet */
    unsigned int ctr_c;  /* strips that finished the fused r update of the forward sweep */
    int zpar;            /* which generation of the z / s buffers is current (k_step_a flips it; never reset) */
};

/* slab decomposition (ifl_comm.cu): rank r owns the strips [s0, s1) of the pressure solve */
struct Comm {
    int rank = 0, world = 1;
    void* nccl = nullptr;             /* ncclComm_t */
    cudaStream_t stream = nullptr;
    void* peer_prev_base = nullptr;   /* IPC mapping of rank-1's / rank+1's inbox allocation (inbox_f | inbox_b) */
    void* peer_next_base = nullptr;
};
const char* comm_last_error();
int comm_unique_id(unsigned char out[128]);
int comm_init(Comm& c, int rank, int world, const unsigned char id_bytes[128], cudaStream_t st);
void comm_destroy(Comm& c);
int comm_map_neighbour_inboxes(Comm& c, void* inbox_base, void* scratch_dev);
int comm_all_gather(Comm& c, const void* send, void* recv, size_t bytes_per_rank);
/* every rank's own range of per-strip values -> every rank (in place; ranges follow slab_range) */
int comm_all_gather_strips(Comm& c, double* strip_vals, int nstrips, size_t per_strip);
/* balanced slab partition: the first nstrips % world ranks own one strip more; never empty when nstrips >= world */
static inline void slab_range(int nstrips, int world, int rank, int* s0, int* s1) {
    const int base = nstrips / world, rem = nstrips % world;
    *s0 = rank * base + (rank < rem ? rank : rem);
    *s1 = *s0 + base + (rank < rem ? 1 : 0);
}
int comm_exchange_rows(Comm& c, const double* send_up, double* recv_up, const double* send_down, double* recv_down, size_t n);
/* one grouped set of point-to-point transfers (halo rows of the row-major fields) */
struct CommXfer { int peer; double* ptr; size_t count; int send; };
int comm_transfers(Comm& c, const CommXfer* x, int n);
int comm_all_gather_u64(Comm& c, unsigned long long* vals /* [world], own value at [rank] */);

/* event-sampled kernel timing (ifl_set_profiling / ifl_kernel_timing) */
struct Profiler {
    static const int NS = 128;        /* samples kept */
    static const int NE = 6;          /* events per sample: before/after the 5 kernels of an iteration */
    int enabled = 0;
    int count = 0;                    /* samples recorded */
    cudaEvent_t ev[NS][NE];
    int kind[NS];                     /* 0 = PCG iteration, 1 = GS sweep */
    int* d_done = nullptr;            /* done flag seen by the sampled iteration */
    bool created = false;
};

struct SolveCtx {
    SkewGeom g;
    /* skewed N-vectors */
    double *r, *p, *adiag, *aplusx, *aplusy, *precon;
    double* p_first;         /* first owned strip of p */
    double *zb[2], *sb[2];   /* z and s, two generations each (k_step_a reads one and writes the other; SolveScalars.zpar) */
    unsigned* strip_ctr;     /* per-strip "last block" counters of the two-level reductions, [strips_alloc] */
    /* per-solve packed operands of the triangular sweeps (products aPlusX*precon, aPlusY*precon, precon) */
    double *pack_f, *pack_b;   /* [strip][tl/16][3][16][32], see k_precon_coeffs */
    /* strip boundary rows (flag-in-data) of the MIC(0) build and Gauss-Seidel wavefronts, [nstrips][w] each; two sets for GS parity */
    double *bnd_f, *bnd_b;
    /* global inbox rows of the triangular sweeps (strip hand-over between clusters / GPUs), [nstrips][inbox_len] each */
    double *inbox_f, *inbox_b;
    int inbox_len;
    double *partials;     /* per-block partial sums */
    int npartials;
    SolveScalars* sc;
    double tolerance;
    double mic_tau, mic_sigma;
    int sequential;       /* reductions in the reference's serial order (bit-exact mode) */
    int separate_r_update;   /* IFL_FLAG_SEPARATE_R_UPDATE: r -= alpha*z as a kernel of its own instead of inside the forward sweep */
    unsigned char* fl;    /* skewed fluid flags of _d._cell (F4+), nullptr for F1-F3 */
    int mask_vector_ops;  /* F6+: dotProduct/scaledAdd/infinityNorm only over fluid cells (F6:999-1064) */
    cudaStream_t stream;
    int const_matrix;     /* F3: the matrix is a function of the grid position alone (k_matvec_dot<true> re-derives it) */
    double const_scale;   /* ... with this `scale` (F3:436) */
    unsigned long long* launches;
    Profiler* prof;
    /* slab decomposition */
    Comm* comm;              /* nullptr on a single GPU */
    double* peer_inbox_f_next; /* rank+1's forward inbox rows (its inbox_f), IPC mapped */
    double* peer_inbox_b_prev; /* rank-1's backward inbox rows (its inbox_b) */
    double* rowbuf;          /* 8 packed rows of w doubles: z,s of my first / last row, z,s received from above / below */
    double* strip_part;      /* per-strip partial sums, [strips_alloc] */
    unsigned long long* rank_max;   /* per-rank max bits, [world] */
    int strips_alloc;        /* strips every vector is allocated for (the whole grid) */
    long long* trace;     /* diagnostic: per-strip timestamps of the next apply_precon, or nullptr */
};

/* ---- grid (row-major) kernels: ifl_grid_kernels.cu ---- */
void launch_advect(cudaStream_t st, int variant, QGeom q, const double* src, double* dst,
                   QGeom qu, const double* u, QGeom qv, const double* v, double timestep, double hx, int row0, int row1);
void launch_max_abs_rows(cudaStream_t st, const double* v, int w, int row0, int row1, unsigned long long* out_bits);
void launch_add_inflow(cudaStream_t st, int variant, QGeom q, double* src, double hx,
                       double x0, double y0, double x1, double y1, double value);
void launch_build_rhs(cudaStream_t st, const SkewGeom& g, const double* u, const double* v, double* r, double hx);
void launch_build_matrix_f3(cudaStream_t st, const SkewGeom& g, double* adiag, double* aplusx, double* aplusy, double scale);
void launch_apply_pressure(cudaStream_t st, const SkewGeom& g, const double* p, double* u, double* v, double scale);
/* row1 < 0: every row; otherwise rows [row0, row1) <-> a row-major buffer that starts at row0 */
void launch_skew_to_rows(cudaStream_t st, const SkewGeom& g, const double* skew, double* rows, int row0 = 0, int row1 = -1);
void launch_rows_to_skew(cudaStream_t st, const SkewGeom& g, const double* rows, double* skew, int row0 = 0, int row1 = -1);
void launch_max_velocity(cudaStream_t st, int w, int h, const double* u, const double* v, unsigned long long* out_bits);
void launch_to_image(cudaStream_t st, int variant, int w, int h, const double* d, const double* t, const double* volume,
                     const unsigned char* cell, double t_ambient, int render_heat, int* rgb);

/* ---- F4-F7 grid kernels: ifl_solid_kernels.cu ---- */
void launch_fill_solid_fields(cudaStream_t st, const SolidQ& q, const ifl_body* bodies, int nbodies, double hx, int variant);
void launch_cell_flags(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, unsigned char* fl);
void launch_set_boundary(cudaStream_t st, int w, int h, const unsigned char* cell, const int* body,
                         const ifl_body* bodies, double* u, double* v, double hx);
void launch_build_rhs_solid(cudaStream_t st, const SkewGeom& g, const SolidView& d, const SolidView& uq, const SolidView& vq,
                            const double* u, const double* v, const ifl_body* bodies, int nbodies, double* r, double hx, int variant);
void launch_compute_densities(cudaStream_t st, int w, int h, const double* dsrc, const double* tsrc, double* ud, double* vd,
                              double rho_air, double rho_soot, double t_amb);
void launch_build_matrix_solid(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, const double* uvol, const double* vvol,
                               const double* ud, const double* vd, double scale, int mode,
                               double* adiag, double* aplusx, double* aplusy);
void launch_apply_pressure_solid(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, const double* p, double* u, double* v,
                                 const double* ud, const double* vd, double scale, int use_density);
void launch_add_buoyancy(cudaStream_t st, int w, int h, const double* dsrc, const double* tsrc, double* v,
                         double timestep, double gravity, double rho_air, double rho_soot, double t_amb);
void launch_extrap_fill_mask(cudaStream_t st, const SolidQ& q, unsigned char* state);
void launch_extrap_sweep(cudaStream_t st, const SolidQ& q, double* src, unsigned char* state, unsigned int* ctl /* [4] */, int k);
void launch_advect_solid(cudaStream_t st, QGeom q, const double* src, double* dst, QGeom qu, const double* u, QGeom qv,
                         const double* v, const unsigned char* cell, const int* body, const ifl_body* bodies,
                         double timestep, double hx);
void launch_fill_value(cudaStream_t st, double* p, size_t n, double v);
void launch_fill_u8(cudaStream_t st, unsigned char* p, size_t n, unsigned char v);
void launch_u8_to_double(cudaStream_t st, const unsigned char* a, double* out, size_t n);
void launch_i32_to_double(cudaStream_t st, const int* a, double* out, size_t n);

/* ---- F8 FLIP: ifl_flip_kernels.cu ---- */
struct FlipCtx {
    int w = 0, h = 0;
    int count = 0;                    /* _particleCount (host mirror, updated at the few points it can change) */
    int max_particles = 0;            /* _maxParticles = w*h*_MaxPerCell (F8:1549) */
    unsigned long long rng_counter = 0;   /* draws consumed from the injected Math.random() stream */
    double *posX = nullptr, *posY = nullptr, *prop[4] = {nullptr, nullptr, nullptr, nullptr};
    int *counts = nullptr, *cell_start = nullptr, *cell_tmp = nullptr;
    int *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
    int *d_ints = nullptr, *h_ints = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int sequential_extrapolations = 0;
    bool ready = false;
};
cudaError_t flip_alloc(FlipCtx& f, int w, int h, int maxpc);
void flip_free(FlipCtx& f);
cudaError_t flip_init_particles(FlipCtx& f, cudaStream_t st, int avg, unsigned long long seed, const ifl_body* bodies, int nbodies,
                                double hx, unsigned long long* launches);
void flip_grid_to_particles(FlipCtx& f, cudaStream_t st, double alpha, QGeom qd, const double* d, QGeom qt, const double* t,
                            QGeom qu, const double* u, QGeom qv, const double* v, unsigned long long* launches);
void flip_advect_particles(FlipCtx& f, cudaStream_t st, double timestep, double hx, QGeom qu, const double* u, QGeom qv,
                           const double* v, const ifl_body* bodies, int nbodies, unsigned long long* launches);
cudaError_t flip_bin(FlipCtx& f, cudaStream_t st, QGeom q, unsigned long long* launches);
void flip_gather(FlipCtx& f, cudaStream_t st, QGeom q, const double* prop, double* src, unsigned char* cell, unsigned long long* launches);
void flip_diff(cudaStream_t st, double* src, const double* old, size_t n, double alpha, int sign, unsigned long long* launches);
cudaError_t flip_count_prune_seed(FlipCtx& f, cudaStream_t st, int maxpc, int minpc, unsigned long long seed,
                                  const ifl_body* bodies, int nbodies, double hx,
                                  QGeom qd, const double* d, QGeom qt, const double* t, QGeom qu, const double* u,
                                  QGeom qv, const double* v, unsigned long long* launches);
cudaError_t flip_extrapolate(FlipCtx& f, cudaStream_t st, const SolidQ& q, double* src, unsigned char* state,
                             unsigned int* d_progress, unsigned int* h_progress, int force_sequential, unsigned long long* launches);

/* ---- solver kernels: ifl_solve_kernels.cu ---- */
/* the int-returning helpers return 0, or -1 after a communication failure (comm_last_error()) */
int  solve_partials_needed(const SkewGeom& g);
int  solve_parity(SolveCtx& c);            /* current generation of z / s (synchronises the stream), -1 on error */
int  launch_build_precon(SolveCtx& c);
int  launch_apply_precon(SolveCtx& c, int as_in_loop = 0);     /* z = M^-1 r on the current generation */
int  pcg_enqueue_init(SolveCtx& c);
int  pcg_enqueue_iterations(SolveCtx& c, int n);
int  pcg_enqueue_finish(SolveCtx& c, int limit);
int  pcg_exchange_pressure_halo(SolveCtx& c);   /* the last row of p of rank-1 -> the row above my slab */
int  launch_matvec(SolveCtx& c);           /* z = A s on the current generation */
void gs_enqueue_begin(SolveCtx& c);
int  gs_enqueue_iterations(SolveCtx& c, int first_iter, int n, double scale, int limit);   /* iterations enqueued */
void gs_enqueue_finish(SolveCtx& c, int limit);

} // namespace ifl
