/*
 * ifl_api.cu -- the C ABI of include/incfluid_b200.h: handle lifetime, device-resident
 * state and the orchestration of FluidSolver.update() as a
 * stream of kernels.  There is no CPU fallback: every entry point that needs the device
 * fails with IFL_ERR_CUDA when CUDA is unusable.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include <new>
#include "ifl_internal.h"

using namespace ifl;

static thread_local char g_err[512] = "";
static void set_err(const char* fmt, const char* a, const char* b, int line) {
    snprintf(g_err, sizeof g_err, fmt, a, b, line);
}
#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_err("CUDA error: %s at %s (ifl_api.cu:%d)", cudaGetErrorString(e_), #call, __LINE__); \
            return IFL_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)
#define COMM(call)                                                                        \
    do {                                                                                  \
        if ((call) != 0) { snprintf(g_err, sizeof g_err, "%s", comm_last_error()); return IFL_ERR_COMM; } \
    } while (0)
#define INVALID(msg)                                                                      \
    do { snprintf(g_err, sizeof g_err, "invalid argument: %s", msg); return IFL_ERR_INVALID; } while (0)

struct Quantity {             /* FluidQuantity: _src/_dst ping-pong (F1:105-147) */
    QGeom q{};
    double* buf[2] = {nullptr, nullptr};
    int cur = 0;
    SolidQ sq{};              /* F4+: _phi/_volume/_normalX/_normalY/_cell/_body/_mask (F4:138-153, F5:241-266) */
    bool present = false;
    SolidView view() const { return SolidView{sq.cell, sq.body, sq.volume}; }
    double* src() const { return buf[cur]; }
    double* dst() const { return buf[cur ^ 1]; }
    void swap() { cur ^= 1; }                       /* F1:143 */
    size_t count() const { return (size_t)q.w * q.h; }
};

struct ifl_solver {
    ifl_config cfg;
    int w, h;
    double hx;
    cudaStream_t stream = nullptr;
    Quantity d, t, u, v;
    SolveCtx sx{};
    double* skew_block = nullptr;      /* one allocation holding the skewed N-vectors */
    double* pack_alloc = nullptr;      /* packed sweep coefficients of this rank's strips (forward | backward) */
    size_t slab_bytes = 0;             /* bytes of the solver vectors + packed coefficients on this rank */
    double* scratch = nullptr;         /* (w+1)*(h+1) doubles for layout conversions */
    SolveScalars* h_sc = nullptr;      /* pinned mirrors for convergence polling (2 in flight) */
    unsigned long long* d_maxbits = nullptr;
    cudaEvent_t ev_chunk[2] = {nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   /* start, after projection, after advection */
    bool ev_valid = false;
    unsigned long long launches = 0;
    ifl_step_status last{};
    int chunk_pcg = 32, chunk_gs = 64;
    Profiler prof;
    /* F4+ */
    ifl_body* d_bodies = nullptr;
    int nbodies = 0, bodies_cap = 0;
    double *udens = nullptr, *vdens = nullptr;      /* _uDensity/_vDensity F7:757-758 */
    unsigned char* extrap_state = nullptr;
    unsigned int* d_progress = nullptr;
    unsigned int* h_progress = nullptr;
    unsigned int* d_extrap_ctl = nullptr;          /* [4]: three rotating progress counters + the finished flag (k_extrap_sweep) */
    unsigned int* h_extrap_flag = nullptr;         /* pinned, [2] */
    SolveScalars* sc_heat = nullptr;                /* device copy of the heat solve's final scalars */
    SolveScalars* h_sc_heat = nullptr;
    /* F8 */
    FlipCtx flip;
    Comm comm;
    double inflow[8] = {0.45, 0.2, 0.2, 0.05, 1.0, 294.0, 0.0, 0.0};   /* F8:1330 */
    int last_halo = 0;                 /* slab decomposition: halo rows of the last advection */
};

static bool uses_pcg(const ifl_solver* s) { return s->cfg.variant >= IFL_F3_CONJUGATE_GRADIENTS; }
static bool has_solids(const ifl_solver* s) { return s->cfg.variant >= IFL_F4_SOLID_BOUNDARIES; }
static bool has_heat(const ifl_solver* s) { return s->cfg.variant >= IFL_F6_HEAT; }
static bool has_vardens(const ifl_solver* s) { return s->cfg.variant >= IFL_F7_VARIABLE_DENSITY; }
static bool is_flip(const ifl_solver* s) { return s->cfg.variant == IFL_F8_FLIP; }
/* several GPUs, F4-F7: every rank holds the whole state and repeats every stage except the solver loop, which runs on the
 * rank's strips (the MIC(0)-PCG iterations are >= 98 % of such a step); F3 instead partitions the fields themselves */
static bool replicated_stages(const ifl_solver* s) { return s->cfg.world_size > 1 && has_solids(s) && !is_flip(s); }

extern "C" {

int ifl_abi_version(void) { return IFL_ABI_VERSION; }

int ifl_comm_unique_id(uint8_t out[128]) {
    if (!out) INVALID("NULL argument");
    if (comm_unique_id(out)) { snprintf(g_err, sizeof g_err, "%s", comm_last_error()); return IFL_ERR_COMM; }
    return IFL_OK;
}
const char* ifl_last_error(void) { return g_err; }

int ifl_default_config(ifl_config* cfg, int32_t variant, int32_t w, int32_t h) {
    if (!cfg) INVALID("cfg is NULL");
    if (variant < IFL_F1_MATRIXLESS || variant > IFL_F8_FLIP) INVALID("unknown variant");
    memset(cfg, 0, sizeof *cfg);
    cfg->abi_version = IFL_ABI_VERSION;
    cfg->variant = variant;
    cfg->w = w; cfg->h = h;
    cfg->density = 0.1;                                         /* F1:62, F7:75 */
    cfg->density_soot = (variant == IFL_F8_FLIP) ? 0.25 : 1.0;  /* F8:76 / F6:77 */
    cfg->diffusion = 0.01;
    cfg->t_ambient = 294.0;
    cfg->gravity = 9.81;
    cfg->flip_alpha = 0.001;
    cfg->tolerance = 1e-5;
    cfg->solver_limit = variant <= IFL_F3_CONJUGATE_GRADIENTS ? 600 : 2000;
    cfg->mic_tau = 0.97;
    cfg->mic_sigma = 0.25;
    cfg->particles_avg_per_cell = 4;
    cfg->particles_max_per_cell = 12;
    cfg->particles_min_per_cell = 3;
    cfg->rng_seed = 0x1234;
    cfg->device = 0;
    cfg->rank = 0;
    cfg->world_size = 1;
    cfg->flags = 0;
    return IFL_OK;
}

static int alloc_quantity(Quantity& q, int w, int h, double ox, double oy, int variant = 0, cudaStream_t st = nullptr) {
    q.q = QGeom{w, h, ox, oy};
    q.present = true;
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc(&q.buf[i], sizeof(double) * q.count()));
        CU(cudaMemset(q.buf[i], 0, sizeof(double) * q.count()));
    }
    q.cur = 0;
    if (variant >= IFL_F4_SOLID_BOUNDARIES) {
        SolidQ& s = q.sq;
        s.w = w; s.h = h; s.ox = ox; s.oy = oy;
        const size_t n = q.count(), np = (size_t)(w + 1) * (h + 1);
        CU(cudaMalloc(&s.phi, sizeof(double) * np));       CU(cudaMemset(s.phi, 0, sizeof(double) * np));
        CU(cudaMalloc(&s.volume, sizeof(double) * n));
        CU(cudaMalloc(&s.normalX, sizeof(double) * n));    CU(cudaMemset(s.normalX, 0, sizeof(double) * n));
        CU(cudaMalloc(&s.normalY, sizeof(double) * n));    CU(cudaMemset(s.normalY, 0, sizeof(double) * n));
        CU(cudaMalloc(&s.cell, n));
        CU(cudaMalloc(&s.body, sizeof(int) * n));          CU(cudaMemset(s.body, 0, sizeof(int) * n));
        CU(cudaMalloc(&s.mask, sizeof(int) * n));          CU(cudaMemset(s.mask, 0, sizeof(int) * n));
        /* F5:313-316: _cell = FLUID, _volume = 1.0; F4:183 leaves _cell null (3 = neither fluid nor solid) */
        launch_fill_value(st, s.volume, n, 1.0);
        launch_fill_u8(st, s.cell, n, variant == IFL_F4_SOLID_BOUNDARIES ? 3 : IFL_CELL_FLUID);
    }
    return IFL_OK;
}

void ifl_destroy(ifl_solver* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (Quantity* q : {&s->d, &s->t, &s->u, &s->v}) {
        for (int i = 0; i < 2; i++) if (q->buf[i]) cudaFree(q->buf[i]);
        void* ptrs[] = {q->sq.phi, q->sq.volume, q->sq.normalX, q->sq.normalY, q->sq.cell, q->sq.body, q->sq.mask};
        for (void* p : ptrs) if (p) cudaFree(p);
    }
    if (s->h_extrap_flag) cudaFreeHost(s->h_extrap_flag);
    void* extra[] = {s->d_bodies, s->udens, s->vdens, s->extrap_state, s->d_progress, s->d_extrap_ctl, s->sc_heat,
                     s->sx.fl ? (void*)(s->sx.fl - s->sx.g.ss) : nullptr};
    for (void* p : extra) if (p) cudaFree(p);
    if (s->h_progress) cudaFreeHost(s->h_progress);
    if (s->h_sc_heat) cudaFreeHost(s->h_sc_heat);
    flip_free(s->flip);
    if (s->skew_block) cudaFree(s->skew_block);
    if (s->scratch) cudaFree(s->scratch);
    if (s->pack_alloc) cudaFree(s->pack_alloc);
    comm_destroy(s->comm);
    if (s->sx.bnd_f) cudaFree(s->sx.bnd_f);
    if (s->sx.inbox_f) cudaFree(s->sx.inbox_f);
    if (s->sx.strip_part) cudaFree(s->sx.strip_part);
    if (s->sx.strip_ctr) cudaFree(s->sx.strip_ctr);
    if (s->sx.rank_max) cudaFree(s->sx.rank_max);
    if (s->sx.rowbuf) cudaFree(s->sx.rowbuf);
    if (s->sx.partials) cudaFree(s->sx.partials);
    if (s->sx.sc) cudaFree(s->sx.sc);
    if (s->d_maxbits) cudaFree(s->d_maxbits);
    if (s->h_sc) cudaFreeHost(s->h_sc);
    if (s->prof.created) {
        for (int i = 0; i < Profiler::NS; i++) for (int j = 0; j < Profiler::NE; j++) cudaEventDestroy(s->prof.ev[i][j]);
        cudaFree(s->prof.d_done);
    }
    for (auto& e : s->ev_chunk) if (e) cudaEventDestroy(e);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static int create_impl(const ifl_config* cfg, ifl_solver* s) {
    s->cfg = *cfg;
    s->w = cfg->w; s->h = cfg->h;
    s->hx = 1.0 / (cfg->w < cfg->h ? cfg->w : cfg->h);          /* F1:266 */
    CU(cudaSetDevice(cfg->device));
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    int rc;
    const int var = cfg->variant;
    if ((rc = alloc_quantity(s->d, s->w, s->h, 0.5, 0.5, var, s->stream))) return rc;       /* F1:268 */
    if ((rc = alloc_quantity(s->u, s->w + 1, s->h, 0.0, 0.5, var, s->stream))) return rc;   /* F1:269 */
    if ((rc = alloc_quantity(s->v, s->w, s->h + 1, 0.5, 0.0, var, s->stream))) return rc;   /* F1:270 */
    if (has_heat(s)) {
        if ((rc = alloc_quantity(s->t, s->w, s->h, 0.5, 0.5, var, s->stream))) return rc;   /* F6:792 */
        launch_fill_value(s->stream, s->t.src(), s->t.count(), cfg->t_ambient);             /* F6:797-799 */
    }
    if (has_vardens(s)) {
        CU(cudaMalloc(&s->udens, sizeof(double) * s->u.count())); CU(cudaMemset(s->udens, 0, sizeof(double) * s->u.count()));
        CU(cudaMalloc(&s->vdens, sizeof(double) * s->v.count())); CU(cudaMemset(s->vdens, 0, sizeof(double) * s->v.count()));
    }
    if (is_flip(s)) {
        CU(flip_alloc(s->flip, s->w, s->h, cfg->particles_max_per_cell));
        s->inflow[5] = cfg->t_ambient;
    }
    if (has_solids(s)) {
        CU(cudaMalloc(&s->extrap_state, (size_t)(s->w + 1) * (s->h + 1)));
        CU(cudaMalloc(&s->d_progress, sizeof(unsigned int)));
        CU(cudaMalloc(&s->d_extrap_ctl, 4 * sizeof(unsigned int)));
        CU(cudaMallocHost(&s->h_extrap_flag, 2 * sizeof(unsigned int)));
        CU(cudaMallocHost(&s->h_progress, sizeof(unsigned int)));
        CU(cudaMalloc(&s->sc_heat, sizeof(SolveScalars))); CU(cudaMemset(s->sc_heat, 0, sizeof(SolveScalars)));
        CU(cudaMallocHost(&s->h_sc_heat, sizeof(SolveScalars))); memset(s->h_sc_heat, 0, sizeof(SolveScalars));
    }

    SkewGeom& g = s->sx.g;
    g.w = s->w; g.h = s->h;
    g.nstrips = (s->h + 31) / 32;
    g.tl = (((s->w + IFL_SKB - 1) / IFL_SKB + 31 + 7) / 8) * 8 + 8;      /* m-rows: ceil(w/SKB)+31, padded to the staging block */
    g.ss = (long long)g.tl * 32 * IFL_SKB;
    /* slab decomposition: balanced contiguous ranges of strips (slab_range): no rank is ever empty */
    const int world = cfg->world_size;
    const int strips_alloc = g.nstrips;
    slab_range(g.nstrips, world, cfg->rank, &g.s0, &g.s1);
    g.total = (long long)strips_alloc * g.ss;
    /* one all-zero strip before and after each vector: the wavefront kernels read the
     * "row above the first" / "row below the last" from there instead of branching.
     * Slab decomposition: the vectors only the solver loop touches (r p z s, both generations) and the packed sweep
     * coefficients hold this rank's strips plus one halo strip on either side; they are addressed with the GLOBAL strip
     * index through a base pointer shifted by -s0 strips.  The matrix and the MIC(0) factor stay whole-grid: the factor
     * is a recurrence from strip 0 that every rank repeats for itself (its inputs depend on the grid position only).
     * Variants with solid bodies (F4-F7) on several GPUs keep every field and vector whole on every rank: all stages but
     * the solver loop are repeated by every rank (replicated_stages()), so r and p must be whole-grid. */
    const size_t guard = (size_t)g.ss;
    const size_t per = (size_t)g.total + 2 * guard;
    const int nown = g.s1 - g.s0;
    const bool slab_local = world > 1 && cfg->variant == IFL_F3_CONJUGATE_GRADIENTS;
    const int strip_shift = slab_local ? g.s0 : 0;
    const size_t per_slab = slab_local ? (size_t)(nown + 2) * g.ss : per;
    const int nfull = 4, nslab = 6;
    s->slab_bytes = sizeof(double) * (per * nfull + per_slab * nslab);
    CU(cudaMalloc(&s->skew_block, s->slab_bytes));
    CU(cudaMemset(s->skew_block, 0, s->slab_bytes));
    double** full_vecs[nfull] = {&s->sx.adiag, &s->sx.aplusx, &s->sx.aplusy, &s->sx.precon};
    for (int i = 0; i < nfull; i++) *full_vecs[i] = s->skew_block + per * i + guard;
    double** slab_vecs[nslab] = {&s->sx.r, &s->sx.p, &s->sx.zb[0], &s->sx.zb[1], &s->sx.sb[0], &s->sx.sb[1]};
    double* slab0 = s->skew_block + per * nfull;
    for (int i = 0; i < nslab; i++) *slab_vecs[i] = slab0 + per_slab * i + (long long)(1 - strip_shift) * g.ss;
    s->sx.p_first = s->sx.p + (long long)g.s0 * g.ss;       /* this rank's strips of p (what a solve starts by clearing) */
    const size_t pstrip = (size_t)g.tl * 32 * IFL_SKB * 3;
    const size_t npack = (size_t)nown * pstrip;
    CU(cudaMalloc(&s->pack_alloc, sizeof(double) * npack * 2));
    CU(cudaMemset(s->pack_alloc, 0, sizeof(double) * npack * 2));
    s->sx.pack_f = s->pack_alloc - (long long)g.s0 * pstrip;
    s->sx.pack_b = s->pack_alloc + npack - (long long)g.s0 * pstrip;
    s->slab_bytes += sizeof(double) * npack * 2;
    size_t nb = (size_t)strips_alloc * g.w;
    CU(cudaMalloc(&s->sx.bnd_f, sizeof(double) * nb * 2));
    s->sx.bnd_b = s->sx.bnd_f + nb;
    /* both inbox-row sets of the sweeps in ONE allocation: the neighbour ranks map it through CUDA IPC */
    s->sx.inbox_len = 2 * (g.tl + 64);
    const size_t nin = (size_t)strips_alloc * s->sx.inbox_len;
    CU(cudaMalloc(&s->sx.inbox_f, sizeof(double) * nin * 2));
    s->sx.inbox_b = s->sx.inbox_f + nin;
    s->sx.strips_alloc = strips_alloc;
    CU(cudaMalloc(&s->sx.strip_part, sizeof(double) * strips_alloc));
    CU(cudaMemset(s->sx.strip_part, 0, sizeof(double) * strips_alloc));
    CU(cudaMalloc(&s->sx.strip_ctr, sizeof(unsigned) * strips_alloc));
    CU(cudaMemset(s->sx.strip_ctr, 0, sizeof(unsigned) * strips_alloc));
    CU(cudaMalloc(&s->sx.rank_max, sizeof(unsigned long long) * (world > 16 ? world : 16) * 8));
    CU(cudaMemset(s->sx.rank_max, 0, sizeof(unsigned long long) * (world > 16 ? world : 16) * 8));
    CU(cudaMalloc(&s->sx.rowbuf, sizeof(double) * 8 * g.w));
    s->sx.comm = nullptr; s->sx.peer_inbox_f_next = nullptr; s->sx.peer_inbox_b_prev = nullptr;
    if (world > 1) {
        if (comm_init(s->comm, cfg->rank, world, cfg->comm_id, s->stream)) {
            snprintf(g_err, sizeof g_err, "%s", comm_last_error());
            return IFL_ERR_COMM;
        }
        if (comm_map_neighbour_inboxes(s->comm, s->sx.inbox_f, s->sx.rank_max)) {
            snprintf(g_err, sizeof g_err, "%s", comm_last_error());
            return IFL_ERR_COMM;
        }
        s->sx.comm = &s->comm;
        if (s->comm.peer_next_base) s->sx.peer_inbox_f_next = (double*)s->comm.peer_next_base;
        if (s->comm.peer_prev_base) s->sx.peer_inbox_b_prev = (double*)s->comm.peer_prev_base + nin;
        CU(cudaMemset(s->sx.rank_max, 0, sizeof(unsigned long long) * (world > 16 ? world : 16) * 8));
    }
    s->sx.npartials = solve_partials_needed(g);
    CU(cudaMalloc(&s->sx.partials, sizeof(double) * s->sx.npartials));
    CU(cudaMalloc(&s->sx.sc, sizeof(SolveScalars)));
    CU(cudaMemset(s->sx.sc, 0, sizeof(SolveScalars)));
    CU(cudaMalloc(&s->d_maxbits, sizeof(unsigned long long)));
    CU(cudaMallocHost(&s->h_sc, sizeof(SolveScalars) * 2));
    memset(s->h_sc, 0, sizeof(SolveScalars) * 2);
    CU(cudaMalloc(&s->scratch, sizeof(double) * (size_t)(s->w + 2) * (s->h + 2)));
    for (auto& e : s->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : s->ev) CU(cudaEventCreate(&e));
    s->sx.tolerance = cfg->tolerance;
    s->sx.mic_tau = cfg->mic_tau;
    s->sx.mic_sigma = cfg->mic_sigma;
    s->sx.sequential = (cfg->flags & IFL_FLAG_SEQUENTIAL_REDUCTIONS) ? 1 : 0;
    s->sx.separate_r_update = (cfg->flags & IFL_FLAG_SEPARATE_R_UPDATE) ? 1 : 0;
    s->sx.fl = nullptr;
    s->sx.mask_vector_ops = has_heat(s) ? 1 : 0;                /* F6:999-1064 vs F4:825-878 */
    if (has_solids(s)) {
        CU(cudaMalloc(&s->sx.fl, (size_t)per));
        CU(cudaMemset(s->sx.fl, 0, (size_t)per));
        s->sx.fl += guard;
    }
    s->sx.stream = s->stream;
    s->sx.const_matrix = 0; s->sx.const_scale = 0.0;
    s->sx.launches = &s->launches;
    s->sx.prof = &s->prof;
    s->sx.trace = nullptr;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaDeviceSynchronize());
    return IFL_OK;
}

int ifl_create(const ifl_config* cfg, ifl_solver** out) {
    if (!cfg || !out) INVALID("NULL argument");
    *out = nullptr;
    if (cfg->abi_version != IFL_ABI_VERSION) INVALID("abi_version mismatch");
    if (cfg->w < 2 || cfg->h < 2) INVALID("grid must be at least 2x2");
    if (cfg->variant < IFL_F1_MATRIXLESS || cfg->variant > IFL_F8_FLIP) INVALID("unknown variant");
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) INVALID("bad rank / world_size");
    if (cfg->solver_limit < 0) INVALID("solver_limit < 0");
    if ((long long)(cfg->w + 1) * (long long)(cfg->h + 1) > 2147483647LL) INVALID("grid too large: (w+1)*(h+1) must fit a 32-bit int, like the reference's array indices");
    if (cfg->variant == IFL_F8_FLIP) {
        /* ParticleQuantities sizes every particle array with w*h*_MaxPerCell (F8:1549) and indexes it with ints */
        if (cfg->particles_avg_per_cell <= 0 || cfg->particles_max_per_cell <= 0 || cfg->particles_min_per_cell < 0)
            INVALID("particle counts per cell must be positive");
        if (cfg->particles_avg_per_cell > cfg->particles_max_per_cell) INVALID("particles_avg_per_cell > particles_max_per_cell");
        if (cfg->particles_min_per_cell > cfg->particles_max_per_cell) INVALID("particles_min_per_cell > particles_max_per_cell");
        if ((long long)cfg->w * cfg->h * cfg->particles_max_per_cell > 2147483647LL)
            INVALID("w*h*particles_max_per_cell must fit a 32-bit int (F8:1549)");
    }
    if (cfg->world_size > 1) {
        if (cfg->variant < IFL_F3_CONJUGATE_GRADIENTS || cfg->variant > IFL_F7_VARIABLE_DENSITY) {
            snprintf(g_err, sizeof g_err, "several GPUs: implemented for the MIC(0)-PCG variants 3 (slab decomposition) and 4-7 "
                     "(slab-decomposed solver loop), not for variant %d", cfg->variant);
            return IFL_ERR_UNSUPPORTED;
        }
        if (cfg->flags & IFL_FLAG_SEQUENTIAL_REDUCTIONS) INVALID("IFL_FLAG_SEQUENTIAL_REDUCTIONS is single-GPU only");
        if ((cfg->h + 31) / 32 < cfg->world_size) INVALID("fewer 32-row strips than ranks");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        snprintf(g_err, sizeof g_err, "no usable CUDA device (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return IFL_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) INVALID("device ordinal out of range");
    ifl_solver* s = new (std::nothrow) ifl_solver();
    if (!s) INVALID("out of host memory");
    int rc = create_impl(cfg, s);
    if (rc != IFL_OK) { ifl_destroy(s); return rc; }
    *out = s;
    return IFL_OK;
}

int ifl_set_bodies(ifl_solver* s, const ifl_body* bodies, int32_t count) {
    if (!s) INVALID("NULL handle");
    if (count < 0 || (count > 0 && !bodies)) INVALID("bad body table");
    if (!has_solids(s)) {
        if (count == 0) return IFL_OK;
        snprintf(g_err, sizeof g_err, "variant %d has no solid bodies", s->cfg.variant);
        return IFL_ERR_UNSUPPORTED;
    }
    CU(cudaSetDevice(s->cfg.device));
    if (count > s->bodies_cap) {
        if (s->d_bodies) CU(cudaFree(s->d_bodies));
        CU(cudaMalloc(&s->d_bodies, sizeof(ifl_body) * count));
        s->bodies_cap = count;
    }
    CU(cudaStreamSynchronize(s->stream));
    if (count > 0) CU(cudaMemcpy(s->d_bodies, bodies, sizeof(ifl_body) * count, cudaMemcpyHostToDevice));
    s->nbodies = count;
    return IFL_OK;
}

int ifl_add_inflow(ifl_solver* s, double x, double y, double w, double h, double d, double t, double u, double v) {
    if (!s) INVALID("NULL handle");
    if (is_flip(s)) {            /* F8 calls addInflow inside update() (F8:1330): this sets that rectangle */
        double r[8] = {x, y, w, h, d, t, u, v};
        memcpy(s->inflow, r, sizeof r);
        return IFL_OK;
    }
    CU(cudaSetDevice(s->cfg.device));
    /* FluidSolver.addInflow F1:298 / F3:385 / F6:1233 */
    launch_add_inflow(s->stream, s->cfg.variant, s->d.q, s->d.src(), s->hx, x, y, x + w, y + h, d);
    if (has_heat(s)) {
        launch_add_inflow(s->stream, s->cfg.variant, s->t.q, s->t.src(), s->hx, x, y, x + w, y + h, t);
        s->launches++;
    }
    launch_add_inflow(s->stream, s->cfg.variant, s->u.q, s->u.src(), s->hx, x, y, x + w, y + h, u);
    launch_add_inflow(s->stream, s->cfg.variant, s->v.q, s->v.src(), s->hx, x, y, x + w, y + h, v);
    s->launches += 3;
    CU(cudaGetLastError());
    return IFL_OK;
}

/* Enqueue the solver loop in chunks; at most two chunks are in flight while the host
 * polls the `done` flag of the older one, so the device never idles on the host. */
static int run_solver_loop(ifl_solver* s, double timestep) {
    SolveCtx& c = s->sx;
    const int limit = s->cfg.solver_limit;
    const bool pcg = uses_pcg(s);
    const int chunk = pcg ? s->chunk_pcg : s->chunk_gs;
    const double gs_scale = timestep / (s->cfg.density * s->hx * s->hx);   /* F1:381 */
    int enq = 0, nchunks = 0;
    bool done = false;
    while (enq < limit && !done) {
        int n = limit - enq < chunk ? limit - enq : chunk;
        if (pcg) COMM(pcg_enqueue_iterations(c, n));
        else n = gs_enqueue_iterations(c, enq, n, gs_scale, limit);
        enq += n;
        int slot = nchunks & 1;
        CU(cudaMemcpyAsync(&s->h_sc[slot], c.sc, sizeof(SolveScalars), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaEventRecord(s->ev_chunk[slot], s->stream));
        nchunks++;
        if (nchunks >= 2) {
            int older = (nchunks - 2) & 1;
            CU(cudaEventSynchronize(s->ev_chunk[older]));
            if (s->h_sc[older].done) done = true;
        }
    }
    CU(cudaGetLastError());
    return IFL_OK;
}

static int fetch_status(ifl_solver* s, ifl_step_status* st) {
    CU(cudaMemcpyAsync(&s->h_sc[0], s->sx.sc, sizeof(SolveScalars), cudaMemcpyDeviceToHost, s->stream));
    if (has_heat(s)) CU(cudaMemcpyAsync(s->h_sc_heat, s->sc_heat, sizeof(SolveScalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));                       /* the one host wait of a step, and only when asked for */
    const SolveScalars& h = s->h_sc[0];
    if (is_flip(s)) s->last.particle_count = s->flip.count;
    s->last.iterations_pressure = h.iterations;
    s->last.residual_pressure = h.max_error;
    s->last.exceeded_budget = h.exceeded ? 1 : 0;
    s->last.nan_residual = h.nan;
    if (has_heat(s)) {
        const SolveScalars& hh = *s->h_sc_heat;
        s->last.iterations_heat = hh.iterations;
        s->last.residual_heat = hh.max_error;
        if (hh.exceeded) s->last.exceeded_budget |= 2;
        if (hh.nan) s->last.nan_residual = 1;
    }
    if (st) *st = s->last;
    return s->last.nan_residual ? IFL_ERR_NAN : IFL_OK;
}

static int stage_build_rhs(ifl_solver* s) {
    launch_build_rhs(s->stream, s->sx.g, s->u.src(), s->v.src(), s->sx.r, s->hx);
    s->launches++;
    return IFL_OK;
}
static int stage_build_matrix(ifl_solver* s, double timestep) {
    double scale = timestep / (s->cfg.density * s->hx * s->hx);             /* F3:436 */
    launch_build_matrix_f3(s->stream, s->sx.g, s->sx.adiag, s->sx.aplusx, s->sx.aplusy, scale);
    s->sx.const_matrix = 1;       /* the arrays are still filled: MIC(0) build and read_field use them; writing one of
                                   * them through ifl_write_field switches the solver to the array-reading kernels */
    s->sx.const_scale = scale;
    s->launches++;
    return IFL_OK;
}
static int stage_project(ifl_solver* s, double timestep) {
    SolveCtx& c = s->sx;
    if (uses_pcg(s)) COMM(pcg_enqueue_init(c)); else gs_enqueue_begin(c);
    int rc = run_solver_loop(s, timestep);
    if (rc) return rc;
    if (uses_pcg(s)) COMM(pcg_enqueue_finish(c, s->cfg.solver_limit));
    else gs_enqueue_finish(c, s->cfg.solver_limit);
    /* F4-F7 on several GPUs: every rank solved for its strips of p; the stages that follow (applyPressure, T <- p) are
     * repeated by every rank on the whole grid, so every rank needs every strip (strips are contiguous in the skewed
     * layout: one grouped broadcast, in place) */
    if (replicated_stages(s)) COMM(comm_all_gather_strips(s->comm, c.p, c.g.nstrips, (size_t)c.g.ss));
    return IFL_OK;
}
static int stage_apply_pressure(ifl_solver* s, double timestep) {
    double scale = timestep / (s->cfg.density * s->hx);                      /* F1:433 */
    launch_apply_pressure(s->stream, s->sx.g, s->sx.p, s->u.src(), s->v.src(), scale);
    s->launches++;
    return IFL_OK;
}
/* rows of quantity q that `rank` owns in the slab decomposition: the rows of its strips; the extra row h of v belongs to
 * the rank that owns row h-1.  (Everything on a single GPU.) */
static void owned_rows(const ifl_solver* s, const Quantity& q, int rank, int* r0, int* r1) {
    int s0, s1;
    slab_range(s->sx.g.nstrips, s->cfg.world_size, rank, &s0, &s1);
    *r0 = 32 * s0 < s->h ? 32 * s0 : s->h;
    *r1 = 32 * s1 < s->h ? 32 * s1 : s->h;
    if (*r1 == s->h) *r1 = q.q.h;
}
static int stage_advect(ifl_solver* s, Quantity& q, double timestep) {
    int r0, r1;
    owned_rows(s, q, s->cfg.rank, &r0, &r1);
    launch_advect(s->stream, s->cfg.variant, q.q, q.src(), q.dst(), s->u.q, s->u.src(), s->v.q, s->v.src(),
                  timestep, s->hx, r0, r1);
    s->launches++;
    return IFL_OK;
}

/* ---- F4-F7 stages ----------------------------------------------------------------------------- */
static int stage_fill_solid_fields(ifl_solver* s) {
    /* FluidQuantity.fillSolidFields returns at once for an empty body list (F4:330) */
    if (s->nbodies > 0) {
        for (Quantity* q : {&s->d, &s->t, &s->u, &s->v}) {
            if (!q->present) continue;
            launch_fill_solid_fields(s->stream, q->sq, s->d_bodies, s->nbodies, s->hx, s->cfg.variant);
            s->launches += s->cfg.variant >= IFL_F5_CURVED_BOUNDARIES ? 2 : 1;
        }
    }
    launch_cell_flags(s->stream, s->sx.g, s->d.sq.cell, s->sx.fl);
    s->launches++;
    return IFL_OK;
}
static int stage_set_boundary(ifl_solver* s) {
    launch_set_boundary(s->stream, s->w, s->h, s->d.sq.cell, s->d.sq.body, s->d_bodies, s->u.src(), s->v.src(), s->hx);
    s->launches++;
    return IFL_OK;
}
static int stage_build_rhs_solid(ifl_solver* s) {
    launch_build_rhs_solid(s->stream, s->sx.g, s->d.view(), s->u.view(), s->v.view(), s->u.src(), s->v.src(),
                           s->d_bodies, s->nbodies, s->sx.r, s->hx, s->cfg.variant);
    s->launches++;
    return IFL_OK;
}
static int stage_compute_densities(ifl_solver* s) {
    launch_compute_densities(s->stream, s->w, s->h, s->d.src(), s->t.src(), s->udens, s->vdens,
                             s->cfg.density, s->cfg.density_soot, s->cfg.t_ambient);
    s->launches++;
    return IFL_OK;
}
static int stage_build_matrix_solid(ifl_solver* s, double timestep) {
    const int var = s->cfg.variant;
    double scale; int mode;
    if (var >= IFL_F7_VARIABLE_DENSITY) { scale = timestep / (s->hx * s->hx); mode = 2; }                 /* F7:878 */
    else { scale = timestep / (s->cfg.density * s->hx * s->hx); mode = var >= IFL_F5_CURVED_BOUNDARIES ? 1 : 0; }  /* F4:714, F5:826 */
    launch_build_matrix_solid(s->stream, s->sx.g, s->d.sq.cell, s->u.sq.volume, s->v.sq.volume, s->udens, s->vdens,
                              scale, mode, s->sx.adiag, s->sx.aplusx, s->sx.aplusy);
    s->launches++;
    return IFL_OK;
}
__global__ void k_copy_scalars(const SolveScalars* a, SolveScalars* b) { *b = *a; }
static int stage_project(ifl_solver* s, double timestep);
/* F6:1190-1201 (=F7:1221-1232): r <- T; heat matrix; MIC; project; T <- p */
static int stage_heat_solve(ifl_solver* s, double timestep) {
    launch_rows_to_skew(s->stream, s->sx.g, s->t.src(), s->sx.r);
    double scale = s->cfg.diffusion * timestep * 1.0 / (s->hx * s->hx);                                    /* F6:888 */
    launch_build_matrix_solid(s->stream, s->sx.g, s->d.sq.cell, nullptr, nullptr, nullptr, nullptr, scale, 3,
                              s->sx.adiag, s->sx.aplusx, s->sx.aplusy);
    s->launches += 2;
    launch_build_precon(s->sx);
    int rc = stage_project(s, timestep);
    if (rc) return rc;
    k_copy_scalars<<<1, 1, 0, s->stream>>>(s->sx.sc, s->sc_heat);
    launch_skew_to_rows(s->stream, s->sx.g, s->sx.p, s->t.src());
    s->launches += 2;
    return IFL_OK;
}
static int stage_add_buoyancy(ifl_solver* s, double timestep) {
    launch_add_buoyancy(s->stream, s->w, s->h, s->d.src(), s->t.src(), s->v.src(), timestep, s->cfg.gravity,
                        s->cfg.density, s->cfg.density_soot, s->cfg.t_ambient);
    s->launches++;
    return IFL_OK;
}
static int stage_apply_pressure_solid(ifl_solver* s, double timestep) {
    const bool vd = has_vardens(s);
    double scale = vd ? timestep / s->hx : timestep / (s->cfg.density * s->hx);                           /* F7:1149 / F4:928 */
    launch_apply_pressure_solid(s->stream, s->sx.g, s->d.sq.cell, s->sx.p, s->u.src(), s->v.src(), s->udens, s->vdens,
                                scale, vd ? 1 : 0);
    s->launches++;
    return IFL_OK;
}
/* FluidQuantity.extrapolate F4:452: sweeps until no pending solid cell can be filled any more.  The sweeps are enqueued
 * in chunks; a sweep that finds the previous one idle raises a device flag and every later one returns at once, so the
 * host only polls the flag of the chunk before last (two chunks in flight, the device never waits for the host). */
static int stage_extrapolate(ifl_solver* s, Quantity& q) {
    launch_extrap_fill_mask(s->stream, q.sq, s->extrap_state);
    CU(cudaMemsetAsync(s->d_extrap_ctl, 0, 4 * sizeof(unsigned int), s->stream));
    s->launches++;
    const int chunk = 16;
    const int max_sweeps = q.q.w + q.q.h + 2;
    int k = 0, nchunks = 0;
    bool done = false;
    while (k < max_sweeps && !done) {
        for (int i = 0; i < chunk; i++, k++) launch_extrap_sweep(s->stream, q.sq, q.src(), s->extrap_state, s->d_extrap_ctl, k);
        s->launches += chunk;
        const int slot = nchunks & 1;
        CU(cudaMemcpyAsync(&s->h_extrap_flag[slot], s->d_extrap_ctl + 3, sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaEventRecord(s->ev_chunk[slot], s->stream));
        nchunks++;
        if (nchunks >= 2) {
            const int older = (nchunks - 2) & 1;
            CU(cudaEventSynchronize(s->ev_chunk[older]));
            if (s->h_extrap_flag[older]) done = true;
        }
    }
    return IFL_OK;
}
static int stage_advect_solid(ifl_solver* s, Quantity& q, double timestep) {
    launch_advect_solid(s->stream, q.q, q.src(), q.dst(), s->u.q, s->u.src(), s->v.q, s->v.src(),
                        q.sq.cell, q.sq.body, s->d_bodies, timestep, s->hx);
    s->launches++;
    return IFL_OK;
}
static int stage_project(ifl_solver* s, double timestep);
static int stage_apply_pressure(ifl_solver* s, double timestep);

/* FluidSolver.update F4:617 / F5:1089 / F6:1181 / F7:1212 */
static int update_solid(ifl_solver* s, double timestep) {
    int rc;
    CU(cudaEventRecord(s->ev[0], s->stream));
    stage_fill_solid_fields(s);
    if (has_heat(s)) {
        if ((rc = stage_heat_solve(s, timestep))) return rc;
        if ((rc = stage_extrapolate(s, s->t))) return rc;
        stage_add_buoyancy(s, timestep);
    }
    stage_set_boundary(s);
    stage_build_rhs_solid(s);
    if (has_vardens(s)) stage_compute_densities(s);
    stage_build_matrix_solid(s, timestep);
    launch_build_precon(s->sx);
    if ((rc = stage_project(s, timestep))) return rc;
    stage_apply_pressure_solid(s, timestep);
    if ((rc = stage_extrapolate(s, s->d))) return rc;
    if ((rc = stage_extrapolate(s, s->u))) return rc;
    if ((rc = stage_extrapolate(s, s->v))) return rc;
    stage_set_boundary(s);
    CU(cudaEventRecord(s->ev[1], s->stream));
    stage_advect_solid(s, s->d, timestep);
    if (has_heat(s)) stage_advect_solid(s, s->t, timestep);
    stage_advect_solid(s, s->u, timestep);
    stage_advect_solid(s, s->v, timestep);
    s->d.swap(); s->u.swap(); s->v.swap();
    if (has_heat(s)) s->t.swap();
    CU(cudaEventRecord(s->ev[2], s->stream));
    s->ev_valid = true;
    CU(cudaGetLastError());
    return IFL_OK;
}

/* ---- F8 ----------------------------------------------------------------------------------------- */
static int inflow_now(ifl_solver* s, const double* r) {
    const double x = r[0], y = r[1], w = r[2], h = r[3];
    launch_add_inflow(s->stream, s->cfg.variant, s->d.q, s->d.src(), s->hx, x, y, x + w, y + h, r[4]);
    launch_add_inflow(s->stream, s->cfg.variant, s->t.q, s->t.src(), s->hx, x, y, x + w, y + h, r[5]);
    launch_add_inflow(s->stream, s->cfg.variant, s->u.q, s->u.src(), s->hx, x, y, x + w, y + h, r[6]);
    launch_add_inflow(s->stream, s->cfg.variant, s->v.q, s->v.src(), s->hx, x, y, x + w, y + h, r[7]);
    s->launches += 4;
    return IFL_OK;
}
/* the part of the F8 constructor that needs the body list: initParticles + gridToParticles(1.0) (F8:887-891) */
static int flip_ensure(ifl_solver* s) {
    FlipCtx& f = s->flip;
    if (f.ready) return IFL_OK;
    CU(flip_init_particles(f, s->stream, s->cfg.particles_avg_per_cell, s->cfg.rng_seed, s->d_bodies, s->nbodies, s->hx, &s->launches));
    flip_grid_to_particles(f, s->stream, 1.0, s->d.q, s->d.src(), s->t.q, s->t.src(), s->u.q, s->u.src(), s->v.q, s->v.src(), &s->launches);
    f.ready = true;
    return IFL_OK;
}
static int flip_extrapolate_q(ifl_solver* s, Quantity& q) {
    CU(flip_extrapolate(s->flip, s->stream, q.sq, q.src(), s->extrap_state, s->d_progress, s->h_progress,
                        (s->cfg.flags & IFL_FLAG_SEQUENTIAL_EXTRAPOLATE) ? 1 : 0, &s->launches));
    return IFL_OK;
}
/* ParticleQuantities.particlesToGrid F8:1761 */
static int stage_particles_to_grid(ifl_solver* s) {
    FlipCtx& f = s->flip;
    int rc;
    if ((rc = flip_ensure(s))) return rc;
    Quantity* qs[4] = {&s->d, &s->t, &s->u, &s->v};
    for (int q = 0; q < 4; q++) {
        if (q != 1) CU(flip_bin(f, s->stream, qs[q]->q, &s->launches));        /* d and t share their sample positions */
        flip_gather(f, s->stream, qs[q]->q, f.prop[q], qs[q]->src(), qs[q]->sq.cell, &s->launches);
        if ((rc = flip_extrapolate_q(s, *qs[q]))) return rc;      /* leaves cell_start / vals_b (the binning) untouched */
    }
    CU(flip_count_prune_seed(f, s->stream, s->cfg.particles_max_per_cell, s->cfg.particles_min_per_cell, s->cfg.rng_seed,
                             s->d_bodies, s->nbodies, s->hx, s->d.q, s->d.src(), s->t.q, s->t.src(), s->u.q, s->u.src(),
                             s->v.q, s->v.src(), &s->launches));
    s->last.particle_count = f.count;
    return IFL_OK;
}
static int stage_grid_to_particles(ifl_solver* s) {
    int rc;
    if ((rc = flip_ensure(s))) return rc;
    const double a = s->cfg.flip_alpha;
    Quantity* qs[4] = {&s->d, &s->t, &s->u, &s->v};
    for (Quantity* q : qs) flip_diff(s->stream, q->src(), q->dst(), q->count(), a, -1, &s->launches);          /* F8:346 */
    flip_grid_to_particles(s->flip, s->stream, a, s->d.q, s->d.src(), s->t.q, s->t.src(), s->u.q, s->u.src(), s->v.q, s->v.src(),
                           &s->launches);
    for (Quantity* q : qs) flip_diff(s->stream, q->src(), q->dst(), q->count(), a, +1, &s->launches);          /* F8:355 */
    return IFL_OK;
}
static int stage_advect_particles(ifl_solver* s, double timestep) {
    int rc;
    if ((rc = flip_ensure(s))) return rc;
    flip_advect_particles(s->flip, s->stream, timestep, s->hx, s->u.q, s->u.src(), s->v.q, s->v.src(), s->d_bodies, s->nbodies,
                          &s->launches);
    return IFL_OK;
}
static int stage_extrapolate_any(ifl_solver* s, Quantity& q) {
    return is_flip(s) ? flip_extrapolate_q(s, q) : stage_extrapolate(s, q);
}
/* FluidSolver.update F8:1306 */
static int update_flip(ifl_solver* s, double timestep) {
    int rc;
    CU(cudaEventRecord(s->ev[0], s->stream));
    if ((rc = flip_ensure(s))) return rc;
    stage_fill_solid_fields(s);
    if ((rc = stage_particles_to_grid(s))) return rc;
    launch_cell_flags(s->stream, s->sx.g, s->d.sq.cell, s->sx.fl);      /* _d._cell went FLUID -> EMPTY -> FLUID meanwhile */
    s->launches++;
    for (Quantity* q : {&s->d, &s->t, &s->u, &s->v})                    /* copy() F8:339 */
        CU(cudaMemcpyAsync(q->dst(), q->src(), sizeof(double) * q->count(), cudaMemcpyDeviceToDevice, s->stream));
    inflow_now(s, s->inflow);                                            /* F8:1330 */
    if ((rc = stage_heat_solve(s, timestep))) return rc;
    if ((rc = flip_extrapolate_q(s, s->t))) return rc;
    stage_add_buoyancy(s, timestep);
    stage_set_boundary(s);
    stage_build_rhs_solid(s);
    stage_compute_densities(s);
    stage_build_matrix_solid(s, timestep);
    launch_build_precon(s->sx);
    if ((rc = stage_project(s, timestep))) return rc;
    stage_apply_pressure_solid(s, timestep);
    if ((rc = flip_extrapolate_q(s, s->d))) return rc;
    if ((rc = flip_extrapolate_q(s, s->u))) return rc;
    if ((rc = flip_extrapolate_q(s, s->v))) return rc;
    stage_set_boundary(s);
    CU(cudaEventRecord(s->ev[1], s->stream));
    if ((rc = stage_grid_to_particles(s))) return rc;
    if ((rc = stage_advect_particles(s, timestep))) return rc;
    CU(cudaEventRecord(s->ev[2], s->stream));
    s->ev_valid = true;
    CU(cudaGetLastError());
    return IFL_OK;
}

/* ---- slab decomposition (F3) --------------------------------------------------------------------------------------- */
/* Halo rows of one row-major quantity: every rank receives the rows [r0 - above, r1 + below) of q it does not own from
 * their owners and sends the rows it owns that others need -- one grouped set of ncclSend/ncclRecv straight from / into
 * the field (rows are contiguous).  `above` / `below` are the same on every rank. */
static int slab_exchange_halo(ifl_solver* s, Quantity& q, int above, int below) {
    const int world = s->cfg.world_size, me = s->cfg.rank;
    std::vector<CommXfer> x;
    int m0, m1;
    owned_rows(s, q, me, &m0, &m1);
    const long long width = q.q.w;
    auto clip = [&](long long v) { return v < 0 ? 0LL : (v > q.q.h ? (long long)q.q.h : v); };
    const long long need0 = clip((long long)m0 - above), need1 = clip((long long)m1 + below);
    for (int p = 0; p < world; p++) {
        if (p == me) continue;
        int p0, p1;
        owned_rows(s, q, p, &p0, &p1);
        /* what I need from p */
        long long lo = need0 > p0 ? need0 : p0, hi = need1 < p1 ? need1 : p1;
        if (hi > lo) x.push_back(CommXfer{p, q.src() + lo * width, (size_t)((hi - lo) * width), 0});
        /* what p needs from me */
        const long long pn0 = clip((long long)p0 - above), pn1 = clip((long long)p1 + below);
        lo = pn0 > m0 ? pn0 : m0; hi = pn1 < m1 ? pn1 : m1;
        if (hi > lo) x.push_back(CommXfer{p, q.src() + lo * width, (size_t)((hi - lo) * width), 1});
    }
    COMM(comm_transfers(s->comm, x.data(), (int)x.size()));
    return IFL_OK;
}
/* Rows a semi-Lagrangian backtrace can leave its slab by: ceil(timestep * max|v| / hx) + 4 (the three RK3 samples stay
 * within timestep*max|v|/hx of the start, F3:255-275; the Catmull-Rom stencil adds two rows, the face offsets one).
 * max|v| is the CFL reduction of maxTimestep (F1:309-333) restricted to the vertical component, over ALL ranks. */
static int slab_advection_halo(ifl_solver* s, double timestep, int* halo) {
    int r0, r1;
    owned_rows(s, s->v, s->cfg.rank, &r0, &r1);
    unsigned long long* slot = s->sx.rank_max + s->cfg.rank;
    CU(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), s->stream));
    launch_max_abs_rows(s->stream, s->v.src(), s->w, r0, r1, slot);
    s->launches++;
    COMM(comm_all_gather_u64(s->comm, s->sx.rank_max));
    unsigned long long h_bits[64];
    CU(cudaMemcpyAsync(h_bits, s->sx.rank_max, sizeof(unsigned long long) * s->cfg.world_size, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    double vmax = 0.0;
    bool bad = false;
    for (int r = 0; r < s->cfg.world_size; r++) {
        double v; memcpy(&v, &h_bits[r], sizeof v);
        if (!(v == v) || v > 1e300) bad = true;
        if (v > vmax) vmax = v;
    }
    const double cells = std::fabs(timestep) * vmax / s->hx;
    *halo = (bad || cells > (double)s->h) ? s->h : (int)std::ceil(cells) + 4;
    s->last_halo = *halo;
    return IFL_OK;
}
/* FluidSolver.update F3:361 on a slab: every rank computes only the rows it owns */
static int update_slab(ifl_solver* s, double timestep) {
    int rc;
    CU(cudaEventRecord(s->ev[0], s->stream));
    if ((rc = slab_exchange_halo(s, s->v, 0, 1))) return rc;        /* buildRhs of my last row reads v of the row below */
    stage_build_rhs(s);
    stage_build_matrix(s, timestep);
    launch_build_precon(s->sx);
    if ((rc = stage_project(s, timestep))) return rc;
    COMM(pcg_exchange_pressure_halo(s->sx));                        /* applyPressure of my first row reads p of the row above */
    stage_apply_pressure(s, timestep);
    CU(cudaEventRecord(s->ev[1], s->stream));
    int halo = 0;
    if ((rc = slab_advection_halo(s, timestep, &halo))) return rc;
    for (Quantity* q : {&s->d, &s->u, &s->v})
        if ((rc = slab_exchange_halo(s, *q, halo, halo + 1))) return rc;
    stage_advect(s, s->d, timestep);
    stage_advect(s, s->u, timestep);
    stage_advect(s, s->v, timestep);
    s->d.swap(); s->u.swap(); s->v.swap();
    CU(cudaEventRecord(s->ev[2], s->stream));
    s->ev_valid = true;
    CU(cudaGetLastError());
    return IFL_OK;
}

static int update_impl(ifl_solver* s, double timestep) {
    if (is_flip(s)) return update_flip(s, timestep);
    if (has_solids(s)) return update_solid(s, timestep);
    if (s->cfg.world_size > 1) return update_slab(s, timestep);
    /* FluidSolver.update  F1:276 / F2:322 / F3:361 */
    CU(cudaEventRecord(s->ev[0], s->stream));
    stage_build_rhs(s);
    if (uses_pcg(s)) {
        stage_build_matrix(s, timestep);
        launch_build_precon(s->sx);
    }
    int rc = stage_project(s, timestep);
    if (rc) return rc;
    stage_apply_pressure(s, timestep);
    CU(cudaEventRecord(s->ev[1], s->stream));
    stage_advect(s, s->d, timestep);
    stage_advect(s, s->u, timestep);
    stage_advect(s, s->v, timestep);
    s->d.swap(); s->u.swap(); s->v.swap();
    CU(cudaEventRecord(s->ev[2], s->stream));
    s->ev_valid = true;
    CU(cudaGetLastError());
    return IFL_OK;
}

int ifl_update(ifl_solver* s, double timestep, ifl_step_status* status) {
    if (!s) INVALID("NULL handle");
    CU(cudaSetDevice(s->cfg.device));
    int rc = update_impl(s, timestep);
    if (rc) return rc;
    if (status) return fetch_status(s, status);
    return IFL_OK;
}

int ifl_run_steps(ifl_solver* s, int32_t n_steps, double timestep, double x, double y, double w, double h,
                  double d, double t, double u, double v, ifl_step_status* last_status) {
    if (!s) INVALID("NULL handle");
    if (n_steps < 0) INVALID("n_steps < 0");
    CU(cudaSetDevice(s->cfg.device));
    for (int i = 0; i < n_steps; i++) {
        int rc = ifl_add_inflow(s, x, y, w, h, d, t, u, v);
        if (rc) return rc;
        rc = update_impl(s, timestep);
        if (rc) return rc;
    }
    if (last_status) return fetch_status(s, last_status);
    return IFL_OK;
}

int ifl_run_stage(ifl_solver* s, int32_t stage, double timestep, ifl_step_status* status) {
    if (!s) INVALID("NULL handle");
    CU(cudaSetDevice(s->cfg.device));
    int rc = IFL_OK;
    if (is_flip(s) && ((stage >= IFL_STAGE_ADVECT_D && stage <= IFL_STAGE_SWAP))) goto unsupported;
    switch (stage) {
        case IFL_STAGE_BUILD_RHS: rc = has_solids(s) ? stage_build_rhs_solid(s) : stage_build_rhs(s); break;
        case IFL_STAGE_BUILD_MATRIX:
            if (!uses_pcg(s)) goto unsupported;
            rc = has_solids(s) ? stage_build_matrix_solid(s, timestep) : stage_build_matrix(s, timestep); break;
        case IFL_STAGE_BUILD_PRECON:
            if (!uses_pcg(s)) goto unsupported;
            launch_build_precon(s->sx); break;
        case IFL_STAGE_PROJECT: rc = stage_project(s, timestep); break;
        case IFL_STAGE_APPLY_PRESSURE: rc = has_solids(s) ? stage_apply_pressure_solid(s, timestep) : stage_apply_pressure(s, timestep); break;
        case IFL_STAGE_ADVECT_D: rc = has_solids(s) ? stage_advect_solid(s, s->d, timestep) : stage_advect(s, s->d, timestep); break;
        case IFL_STAGE_ADVECT_T:
            if (!has_heat(s)) goto unsupported;
            rc = stage_advect_solid(s, s->t, timestep); break;
        case IFL_STAGE_ADVECT_U: rc = has_solids(s) ? stage_advect_solid(s, s->u, timestep) : stage_advect(s, s->u, timestep); break;
        case IFL_STAGE_ADVECT_V: rc = has_solids(s) ? stage_advect_solid(s, s->v, timestep) : stage_advect(s, s->v, timestep); break;
        case IFL_STAGE_SWAP: s->d.swap(); s->u.swap(); s->v.swap(); if (has_heat(s)) s->t.swap(); break;
        case IFL_STAGE_FILL_SOLID_FIELDS: if (!has_solids(s)) goto unsupported; rc = stage_fill_solid_fields(s); break;
        case IFL_STAGE_SET_BOUNDARY: if (!has_solids(s)) goto unsupported; rc = stage_set_boundary(s); break;
        case IFL_STAGE_EXTRAPOLATE_D: if (!has_solids(s)) goto unsupported; rc = stage_extrapolate_any(s, s->d); break;
        case IFL_STAGE_EXTRAPOLATE_T: if (!has_heat(s)) goto unsupported; rc = stage_extrapolate_any(s, s->t); break;
        case IFL_STAGE_EXTRAPOLATE_U: if (!has_solids(s)) goto unsupported; rc = stage_extrapolate_any(s, s->u); break;
        case IFL_STAGE_EXTRAPOLATE_V: if (!has_solids(s)) goto unsupported; rc = stage_extrapolate_any(s, s->v); break;
        case IFL_STAGE_HEAT_SOLVE: if (!has_heat(s)) goto unsupported; rc = stage_heat_solve(s, timestep); break;
        case IFL_STAGE_ADD_BUOYANCY: if (!has_heat(s)) goto unsupported; rc = stage_add_buoyancy(s, timestep); break;
        case IFL_STAGE_PARTICLES_TO_GRID: if (!is_flip(s)) goto unsupported; rc = stage_particles_to_grid(s); break;
        case IFL_STAGE_GRID_TO_PARTICLES: if (!is_flip(s)) goto unsupported; rc = stage_grid_to_particles(s); break;
        case IFL_STAGE_ADVECT_PARTICLES: if (!is_flip(s)) goto unsupported; rc = stage_advect_particles(s, timestep); break;
        case IFL_STAGE_COMPUTE_DENSITIES: if (!has_vardens(s)) goto unsupported; rc = stage_compute_densities(s); break;
        case IFL_STAGE_APPLY_PRECON:
            if (!uses_pcg(s)) goto unsupported;
            COMM(launch_apply_precon(s->sx)); break;
        case IFL_STAGE_MATVEC:
            if (!uses_pcg(s)) goto unsupported;
            COMM(launch_matvec(s->sx)); break;
        default: goto unsupported;
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    if (status) return fetch_status(s, status);
    return IFL_OK;
unsupported:
    snprintf(g_err, sizeof g_err, "stage %d is not available for variant %d in this build", stage, s->cfg.variant);
    return IFL_ERR_UNSUPPORTED;
}

int ifl_max_timestep(ifl_solver* s, double* out) {
    if (!s || !out) INVALID("NULL argument");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemsetAsync(s->d_maxbits, 0, sizeof(unsigned long long), s->stream));
    launch_max_velocity(s->stream, s->w, s->h, s->u.src(), s->v.src(), s->d_maxbits);
    s->launches++;
    unsigned long long bits = 0;
    CU(cudaMemcpyAsync(&bits, s->d_maxbits, sizeof bits, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    double maxVelocity;
    memcpy(&maxVelocity, &bits, sizeof bits);
    double maxTimestep = 2.0 * s->hx / maxVelocity;                          /* F1:330 */
    *out = maxTimestep < 1.0 ? maxTimestep : 1.0;
    if (maxTimestep != maxTimestep) *out = maxTimestep;
    return IFL_OK;
}

/* field lookup: device pointer, element count and storage kind */
enum { KIND_ROWS = 0, KIND_SKEW = 1, KIND_U8 = 2, KIND_I32 = 3 };
static int field_ptr(ifl_solver* s, int32_t field, void** ptr, size_t* count, int* kind, size_t* width = nullptr) {
    *kind = KIND_ROWS;
    size_t n = (size_t)s->w * s->h;
    size_t wdummy;
    if (!width) width = &wdummy;
    *width = (size_t)s->w;                        /* row length; u-shaped fields and phi set their own */
    auto bad = [&]() {
        snprintf(g_err, sizeof g_err, "field %d does not exist for variant %d", field, s->cfg.variant);
        return IFL_ERR_INVALID;
    };
    switch (field) {
        case IFL_FIELD_D: *ptr = s->d.src(); *count = s->d.count(); return IFL_OK;
        case IFL_FIELD_U: *ptr = s->u.src(); *count = s->u.count(); *width = (size_t)s->w + 1; return IFL_OK;
        case IFL_FIELD_V: *ptr = s->v.src(); *count = s->v.count(); return IFL_OK;
        case IFL_FIELD_D_DST: *ptr = s->d.dst(); *count = s->d.count(); return IFL_OK;
        case IFL_FIELD_U_DST: *ptr = s->u.dst(); *count = s->u.count(); *width = (size_t)s->w + 1; return IFL_OK;
        case IFL_FIELD_V_DST: *ptr = s->v.dst(); *count = s->v.count(); return IFL_OK;
        case IFL_FIELD_T: if (!has_heat(s)) return bad(); *ptr = s->t.src(); *count = s->t.count(); return IFL_OK;
        case IFL_FIELD_T_DST: if (!has_heat(s)) return bad(); *ptr = s->t.dst(); *count = s->t.count(); return IFL_OK;
        case IFL_FIELD_UDENSITY: if (!has_vardens(s)) return bad(); *ptr = s->udens; *count = s->u.count(); *width = (size_t)s->w + 1; return IFL_OK;
        case IFL_FIELD_VDENSITY: if (!has_vardens(s)) return bad(); *ptr = s->vdens; *count = s->v.count(); return IFL_OK;
        case IFL_FIELD_P: *ptr = s->sx.p; break;
        case IFL_FIELD_R: *ptr = s->sx.r; break;
        case IFL_FIELD_Z: case IFL_FIELD_S: {          /* the current generation (SolveScalars.zpar) */
            const int zp = solve_parity(s->sx);
            if (zp < 0) { snprintf(g_err, sizeof g_err, "cannot read the solver state from the device"); return IFL_ERR_CUDA; }
            *ptr = field == IFL_FIELD_Z ? s->sx.zb[zp] : s->sx.sb[zp];
            break;
        }
        case IFL_FIELD_PRECON: *ptr = s->sx.precon; break;
        case IFL_FIELD_ADIAG: *ptr = s->sx.adiag; break;
        case IFL_FIELD_APLUSX: *ptr = s->sx.aplusx; break;
        case IFL_FIELD_APLUSY: *ptr = s->sx.aplusy; break;
        default: {
            if (!has_solids(s) || field < IFL_FIELD_VOLUME_D || field > IFL_FIELD_PHI_V) return bad();
            const int qi = (field - IFL_FIELD_VOLUME_D) % 4, grp = (field - IFL_FIELD_VOLUME_D) / 4;
            Quantity* q = qi == 0 ? &s->d : qi == 1 ? &s->t : qi == 2 ? &s->u : &s->v;
            if (!q->present) return bad();
            *count = q->count();
            *width = (size_t)q->q.w;
            switch (grp) {
                case 0: if (s->cfg.variant < IFL_F5_CURVED_BOUNDARIES) return bad(); *ptr = q->sq.volume; return IFL_OK;
                case 1: *ptr = q->sq.normalX; return IFL_OK;
                case 2: *ptr = q->sq.normalY; return IFL_OK;
                case 3: *ptr = q->sq.cell; *kind = KIND_U8; return IFL_OK;
                case 4: *ptr = q->sq.body; *kind = KIND_I32; return IFL_OK;
                case 5: if (s->cfg.variant < IFL_F5_CURVED_BOUNDARIES) return bad();
                        *ptr = q->sq.phi; *count = (size_t)(q->q.w + 1) * (q->q.h + 1); *width = (size_t)q->q.w + 1; return IFL_OK;
            }
            return bad();
        }
    }
    *count = n;
    *kind = KIND_SKEW;
    return IFL_OK;
}

int ifl_field_size(const ifl_solver* s, int32_t field, size_t* count) {
    if (!s || !count) INVALID("NULL argument");
    void* p; int kind;
    return field_ptr(const_cast<ifl_solver*>(s), field, &p, count, &kind);
}

/* rows [row0, row0 + nrows) of a field; nrows < 0: the whole field */
static int field_io(ifl_solver* s, int32_t field, double* host, size_t count, long long row0, long long nrows, bool write) {
    if (!s || !host) INVALID("NULL argument");
    CU(cudaSetDevice(s->cfg.device));
    void* p; size_t n, width; int kind;
    int rc = field_ptr(s, field, &p, &n, &kind, &width);
    if (rc) return rc;
    const bool whole = nrows < 0;
    if (whole && s->cfg.world_size > 1 && !replicated_stages(s)) {
        snprintf(g_err, sizeof g_err, "slab decomposition: a rank holds only its rows of field %d -- use ifl_%s_field_rows", field,
                 write ? "write" : "read");
        return IFL_ERR_UNSUPPORTED;
    }
    const size_t total_rows = n / width;
    if (whole) { row0 = 0; nrows = (long long)total_rows; }
    if (row0 < 0 || nrows < 0 || (size_t)(row0 + nrows) > total_rows) INVALID("row range outside the field");
    if ((size_t)nrows * width != count) INVALID(whole ? "count does not match ifl_field_size" : "count does not match nrows * row length");
    if (kind == KIND_SKEW && s->cfg.world_size > 1 && !replicated_stages(s)) {
        const long long o0 = 32LL * s->sx.g.s0, o1 = 32LL * s->sx.g.s1 < s->h ? 32LL * s->sx.g.s1 : (long long)s->h;
        if (row0 < o0 || row0 + nrows > o1) INVALID("slab decomposition: a solver vector holds only the rows this rank owns (ifl_slab_rows)");
    }
    if (write) {
        if (kind == KIND_U8 || kind == KIND_I32) {
            snprintf(g_err, sizeof g_err, "field %d (cell/body) is derived by fillSolidFields and is read-only", field);
            return IFL_ERR_UNSUPPORTED;
        }
        if (field == IFL_FIELD_ADIAG || field == IFL_FIELD_APLUSX || field == IFL_FIELD_APLUSY) s->sx.const_matrix = 0;
        if (kind == KIND_SKEW) {
            CU(cudaMemcpyAsync(s->scratch, host, sizeof(double) * count, cudaMemcpyHostToDevice, s->stream));
            launch_rows_to_skew(s->stream, s->sx.g, s->scratch, (double*)p, (int)row0, (int)(row0 + nrows));
            s->launches++;
        } else {
            CU(cudaMemcpyAsync((double*)p + (size_t)row0 * width, host, sizeof(double) * count, cudaMemcpyHostToDevice, s->stream));
        }
        CU(cudaStreamSynchronize(s->stream));
        return IFL_OK;
    }
    const double* srcp = (const double*)p + (size_t)row0 * width;
    if (kind == KIND_SKEW) {
        launch_skew_to_rows(s->stream, s->sx.g, (const double*)p, s->scratch, (int)row0, (int)(row0 + nrows));
        srcp = s->scratch; s->launches++;
    } else if (kind == KIND_U8) {
        launch_u8_to_double(s->stream, (const unsigned char*)p, s->scratch, n); srcp = s->scratch + (size_t)row0 * width; s->launches++;
    } else if (kind == KIND_I32) {
        launch_i32_to_double(s->stream, (const int*)p, s->scratch, n); srcp = s->scratch + (size_t)row0 * width; s->launches++;
    }
    CU(cudaMemcpyAsync(host, srcp, sizeof(double) * count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return IFL_OK;
}
int ifl_read_field(ifl_solver* s, int32_t field, double* host, size_t count) { return field_io(s, field, host, count, 0, -1, false); }
int ifl_write_field(ifl_solver* s, int32_t field, const double* host, size_t count) {
    return field_io(s, field, const_cast<double*>(host), count, 0, -1, true);
}
int ifl_read_field_rows(ifl_solver* s, int32_t field, int32_t row0, int32_t nrows, double* host, size_t count) {
    if (nrows < 0) INVALID("nrows < 0");
    return field_io(s, field, host, count, row0, nrows, false);
}
int ifl_write_field_rows(ifl_solver* s, int32_t field, int32_t row0, int32_t nrows, const double* host, size_t count) {
    if (nrows < 0) INVALID("nrows < 0");
    return field_io(s, field, const_cast<double*>(host), count, row0, nrows, true);
}
int ifl_slab_rows(const ifl_solver* s, int32_t field, int32_t* row0, int32_t* nrows) {
    if (!s || !row0 || !nrows) INVALID("NULL argument");
    void* p; size_t n, width; int kind;
    int rc = field_ptr(const_cast<ifl_solver*>(s), field, &p, &n, &kind, &width);
    if (rc) return rc;
    int s0, s1;
    slab_range(s->sx.g.nstrips, s->cfg.world_size, s->cfg.rank, &s0, &s1);
    int r0 = 32 * s0 < s->h ? 32 * s0 : s->h, r1 = 32 * s1 < s->h ? 32 * s1 : s->h;
    if (r1 == s->h) r1 = (int)(n / width);           /* v and phi have one row more: it belongs to the last rank */
    *row0 = r0; *nrows = r1 - r0;
    return IFL_OK;
}

int ifl_image_size(const ifl_solver* s, int32_t render_heat, size_t* pixels) {
    if (!s || !pixels) INVALID("NULL argument");
    if (render_heat && !has_heat(s)) INVALID("renderHeat needs a variant with a temperature field (F6, F7, F8)");
    *pixels = (size_t)(render_heat ? 2 * s->w : s->w) * (size_t)s->h;
    return IFL_OK;
}

int ifl_to_image(ifl_solver* s, int32_t render_heat, int32_t* rgb, size_t pixels) {
    if (!s || !rgb) INVALID("NULL argument");
    size_t want = 0;
    int rc = ifl_image_size(s, render_heat, &want);
    if (rc) return rc;
    if (pixels != want) INVALID("pixels does not match ifl_image_size");
    CU(cudaSetDevice(s->cfg.device));
    int* d_rgb = nullptr;
    CU(cudaMalloc(&d_rgb, sizeof(int) * 3 * want));
    launch_to_image(s->stream, s->cfg.variant, s->w, s->h, s->d.src(), has_heat(s) ? s->t.src() : nullptr,
                    has_solids(s) ? s->d.sq.volume : nullptr, has_solids(s) ? s->d.sq.cell : nullptr,
                    s->cfg.t_ambient, render_heat ? 1 : 0, d_rgb);
    s->launches++;
    cudaError_t e = cudaMemcpyAsync(rgb, d_rgb, sizeof(int) * 3 * want, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_rgb);
    CU(e);
    return IFL_OK;
}

int ifl_write_ppm(ifl_solver* s, const char* path, int32_t render_heat) {
    if (!s || !path) INVALID("NULL argument");
    size_t pixels = 0;
    int rc = ifl_image_size(s, render_heat, &pixels);
    if (rc) return rc;
    std::vector<int32_t> rgb(3 * pixels);
    rc = ifl_to_image(s, render_heat, rgb.data(), pixels);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s for writing", path); return IFL_ERR_INVALID; }
    /* F1:346 / F6:1257 header, then F1:353-356: the newline test uses the solver width, whatever the frame width */
    std::string out;
    out.reserve(pixels * 12 + 32);
    char buf[64];
    snprintf(buf, sizeof buf, "P3\n%d %d\n%d\n", render_heat ? 2 * s->w : s->w, s->h, 255);
    out += buf;
    for (size_t i = 0; i < pixels; i++) {
        snprintf(buf, sizeof buf, "%d %d %d ", rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        out += buf;
        if (i % (size_t)s->w == 0) out += "\n";
    }
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    if (!ok) { snprintf(g_err, sizeof g_err, "short write to %s", path); return IFL_ERR_INVALID; }
    return IFL_OK;
}

int ifl_particle_count(ifl_solver* s, int32_t* count) {
    if (!s || !count) INVALID("NULL argument");
    if (!is_flip(s)) { snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant); return IFL_ERR_UNSUPPORTED; }
    CU(cudaSetDevice(s->cfg.device));
    int rc = flip_ensure(s);
    if (rc) return rc;
    *count = s->flip.count;
    return IFL_OK;
}
int ifl_read_particles(ifl_solver* s, double* pos_x, double* pos_y, double* prop_d, double* prop_t, double* prop_u, double* prop_v,
                       size_t capacity) {
    if (!s) INVALID("NULL handle");
    if (!is_flip(s)) { snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant); return IFL_ERR_UNSUPPORTED; }
    CU(cudaSetDevice(s->cfg.device));
    int rc = flip_ensure(s);
    if (rc) return rc;
    FlipCtx& f = s->flip;
    if (capacity < (size_t)f.count) INVALID("capacity < particle count");
    double* host[6] = {pos_x, pos_y, prop_d, prop_t, prop_u, prop_v};
    double* dev[6] = {f.posX, f.posY, f.prop[0], f.prop[1], f.prop[2], f.prop[3]};
    for (int i = 0; i < 6; i++)
        if (host[i]) CU(cudaMemcpyAsync(host[i], dev[i], sizeof(double) * f.count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return IFL_OK;
}
int ifl_write_particles(ifl_solver* s, const double* pos_x, const double* pos_y, const double* prop_d, const double* prop_t,
                        const double* prop_u, const double* prop_v, size_t count) {
    if (!s || !pos_x || !pos_y || !prop_d || !prop_t || !prop_u || !prop_v) INVALID("NULL argument");
    if (!is_flip(s)) { snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant); return IFL_ERR_UNSUPPORTED; }
    FlipCtx& f = s->flip;
    if (count > (size_t)f.max_particles) INVALID("count exceeds w*h*particles_max_per_cell");
    CU(cudaSetDevice(s->cfg.device));
    const double* host[6] = {pos_x, pos_y, prop_d, prop_t, prop_u, prop_v};
    double* dev[6] = {f.posX, f.posY, f.prop[0], f.prop[1], f.prop[2], f.prop[3]};
    for (int i = 0; i < 6; i++) CU(cudaMemcpyAsync(dev[i], host[i], sizeof(double) * count, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    f.count = (int)count;
    f.ready = true;
    s->last.particle_count = f.count;
    return IFL_OK;
}

int ifl_synchronize(ifl_solver* s) {
    if (!s) INVALID("NULL handle");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    return IFL_OK;
}

int ifl_launch_count(const ifl_solver* s, uint64_t* count) {
    if (!s || !count) INVALID("NULL argument");
    *count = s->launches;
    return IFL_OK;
}

/* Diagnostic: run applyPreconditioner(z, r) once on the resident state and return, per strip and
 * sweep, {enter, after first block, mid, end} device timestamps (ns), the number of blocks that had
 * to poll their neighbour, block count, ticket slot and SM id: out[(sweep*nstrips + k)*8 + 0..7]. */
int ifl_wavefront_trace(ifl_solver* s, int64_t* out, size_t capacity) {
    if (!s || !out) INVALID("NULL argument");
    if (!uses_pcg(s)) INVALID("wavefront trace needs a PCG variant");
    CU(cudaSetDevice(s->cfg.device));
    size_t n = (size_t)s->sx.g.nstrips * 32;
    if (capacity < n) INVALID("capacity < 32 * nstrips");
    long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof(long long) * n));
    CU(cudaMemsetAsync(d, 0, sizeof(long long) * n, s->stream));
    s->sx.trace = d;
    const int trc = launch_apply_precon(s->sx, 1);
    s->sx.trace = nullptr;
    if (trc) { cudaFree(d); COMM(trc); }
    CU(cudaMemcpyAsync(out, d, sizeof(long long) * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaFree(d));
    return IFL_OK;
}

int ifl_set_profiling(ifl_solver* s, int32_t enable) {
    if (!s) INVALID("NULL handle");
    CU(cudaSetDevice(s->cfg.device));
    Profiler& p = s->prof;
    if (enable && !p.created) {
        for (int i = 0; i < Profiler::NS; i++) for (int j = 0; j < Profiler::NE; j++) CU(cudaEventCreate(&p.ev[i][j]));
        CU(cudaMalloc(&p.d_done, sizeof(int) * Profiler::NS));
        p.created = true;
    }
    CU(cudaStreamSynchronize(s->stream));
    p.count = 0;
    p.enabled = enable ? 1 : 0;
    return IFL_OK;
}

int ifl_kernel_timing(ifl_solver* s, int32_t which, double* avg_ms, int32_t* samples) {
    if (!s || !avg_ms || !samples) INVALID("NULL argument");
    if (which < 0 || which >= IFL_KERNEL_COUNT) INVALID("unknown kernel id");
    CU(cudaSetDevice(s->cfg.device));
    Profiler& p = s->prof;
    *avg_ms = 0.0; *samples = 0;
    if (!p.created || p.count == 0) return IFL_OK;
    CU(cudaStreamSynchronize(s->stream));
    static int h_done[Profiler::NS];
    CU(cudaMemcpy(h_done, p.d_done, sizeof(int) * p.count, cudaMemcpyDeviceToHost));
    const int want_kind = which == IFL_KERNEL_GS_SWEEP ? 1 : 0;
    const int e0 = which == IFL_KERNEL_GS_SWEEP ? 0 : which;
    double sum = 0.0; int n = 0;
    for (int i = 0; i < p.count; i++) {
        if (p.kind[i] != want_kind || h_done[i]) continue;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p.ev[i][e0], p.ev[i][e0 + 1]));
        sum += ms; n++;
    }
    *samples = n;
    *avg_ms = n ? sum / n : 0.0;
    return IFL_OK;
}

int ifl_last_timing(ifl_solver* s, double out_ms[3]) {
    if (!s || !out_ms) INVALID("NULL argument");
    if (!s->ev_valid) INVALID("no update has run yet");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaEventSynchronize(s->ev[2]));
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    CU(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    CU(cudaEventElapsedTime(&c, s->ev[0], s->ev[2]));
    out_ms[0] = a; out_ms[1] = b; out_ms[2] = c;
    return IFL_OK;
}

} // extern "C"
