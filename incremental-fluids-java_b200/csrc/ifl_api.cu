/*
 * ifl_api.cu -- the C ABI of include/incfluid_b200.h: handle lifetime, device-resident
 * state and the orchestration of FluidSolver.update() as a
 * stream of kernels.  There is no CPU fallback: every entry point that needs the device
 * fails with IFL_ERR_CUDA when CUDA is unusable.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
#define INVALID(msg)                                                                      \
    do { snprintf(g_err, sizeof g_err, "invalid argument: %s", msg); return IFL_ERR_INVALID; } while (0)

struct Quantity {             /* FluidQuantity: _src/_dst ping-pong (F1:105-147) */
    QGeom q{};
    double* buf[2] = {nullptr, nullptr};
    int cur = 0;
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
    Quantity d, u, v;
    SolveCtx sx{};
    double* skew_block = nullptr;      /* one allocation holding the 8 skewed N-vectors */
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
};

static bool uses_pcg(const ifl_solver* s) { return s->cfg.variant >= IFL_F3_CONJUGATE_GRADIENTS; }

extern "C" {

int ifl_abi_version(void) { return IFL_ABI_VERSION; }
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

static int alloc_quantity(Quantity& q, int w, int h, double ox, double oy) {
    q.q = QGeom{w, h, ox, oy};
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc(&q.buf[i], sizeof(double) * q.count()));
        CU(cudaMemset(q.buf[i], 0, sizeof(double) * q.count()));
    }
    q.cur = 0;
    return IFL_OK;
}

void ifl_destroy(ifl_solver* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (Quantity* q : {&s->d, &s->u, &s->v})
        for (int i = 0; i < 2; i++) if (q->buf[i]) cudaFree(q->buf[i]);
    if (s->skew_block) cudaFree(s->skew_block);
    if (s->scratch) cudaFree(s->scratch);
    if (s->sx.pack_f) cudaFree(s->sx.pack_f);
    if (s->sx.pack_b) cudaFree(s->sx.pack_b);
    if (s->sx.bnd_f) cudaFree(s->sx.bnd_f);
    if (s->sx.bnd_b) cudaFree(s->sx.bnd_b);
    if (s->sx.partials) cudaFree(s->sx.partials);
    if (s->sx.stripdot) cudaFree(s->sx.stripdot);
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
    if ((rc = alloc_quantity(s->d, s->w, s->h, 0.5, 0.5))) return rc;       /* F1:268 */
    if ((rc = alloc_quantity(s->u, s->w + 1, s->h, 0.0, 0.5))) return rc;   /* F1:269 */
    if ((rc = alloc_quantity(s->v, s->w, s->h + 1, 0.5, 0.0))) return rc;   /* F1:270 */

    SkewGeom& g = s->sx.g;
    g.w = s->w; g.h = s->h;
    g.nstrips = (s->h + 31) / 32;
    g.tl = ((s->w + 31 + 15) / 16) * 16 + 16;
    g.ss = (long long)g.tl * 32;
    g.total = (long long)g.nstrips * g.ss;
    /* one all-zero strip before and after each vector: the wavefront kernels read the
     * "row above the first" / "row below the last" from there instead of branching */
    const size_t guard = (size_t)g.ss;
    const size_t per = (size_t)g.total + 2 * guard;
    CU(cudaMalloc(&s->skew_block, sizeof(double) * per * 8));
    CU(cudaMemset(s->skew_block, 0, sizeof(double) * per * 8));
    double** vecs[8] = {&s->sx.r, &s->sx.p, &s->sx.z, &s->sx.s, &s->sx.adiag, &s->sx.aplusx, &s->sx.aplusy, &s->sx.precon};
    for (int i = 0; i < 8; i++) *vecs[i] = s->skew_block + per * i + guard;
    const size_t npack = (size_t)g.nstrips * g.tl * 32 * 3;
    CU(cudaMalloc(&s->sx.pack_f, sizeof(double) * npack));
    CU(cudaMalloc(&s->sx.pack_b, sizeof(double) * npack));
    CU(cudaMemset(s->sx.pack_f, 0, sizeof(double) * npack));
    CU(cudaMemset(s->sx.pack_b, 0, sizeof(double) * npack));
    size_t nb = (size_t)g.nstrips * g.w;
    CU(cudaMalloc(&s->sx.bnd_f, sizeof(double) * nb));
    CU(cudaMalloc(&s->sx.bnd_b, sizeof(double) * nb));
    s->sx.npartials = solve_partials_needed(g);
    CU(cudaMalloc(&s->sx.partials, sizeof(double) * s->sx.npartials));
    CU(cudaMalloc(&s->sx.stripdot, sizeof(double) * g.nstrips));
    CU(cudaMalloc(&s->sx.sc, sizeof(SolveScalars)));
    CU(cudaMemset(s->sx.sc, 0, sizeof(SolveScalars)));
    CU(cudaMalloc(&s->d_maxbits, sizeof(unsigned long long)));
    CU(cudaMallocHost(&s->h_sc, sizeof(SolveScalars) * 2));
    memset(s->h_sc, 0, sizeof(SolveScalars) * 2);
    CU(cudaMalloc(&s->scratch, sizeof(double) * (size_t)(s->w + 1) * (s->h + 1)));
    for (auto& e : s->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : s->ev) CU(cudaEventCreate(&e));
    s->sx.tolerance = cfg->tolerance;
    s->sx.mic_tau = cfg->mic_tau;
    s->sx.mic_sigma = cfg->mic_sigma;
    s->sx.sequential = (cfg->flags & IFL_FLAG_SEQUENTIAL_REDUCTIONS) ? 1 : 0;
    s->sx.stream = s->stream;
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
    if (cfg->variant > IFL_F3_CONJUGATE_GRADIENTS) {
        snprintf(g_err, sizeof g_err, "variant %d is not implemented by this build", cfg->variant);
        return IFL_ERR_UNSUPPORTED;
    }
    if (cfg->world_size != 1) {
        snprintf(g_err, sizeof g_err, "slab decomposition (world_size=%d) is not implemented by this build", cfg->world_size);
        return IFL_ERR_UNSUPPORTED;
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
    (void)bodies;
    if (count == 0) return IFL_OK;
    snprintf(g_err, sizeof g_err, "variant %d has no solid bodies", s->cfg.variant);
    return IFL_ERR_UNSUPPORTED;
}

int ifl_add_inflow(ifl_solver* s, double x, double y, double w, double h, double d, double t, double u, double v) {
    if (!s) INVALID("NULL handle");
    (void)t;
    CU(cudaSetDevice(s->cfg.device));
    /* FluidSolver.addInflow F1:298 / F3:385 */
    launch_add_inflow(s->stream, s->cfg.variant, s->d.q, s->d.src(), s->hx, x, y, x + w, y + h, d);
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
        if (pcg) pcg_enqueue_iterations(c, n);
        else gs_enqueue_iterations(c, enq, n, gs_scale);
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
    CU(cudaStreamSynchronize(s->stream));
    const SolveScalars& h = s->h_sc[0];
    s->last.iterations_pressure = h.iterations;
    s->last.residual_pressure = h.max_error;
    s->last.exceeded_budget = h.exceeded ? 1 : 0;
    s->last.nan_residual = h.nan;
    if (st) *st = s->last;
    return h.nan ? IFL_ERR_NAN : IFL_OK;
}

static int stage_build_rhs(ifl_solver* s) {
    launch_build_rhs(s->stream, s->sx.g, s->u.src(), s->v.src(), s->sx.r, s->hx);
    s->launches++;
    return IFL_OK;
}
static int stage_build_matrix(ifl_solver* s, double timestep) {
    double scale = timestep / (s->cfg.density * s->hx * s->hx);             /* F3:436 */
    launch_build_matrix_f3(s->stream, s->sx.g, s->sx.adiag, s->sx.aplusx, s->sx.aplusy, scale);
    s->launches++;
    return IFL_OK;
}
static int stage_project(ifl_solver* s, double timestep) {
    SolveCtx& c = s->sx;
    if (uses_pcg(s)) pcg_enqueue_init(c); else gs_enqueue_begin(c);
    int rc = run_solver_loop(s, timestep);
    if (rc) return rc;
    pcg_enqueue_finish(c, s->cfg.solver_limit);
    return IFL_OK;
}
static int stage_apply_pressure(ifl_solver* s, double timestep) {
    double scale = timestep / (s->cfg.density * s->hx);                      /* F1:433 */
    launch_apply_pressure(s->stream, s->sx.g, s->sx.p, s->u.src(), s->v.src(), scale);
    s->launches++;
    return IFL_OK;
}
static int stage_advect(ifl_solver* s, Quantity& q, double timestep) {
    launch_advect(s->stream, s->cfg.variant, q.q, q.src(), q.dst(), s->u.q, s->u.src(), s->v.q, s->v.src(),
                  timestep, s->hx);
    s->launches++;
    return IFL_OK;
}

static int update_impl(ifl_solver* s, double timestep) {
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
    switch (stage) {
        case IFL_STAGE_BUILD_RHS: rc = stage_build_rhs(s); break;
        case IFL_STAGE_BUILD_MATRIX:
            if (!uses_pcg(s)) goto unsupported;
            rc = stage_build_matrix(s, timestep); break;
        case IFL_STAGE_BUILD_PRECON:
            if (!uses_pcg(s)) goto unsupported;
            launch_build_precon(s->sx); break;
        case IFL_STAGE_PROJECT: rc = stage_project(s, timestep); break;
        case IFL_STAGE_APPLY_PRESSURE: rc = stage_apply_pressure(s, timestep); break;
        case IFL_STAGE_ADVECT_D: rc = stage_advect(s, s->d, timestep); break;
        case IFL_STAGE_ADVECT_U: rc = stage_advect(s, s->u, timestep); break;
        case IFL_STAGE_ADVECT_V: rc = stage_advect(s, s->v, timestep); break;
        case IFL_STAGE_SWAP: s->d.swap(); s->u.swap(); s->v.swap(); break;
        case IFL_STAGE_APPLY_PRECON:
            if (!uses_pcg(s)) goto unsupported;
            launch_apply_precon(s->sx, s->sx.r, s->sx.z, 0); break;
        case IFL_STAGE_MATVEC:
            if (!uses_pcg(s)) goto unsupported;
            launch_matvec(s->sx, s->sx.s, s->sx.z); break;
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

/* field lookup: returns device pointer and element count; skewed != 0 for solve vectors */
static int field_ptr(ifl_solver* s, int32_t field, double** ptr, size_t* count, int* skewed) {
    *skewed = 0;
    size_t n = (size_t)s->w * s->h;
    switch (field) {
        case IFL_FIELD_D: *ptr = s->d.src(); *count = s->d.count(); return IFL_OK;
        case IFL_FIELD_U: *ptr = s->u.src(); *count = s->u.count(); return IFL_OK;
        case IFL_FIELD_V: *ptr = s->v.src(); *count = s->v.count(); return IFL_OK;
        case IFL_FIELD_D_DST: *ptr = s->d.dst(); *count = s->d.count(); return IFL_OK;
        case IFL_FIELD_U_DST: *ptr = s->u.dst(); *count = s->u.count(); return IFL_OK;
        case IFL_FIELD_V_DST: *ptr = s->v.dst(); *count = s->v.count(); return IFL_OK;
        case IFL_FIELD_P: *ptr = s->sx.p; break;
        case IFL_FIELD_R: *ptr = s->sx.r; break;
        case IFL_FIELD_Z: *ptr = s->sx.z; break;
        case IFL_FIELD_S: *ptr = s->sx.s; break;
        case IFL_FIELD_PRECON: *ptr = s->sx.precon; break;
        case IFL_FIELD_ADIAG: *ptr = s->sx.adiag; break;
        case IFL_FIELD_APLUSX: *ptr = s->sx.aplusx; break;
        case IFL_FIELD_APLUSY: *ptr = s->sx.aplusy; break;
        default:
            snprintf(g_err, sizeof g_err, "field %d does not exist for variant %d", field, s->cfg.variant);
            return IFL_ERR_INVALID;
    }
    *count = n;
    *skewed = 1;
    return IFL_OK;
}

int ifl_field_size(const ifl_solver* s, int32_t field, size_t* count) {
    if (!s || !count) INVALID("NULL argument");
    double* p; int sk;
    return field_ptr(const_cast<ifl_solver*>(s), field, &p, count, &sk);
}

int ifl_read_field(ifl_solver* s, int32_t field, double* host, size_t count) {
    if (!s || !host) INVALID("NULL argument");
    CU(cudaSetDevice(s->cfg.device));
    double* p; size_t n; int sk;
    int rc = field_ptr(s, field, &p, &n, &sk);
    if (rc) return rc;
    if (n != count) INVALID("count does not match ifl_field_size");
    if (sk) {
        launch_skew_to_rows(s->stream, s->sx.g, p, s->scratch);
        s->launches++;
        p = s->scratch;
    }
    CU(cudaMemcpyAsync(host, p, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return IFL_OK;
}

int ifl_write_field(ifl_solver* s, int32_t field, const double* host, size_t count) {
    if (!s || !host) INVALID("NULL argument");
    CU(cudaSetDevice(s->cfg.device));
    double* p; size_t n; int sk;
    int rc = field_ptr(s, field, &p, &n, &sk);
    if (rc) return rc;
    if (n != count) INVALID("count does not match ifl_field_size");
    if (sk) {
        CU(cudaMemcpyAsync(s->scratch, host, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        launch_rows_to_skew(s->stream, s->sx.g, s->scratch, p);
        s->launches++;
    } else {
        CU(cudaMemcpyAsync(p, host, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    return IFL_OK;
}

int ifl_particle_count(ifl_solver* s, int32_t* count) {
    if (!s || !count) INVALID("NULL argument");
    snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant);
    return IFL_ERR_UNSUPPORTED;
}
int ifl_read_particles(ifl_solver* s, double*, double*, double*, double*, double*, double*, size_t) {
    if (!s) INVALID("NULL handle");
    snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant);
    return IFL_ERR_UNSUPPORTED;
}
int ifl_write_particles(ifl_solver* s, const double*, const double*, const double*, const double*, const double*,
                        const double*, size_t) {
    if (!s) INVALID("NULL handle");
    snprintf(g_err, sizeof g_err, "variant %d has no particles", s->cfg.variant);
    return IFL_ERR_UNSUPPORTED;
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
    size_t n = (size_t)s->sx.g.nstrips * 16;
    if (capacity < n) INVALID("capacity < 16 * nstrips");
    long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof(long long) * n));
    CU(cudaMemsetAsync(d, 0, sizeof(long long) * n, s->stream));
    s->sx.trace = d;
    launch_apply_precon(s->sx, s->sx.r, s->sx.z, 0);
    s->sx.trace = nullptr;
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
