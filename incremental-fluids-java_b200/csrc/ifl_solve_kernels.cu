/*
 * ifl_solve_kernels.cu -- the pressure solve on the strip-skewed layout (ifl_device.cuh):
 *   - element-wise / stencil kernels of the MIC(0)-PCG loop with fused deterministic
 *     reductions (matrixVectorProduct F3:544 + dotProduct F3:532, scaledAdd F3:570 +
 *     infinityNorm F3:579);
 *   - the three lexicographic recurrences as warp-per-strip anti-diagonal wavefronts:
 *     buildPreconditioner F3:464, applyPreconditioner F3:495 (forward and backward
 *     sweep) and the in-place Gauss-Seidel sweep of project() F1:380.
 *
 * Wavefront scheme.  A warp owns a strip of 32 rows (lane l = row l) and advances one
 * anti-diagonal (one 256-byte t-row of the skewed layout) per step: lane l works on column
 * x = t - l, so its left neighbour is its own previous result and its upper neighbour is
 * lane l-1's previous result (one shuffle).  Any schedule that respects the left/up
 * dependencies performs exactly the reference's per-cell operations, so results are
 * bit-identical to the lexicographic loops.  Strip k+1 consumes the last row of strip k
 * through a "flag-in-data" boundary row: it is pre-filled with a NaN sentinel and the
 * consumer simply re-reads an element until it is no longer the sentinel -- no fences on
 * the critical path.  Strips are handed out through an atomic ticket so that a waiting
 * strip's producer is always resident (no reliance on block scheduling order).
 *
 * Compiled with -fmad=false.
 */
#include <cstdlib>
#include <cooperative_groups.h>
#include "ifl_internal.h"

namespace ifl {

static constexpr int EW_THREADS = 256;
static constexpr int EW_TROWS = 64;    /* t-rows per element-wise block: 2048 elements, 8 per thread */
static constexpr int WF_WARPS = 4;     /* strips (warps) per wavefront block */
static constexpr int WF_U = 8;         /* register prefetch distance in steps */

static dim3 ew_grid(const SkewGeom& g) { return dim3((g.tl + EW_TROWS - 1) / EW_TROWS, g.s1 - g.s0); }
int solve_partials_needed(const SkewGeom& g) { return (int)(((g.tl + EW_TROWS - 1) / EW_TROWS) * g.nstrips); }
static inline void count(SolveCtx& c, int n = 1) { if (c.launches) *c.launches += n; }
static inline SkewGeom full_geom(const SkewGeom& g) { SkewGeom f = g; f.s0 = 0; f.s1 = g.nstrips; return f; }
static inline int single_gpu(const SolveCtx& c) { return c.comm == nullptr; }

/* iteration helper of the element-wise kernels: thread handles 8 elements, t-rows
 * t0 + e*8 + warp, lane = l.  Visits in a fixed order (e ascending). */
#define IFL_EW_LOOP_BEGIN(g)                                                            \
    const int k_ = blockIdx.y + (g).s0;                                                 \
    const long long base_ = (long long)k_ * (g).ss;                                     \
    const int l_ = threadIdx.x & 31;                                                    \
    const int y_ = k_ * 32 + l_;                                                        \
    _Pragma("unroll") for (int e_ = 0; e_ < 8; e_++) {                                  \
        const int t_ = blockIdx.x * EW_TROWS + e_ * 8 + (threadIdx.x >> 5);             \
        const int x_ = t_ - l_;                                                         \
        const long long i_ = base_ + (long long)t_ * 32 + l_;                           \
        const bool ok_ = (t_ < (g).tl) && x_ >= 0 && x_ < (g).w && y_ < (g).h;
#define IFL_EW_LOOP_END }

enum { DOT_NONE = 0, DOT_ZS = 1, DOT_SIGMA_INIT = 2, DOT_SIGMA_NEW = 3 };

__device__ __forceinline__ void apply_dot_result(SolveScalars* sc, int mode, double total) {
    if (mode == DOT_ZS) {                       /* alpha = sigma / dotProduct(z, s)  F3:605 */
        sc->zs = total;
        double a = sc->sigma / total;
        sc->alpha = a;
        sc->neg_alpha = -a;
    } else if (mode == DOT_SIGMA_INIT) {        /* sigma = dotProduct(z, r)  F3:602 */
        sc->sigma = total;
    } else if (mode == DOT_SIGMA_NEW) {         /* sigmaNew; beta = sigmaNew / sigma; sigma = sigmaNew  F3:625-627 */
        sc->beta = total / sc->sigma;
        sc->sigma = total;
    }
}

/* "last block" reduction, fixed order and independent of the slab decomposition: the partials of the blocks of one
 * strip are added in block order (per-strip sum), then the strips are added in global strip order.  On several GPUs
 * the second level runs in k_apply_dot after the per-strip sums of all ranks have been all-gathered. */
__device__ __forceinline__ void finish_partials(double blocksum, double* partials, const SkewGeom& g, double* strip_part,
                                                SolveScalars* sc, int mode, int single_gpu, bool* s_last) {
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    const int nblocks = gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        partials[bid] = blocksum;
        __threadfence();
        unsigned prev = atomicAdd(&sc->ctr_a, 1u);
        *s_last = (prev == (unsigned)nblocks - 1u);
    }
    __syncthreads();
    if (*s_last) {
        __threadfence();
        for (int ks = threadIdx.x; ks < (int)gridDim.y; ks += EW_THREADS) {
            double v = 0.0;
            for (int j = 0; j < (int)gridDim.x; j++) v += ld_cg(partials + ks * gridDim.x + j);
            strip_part[g.s0 + ks] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (single_gpu) {
                double total = 0.0;
                for (int kk = 0; kk < g.nstrips; kk++) total += strip_part[kk];
                apply_dot_result(sc, mode, total);
            }
            sc->ctr_a = 0;
        }
    }
}
/* second level of a distributed dot product: strips in global order (every rank computes the same bits) */
__global__ void k_apply_dot(SkewGeom g, const double* __restrict__ strip_part, SolveScalars* sc, int mode, int check_done) {
    if (check_done && sc->done) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double total = 0.0;
    for (int kk = 0; kk < g.nstrips; kk++) total += strip_part[kk];
    apply_dot_result(sc, mode, total);
}
/* convergence decision from the per-rank maxima (F3:609-619), identical on every rank */
__global__ void k_apply_max(const unsigned long long* __restrict__ rank_max, int world, SolveScalars* sc, double tol, int init) {
    if (!init && sc->done) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long mb = 0;
    for (int r = 0; r < world; r++) mb = rank_max[r] > mb ? rank_max[r] : mb;
    double err = (double)__uint_as_float((unsigned)mb);
    sc->max_error = err;
    sc->iterations = init ? 0 : sc->iterations + 1;
    if (err < tol) sc->done = 1;
    if (err != err) { sc->done = 1; sc->nan = 1; }
}

/* z = A s fused with the partial sums of z.s   (matrixVectorProduct F3:544, dotProduct F3:532) */
__global__ void __launch_bounds__(EW_THREADS)
k_matvec_dot(SkewGeom g, const double* __restrict__ adiag, const double* __restrict__ ax,
             const double* __restrict__ ay, const double* __restrict__ b, double* __restrict__ dst,
             double* partials, double* strip_part, SolveScalars* sc, int dot_mode, int check_done,
             const unsigned char* __restrict__ fl, int single_gpu) {
    if (check_done && sc->done) return;
    __shared__ double smem8[8];
    __shared__ bool s_last;
    double acc = 0.0;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_) {
            double bc = b[i_];
            double tv = adiag[i_] * bc;
            if (x_ > 0) tv += ax[i_ - 32] * b[i_ - 32];
            if (y_ > 0) {
                long long j = (l_ > 0) ? (i_ - 33) : (base_ - g.ss + (long long)(t_ + 31) * 32 + 31);
                tv += ay[j] * b[j];
            }
            if (x_ < g.w - 1) tv += ax[i_] * b[i_ + 32];
            if (y_ < g.h - 1) {
                long long j = (l_ < 31) ? (i_ + 33) : (base_ + g.ss + (long long)(t_ - 31) * 32);
                tv += ay[i_] * b[j];
            }
            dst[i_] = tv;                                   /* matrixVectorProduct is never masked (F6:1011) */
            if (fl == nullptr || fl[i_]) acc += tv * bc;    /* dotProduct: fluid cells only from F6 on (F6:999) */
        }
    IFL_EW_LOOP_END
    if (dot_mode == DOT_NONE) return;
    double bs = block_sum_256(acc, smem8);
    finish_partials(bs, partials, g, strip_part, sc, dot_mode, single_gpu, &s_last);
}

/* p += alpha s ; r -= alpha z ; maxError = infinityNorm(r)   (F3:606-608, scaledAdd F3:570,
 * infinityNorm F3:579).  Also re-arms the boundary rows and tickets of the two sweeps
 * that follow.  The last block decides convergence (F3:609-619). */
__global__ void __launch_bounds__(EW_THREADS)
k_update_pr(SkewGeom g, double* __restrict__ p, const double* __restrict__ s, double* __restrict__ r,
            const double* __restrict__ z, double* __restrict__ bnd_f, double* __restrict__ bnd_b,
            SolveScalars* sc, double tol, const unsigned char* __restrict__ fl,
            unsigned long long* rank_max, int rank, int single_gpu) {
    if (sc->done) return;
    __shared__ unsigned s_max[8];
    __shared__ bool s_last;
    const double alpha = sc->alpha, nalpha = sc->neg_alpha;
    unsigned m = 0;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_ && (fl == nullptr || fl[i_])) {
            p[i_] = p[i_] + s[i_] * alpha;
            double rv = r[i_] + z[i_] * nalpha;
            r[i_] = rv;
            unsigned bits = absf_bits(rv);
            m = bits > m ? bits : m;
        }
        if (l_ == 0 && t_ < g.w) {
            bnd_f[(long long)k_ * g.w + t_] = sentinel();
            bnd_b[(long long)k_ * g.w + t_] = sentinel();
            /* slab decomposition: the rows my neighbours write into (their edge strips) are re-armed by me */
            if (k_ == g.s0 && g.s0 > 0) bnd_f[(long long)(g.s0 - 1) * g.w + t_] = sentinel();
            if (k_ == g.s1 - 1 && g.s1 < g.nstrips) bnd_b[(long long)g.s1 * g.w + t_] = sentinel();
        }
    IFL_EW_LOOP_END
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned other = __shfl_down_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) m = s_max[i] > m ? s_max[i] : m;
        atomicMax(&sc->max_bits, (unsigned long long)m);
        __threadfence();
        unsigned prev = atomicAdd(&sc->ctr_a, 1u);
        s_last = (prev == gridDim.x * gridDim.y - 1u);
        if (s_last) {
            __threadfence();
            unsigned long long mb = atomicMax(&sc->max_bits, 0ull);
            if (single_gpu) {
                double err = (double)__uint_as_float((unsigned)mb);
                sc->max_error = err;
                sc->iterations = sc->iterations + 1;
                if (err < tol) sc->done = 1;
                if (err != err) { sc->done = 1; sc->nan = 1; }
            } else {
                rank_max[rank] = mb;          /* decision after the all-gather, in k_apply_max */
            }
            sc->max_bits = 0ull;
            sc->ctr_a = 0;
            sc->ctr_b = 0;
            sc->ticket_f = 0;
            sc->ticket_b = 0;
        }
    }
}

/* s = z + s * beta   (F3:626) */
__global__ void __launch_bounds__(EW_THREADS)
k_update_s(SkewGeom g, double* __restrict__ s, const double* __restrict__ z, SolveScalars* sc,
           const unsigned char* __restrict__ fl) {
    if (sc->done) return;
    const double beta = sc->beta;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_ && (fl == nullptr || fl[i_])) s[i_] = z[i_] + s[i_] * beta;
    IFL_EW_LOOP_END
}

/* start of project(): s = z (arraycopy F3:596); maxError = infinityNorm(r) (F3:598) */
__global__ void __launch_bounds__(EW_THREADS)
k_init_s(SkewGeom g, double* __restrict__ s, const double* __restrict__ z, const double* __restrict__ r,
         SolveScalars* sc, double tol, const unsigned char* __restrict__ fl,
         unsigned long long* rank_max, int rank, int single_gpu) {
    __shared__ unsigned s_max[8];
    unsigned m = 0;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_) {
            s[i_] = z[i_];                                  /* arraycopy: every cell */
            if (fl == nullptr || fl[i_]) {                  /* infinityNorm: fluid cells only from F6 on (F6:1054) */
                unsigned bits = absf_bits(r[i_]);
                m = bits > m ? bits : m;
            }
        }
    IFL_EW_LOOP_END
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned other = __shfl_down_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) m = s_max[i] > m ? s_max[i] : m;
        atomicMax(&sc->max_bits, (unsigned long long)m);
        __threadfence();
        unsigned prev = atomicAdd(&sc->ctr_a, 1u);
        if (prev == gridDim.x * gridDim.y - 1u) {
            __threadfence();
            unsigned long long mb = atomicMax(&sc->max_bits, 0ull);
            if (single_gpu) {
                double err = (double)__uint_as_float((unsigned)mb);
                sc->max_error = err;
                sc->iterations = 0;
                if (err < tol) sc->done = 1;
                if (err != err) { sc->done = 1; sc->nan = 1; }
            } else {
                rank_max[rank] = mb;
            }
            sc->max_bits = 0ull;
            sc->ctr_a = 0;
        }
    }
}

/* (re)arm scalars, tickets and both boundary-row sets before a solve */
__global__ void k_solve_reset(SkewGeom g, double* bnd_f, double* bnd_b, SolveScalars* sc) {
    long long n = (long long)g.nstrips * g.w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bnd_f[i] = sentinel();
        bnd_b[i] = sentinel();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->sigma = 0.0; sc->alpha = 0.0; sc->neg_alpha = 0.0; sc->beta = 0.0; sc->zs = 0.0;
        sc->max_error = 0.0; sc->max_bits = 0ull;
        sc->iterations = 0; sc->done = 0; sc->nan = 0; sc->exceeded = 0;
        sc->ctr_a = 0; sc->ctr_b = 0; sc->ticket_f = 0; sc->ticket_b = 0;
    }
}

/* re-arm only the boundary rows/tickets (between two stand-alone sweeps) */
__global__ void k_sweep_rearm(SkewGeom g, double* bnd_f, double* bnd_b, SolveScalars* sc) {
    long long n = (long long)g.nstrips * g.w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bnd_f[i] = sentinel();
        bnd_b[i] = sentinel();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->ticket_f = 0; sc->ticket_b = 0; sc->ctr_b = 0; }
}

__global__ void k_solve_finish(SolveScalars* sc, int limit) {
    if (!sc->done && sc->iterations >= limit) sc->exceeded = 1;
}

/* Reductions in the reference's serial order (bit-exact mode, small grids / tests):
 * one thread walks the cells lexicographically.  dotProduct F3:532. */
__global__ void k_dot_sequential(SkewGeom g, const double* __restrict__ a, const double* __restrict__ b,
                                 SolveScalars* sc, int mode, int check_done, const unsigned char* __restrict__ fl) {
    if (check_done && sc->done) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double result = 0.0;
    for (int y = 0; y < g.h; y++) {
        long long i = skew_index(g, 0, y);
        for (int x = 0; x < g.w; x++, i += 32)
            if (fl == nullptr || fl[i]) result += a[i] * b[i];
    }
    apply_dot_result(sc, mode, result);
}

/* ---------------------------------------------------------------------------------
 * Wavefront machinery
 * ------------------------------------------------------------------------------- */

/* Claims the next strip slot for every warp of the block.  Returns the slot or -1. */
__device__ __forceinline__ int claim_slot(unsigned int* ticket, int nstrips) {
    __shared__ unsigned s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    int slot = (int)s_ticket * WF_WARPS + (threadIdx.x >> 5);
    return slot < nstrips ? slot : -1;
}

/* Flag-in-data feed of the neighbouring strip's edge row.  Lane j (< U) owns the value the
 * edge lane needs at step j of the current block; `get(j)` spins (warp-uniformly) until
 * the producer has replaced the sentinel. */
template <int U>
struct EdgeFeed {
    const double* row;   /* neighbour strip's boundary row or nullptr */
    int w;
    double cur, nxt;
    const double* acur;
    const double* anxt;
    __device__ __forceinline__ void init(const double* r, int w_) {
        row = r; w = w_; cur = 0.0; nxt = 0.0; acur = nullptr; anxt = nullptr;
    }
    /* start fetching the values of the block whose step j needs column xin(j) */
    __device__ __forceinline__ void prefetch(int xin_for_my_lane, int lane) {
        anxt = nullptr; nxt = 0.0;
        if (row != nullptr && lane < U && xin_for_my_lane >= 0 && xin_for_my_lane < w) {
            anxt = row + xin_for_my_lane;
            nxt = ld_volatile(anxt);
        }
    }
    __device__ __forceinline__ void advance() { cur = nxt; acur = anxt; }
    __device__ __forceinline__ double get(int j, int lane) {
        double v = __shfl_sync(0xffffffffu, cur, j);
        while (is_sentinel(v)) {
            /* one round trip re-reads every value of the block that is still missing */
            if (lane >= j && lane < U && acur != nullptr && is_sentinel(cur)) cur = ld_volatile(acur);
            v = __shfl_sync(0xffffffffu, cur, j);
        }
        return v;
    }
};


/* Edge feed of the two triangular sweeps.  The lane whose result leaves the strip (E_OUT: lane 31
 * forward, lane 0 backward) is also the lane that READS the neighbouring strip's edge row: it puts
 * that value into the register every lane shuffles, and the shuffle is a rotation, so the lane on
 * the other edge (E_IN) receives the boundary value with the same single shuffle that hands every
 * other lane its neighbour's result -- no second shuffle or select on the dependency chain. */
template <int U>
struct RotFeed {
    const double* row;   /* neighbour strip's boundary row (only the feeder lane dereferences it) */
    int w;
    bool feeder;
    double cur[U], nxt[U];
    __device__ __forceinline__ void init(const double* r, int w_, bool feeder_) {
        row = r; w = w_; feeder = feeder_ && (r != nullptr);
#pragma unroll
        for (int j = 0; j < U; j++) { cur[j] = 0.0; nxt[j] = 0.0; }
    }
    /* x0 = column needed at step 0 of the block, columns advance by dir per step */
    __device__ __forceinline__ void prefetch(int x0, int dir) {
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int x = x0 + dir * j;
            nxt[j] = (feeder && x >= 0 && x < w) ? ld_volatile(row + x) : 0.0;
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int j = 0; j < U; j++) cur[j] = nxt[j];
    }
    /* warp-uniform: are all values of the current block present? */
    __device__ __forceinline__ bool ready() const {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < U; j++) ok = ok && !is_sentinel(cur[j]);
        return __all_sync(0xffffffffu, ok);
    }
    /* spin until the value of step j (column x) has arrived; re-reads all missing later ones too */
    __device__ __forceinline__ void wait(int j0, int x0, int dir) {
        /* bounded: a producer that never shows up (a dead peer rank) must not hang the GPU; after ~2^22 polls the
         * strip gives up and continues with 0.0 -- the result is then wrong, which the callers' checks expose */
        for (unsigned spins = 0;; spins++) {
            bool ok = !is_sentinel(cur[j0]) || spins > (1u << 22);
            if (__all_sync(0xffffffffu, ok)) { if (is_sentinel(cur[j0])) cur[j0] = 0.0; return; }
#pragma unroll
            for (int j = 0; j < U; j++)
                if (j >= j0 && is_sentinel(cur[j])) cur[j] = ld_volatile(row + x0 + dir * j);
        }
    }
};

/* ---- TMA (bulk async copy) staging ------------------------------------------------
 * In the skewed layout the operands a warp needs for U consecutive steps are, per array, ONE
 * contiguous run of U*256 bytes.  Lane 0 streams them into a per-warp shared-memory ring with
 * cp.async.bulk (mbarrier completion), several blocks ahead; the lanes then read their
 * operands with conflict-free 8-byte LDS.  (Register-rotating LDG prefetch does not work
 * here: with 6 scoreboard slots the wait for the oldest load also waits for the youngest.) */
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "IFL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra IFL_DONE;\n"
        "bra IFL_WAIT;\n"
        "IFL_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

static constexpr int SW_U = 16;                     /* steps (t-rows) per staged block of the triangular sweeps */
static constexpr int SW_STAGES = 2;                 /* ring depth: block b+1 streams in while block b is computed */
static constexpr int SW_TILE = SW_U * 32;           /* doubles per array per block (4 KB) */
static constexpr int SW_PACK = 3 * SW_TILE;         /* packed coefficient block: 3 arrays (12 KB) */
static constexpr int SW_FG = 8;                     /* steps per edge-feed group: granularity of the strip hand-over */

/* Packed sweep coefficients.  applyPreconditioner multiplies `aPlusX[i]*precon[i]` and
 * `aPlusY[i]*precon[i]` before it touches dst (F3:500,503,514,517); those products are the same in
 * every PCG iteration, so they are formed once per solve -- same operands, same rounding -- and
 * stored per strip in blocks of SW_U t-rows, [3][SW_U][32] doubles, so that ONE bulk copy fetches
 * every static operand of SW_U wavefront steps:
 *     forward pack : cx = aPlusX*precon (own cell), cy = (aPlusY*precon) of the cell above, precon
 *     backward pack: cx = aPlusX*precon,            cy = aPlusY*precon (own cell),          precon */
__global__ void __launch_bounds__(EW_THREADS)
k_precon_coeffs(SkewGeom g, const double* __restrict__ ax, const double* __restrict__ ay,
                const double* __restrict__ pre, double* __restrict__ pack_f, double* __restrict__ pack_b,
                const unsigned char* __restrict__ fl) {
    const long long pstrip = (long long)(g.tl / SW_U) * SW_PACK;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_) {
            const long long j = (l_ > 0) ? (i_ - 33) : (base_ - g.ss + (long long)(t_ + 31) * 32 + 31);
            double pr = pre[i_];
            double cx = ax[i_] * pr;
            double cy = ay[i_] * pr;
            double cyu = ay[j] * pre[j];
            if (fl != nullptr) {
                /* F4+: non-fluid cells are skipped by applyPreconditioner and never used as neighbours
                 * (F4:783-807).  They are tagged with a NaN precon (the sweep then neither computes nor
                 * stores them and hands +0.0 to its neighbours) and exact +0.0 products. */
                if (!fl[j]) cyu = 0.0;
                if (!fl[i_]) { cx = 0.0; cy = 0.0; pr = __longlong_as_double(0x7ff8000000000b20ll); }
            }
            const long long o = (long long)k_ * pstrip + (long long)(t_ / SW_U) * SW_PACK + (t_ % SW_U) * 32 + l_;
            pack_f[o] = cx; pack_f[o + SW_TILE] = cyu; pack_f[o + 2 * SW_TILE] = pr;
            pack_b[o] = cx; pack_b[o + SW_TILE] = cy;  pack_b[o + 2 * SW_TILE] = pr;
        }
    IFL_EW_LOOP_END
}

/* ---- strip hand-over inside a thread-block cluster ---------------------------------------------------
 * Consecutive strips run on consecutive warps of a cluster (SW_CL blocks x WF_WARPS warps).  A strip hands its
 * edge row to the next one through a small ring in the CONSUMER's shared memory, written with distributed-
 * shared-memory stores (same block: plain shared stores through the same mapped address):
 *   - flag-in-data again: a slot holds the sentinel until the n-th edge value is stored; the consumer polls its
 *     own shared memory (~30 cycles a poll instead of ~300-700 for a global boundary row) and re-arms the slot;
 *   - credit: the consumer publishes how many values it has consumed into the producer's shared memory once per
 *     block; the producer only checks that word (local) before a block, so it can never lap the ring.
 * Only the first strip of a cluster takes its input from a global boundary row (previous cluster / previous GPU)
 * and only the last one writes one. */
static constexpr int SW_CL_MAX = 8;                /* blocks per cluster (portable maximum) */
static constexpr int SW_RB = 128;                  /* ring slots per strip inbox */

/* sentinel test on the high word only (0xFFF8B200: a NaN payload no arithmetic produces) */
__device__ __forceinline__ bool is_sentinel_hi(double v) { return __double2hiint(v) == (int)0xFFF8B200; }

/* Incoming edge values of one group of SW_FG steps.  In ring mode a value lives in slot (producer step & 127), and
 * the producer's step is the consumer's step + 31; in global-row mode the index is the column. */
struct EdgeSrc {
    const double* src;        /* global boundary row, or this warp's shared-memory ring (generic address) */
    double* ring;             /* the ring, writable (re-arm); nullptr in global mode */
    int w;
    bool feeder;              /* this lane reads the edge values */
    int mode;                 /* 0 none, 1 global row, 2 ring */
    __device__ __forceinline__ int index(int x, int sc) const { return mode == 2 ? ((sc + 31) & (SW_RB - 1)) : x; }
    /* x0 / sc0: column and consumer step count of the group's first step; they advance by (dir, +1) per step */
    __device__ __forceinline__ void fetch(double (&v)[SW_FG], int x0, int sc0, int dir) const {
#pragma unroll
        for (int j = 0; j < SW_FG; j++) {
            const int x = x0 + dir * j;
            v[j] = (feeder && x >= 0 && x < w) ? ld_volatile(src + index(x, sc0 + j)) : 0.0;
        }
    }
    __device__ __forceinline__ bool ready(const double (&v)[SW_FG]) const {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < SW_FG; j++) ok = ok && !is_sentinel_hi(v[j]);
        return __all_sync(0xffffffffu, ok);
    }
    /* spin (bounded) until every value of the group has arrived */
    __device__ __forceinline__ void wait(double (&v)[SW_FG], int x0, int sc0, int dir) const {
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < SW_FG; j++) ok = ok && !is_sentinel_hi(v[j]);
            ok = ok || spins > (1u << 24);
            if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
            for (int j = 0; j < SW_FG; j++)
                if (is_sentinel_hi(v[j])) v[j] = ld_volatile(src + index(x0 + dir * j, sc0 + j));
        }
#pragma unroll
        for (int j = 0; j < SW_FG; j++) if (is_sentinel_hi(v[j])) v[j] = 0.0;     /* gave up: dead producer */
    }
    __device__ __forceinline__ void rearm(int x0, int sc0, int dir) const {
        if (mode != 2 || !feeder) return;
#pragma unroll
        for (int j = 0; j < SW_FG; j++) {
            const int x = x0 + dir * j;
            if (x >= 0 && x < w) ring[(sc0 + 31 + j) & (SW_RB - 1)] = sentinel();
        }
    }
};

/* operands of SW_FG steps, loaded from the staged tiles one group ahead of their use: with one warp per
 * scheduler nothing hides a shared-memory load that sits between two dependent FP64 operations */
template <bool WITH_DOT>
struct StepOps {
    double a[SW_FG], cx[SW_FG], cy[SW_FG], pr[SW_FG], rr[SW_FG];
    template <int DIR>
    __device__ __forceinline__ void load(const double* tile_a, const double* tile_c, const double* tile_r, int j0) {
#pragma unroll
        for (int j = 0; j < SW_FG; j++) {
            const int row = (DIR > 0 ? (j0 + j) : SW_U - 1 - (j0 + j)) * 32;
            a[j] = tile_a[row];
            cx[j] = tile_c[row];
            cy[j] = tile_c[SW_TILE + row];
            pr[j] = tile_c[2 * SW_TILE + row];
            rr[j] = WITH_DOT ? tile_r[row] : 0.0;
        }
    }
};

/* SW_FG consecutive steps of one strip; t0 = t of the first step.  pdst: this lane's dst element of the first step
 * (+DIR*32 per step); pout: where the edge lane stores its result of the first step (+ostep per step). */
template <int DIR, bool WITH_DOT, bool CHECKED, bool MASKED>
__device__ __forceinline__ void sweep_steps(const StepOps<WITH_DOT>& op, const double (&edge)[SW_FG], const int t0,
                                            const int lane, const int w, const bool yok, double* pdst, double* pout,
                                            const int ostep, double& dprev, double& cxprev, double& acc) {
    constexpr int E_OUT = DIR > 0 ? 31 : 0;
    const int srclane = (lane - DIR) & 31;          /* rotation: E_IN reads from E_OUT */
    const bool is_out = lane == E_OUT;
#pragma unroll
    for (int j = 0; j < SW_FG; j++) {
        const double pr = op.pr[j];
        const double give = is_out ? edge[j] : dprev;
        const double dn = __shfl_sync(0xffffffffu, give, srclane);
        double v = op.a[j];
        if (DIR > 0) {
            v = v - cxprev * dprev;      /* (aPlusX[i-1]*precon[i-1]) * dst[i-1]   F3:500 */
            v = v - op.cy[j] * dn;       /* (aPlusY[i-w]*precon[i-w]) * dst[i-w]   F3:503 */
            cxprev = op.cx[j];
        } else {
            v = v - op.cx[j] * dprev;    /* (aPlusX[i]*precon[i]) * dst[i+1]       F3:514 */
            v = v - op.cy[j] * dn;       /* (aPlusY[i]*precon[i]) * dst[i+w]       F3:517 */
        }
        double d = v * pr;
        const bool fluid = MASKED ? (pr == pr) : true;      /* NaN precon tags a non-fluid cell (k_precon_coeffs) */
        if (MASKED) d = fluid ? d : 0.0;                     /* skipped cell: dst stays, neighbours see +0.0 */
        if (CHECKED) {
            const int x = t0 + DIR * j - lane;
            const bool active = yok && x >= 0 && x < w;
            if (active && fluid) pdst[DIR * j * 32] = d;
            if (x >= 0 && x < w && is_out) pout[ostep * j] = d;       /* the edge value is published for every column */
        } else {
            /* interior block: only rows y >= h are inactive; they hold +0.0, and storing +0.0 into
             * their (zero) padding cells keeps the padding zero -- no branch needed */
            if (fluid) pdst[DIR * j * 32] = d;
            if (is_out) pout[ostep * j] = d;
        }
        /* inactive lanes hold +0.0 and rr is zero in the padding: adding them is exact */
        if (WITH_DOT) acc += d * op.rr[j];
        dprev = d;
    }
}

/* applyPreconditioner F3:495: DIR=+1 forward substitution (first loop nest), DIR=-1
 * backward substitution (second loop nest).  WITH_DOT (backward only) accumulates
 * dotProduct(dst, rr) per strip (F3:602 / F3:625).
 *
 * The reference guards the two neighbour terms with x>0 / y>0 (x<w-1 / y<h-1).  Here the
 * guards are implicit: every skewed array is zero outside the grid (padding cells, plus one
 * all-zero strip before and after), and inactive lanes provably carry +0.0, so the guarded
 * products are +0.0 and `v - (+0.0)` returns v bit-for-bit (also for v = -0.0).  For the
 * backward sweep aPlusX(w-1,y) and aPlusY(x,h-1) are +0.0 by construction (F3:447,454). */
template <int DIR, bool WITH_DOT, bool MASKED>
__global__ void __launch_bounds__(WF_WARPS * 32)
k_precon_sweep(SkewGeom g, const double* src, double* dst, const double* __restrict__ pack,
               const double* __restrict__ rr, double* bnd, double* bnd_peer, double* strip_part,
               SolveScalars* sc, int dot_mode, int check_done, long long* trace) {
    namespace cg = cooperative_groups;
    constexpr int U = SW_U;
    constexpr int S = SW_STAGES;
    constexpr int NG = SW_U / SW_FG;                                  /* groups per block */
    constexpr int STAGE = (WITH_DOT ? 2 : 1) * SW_TILE + SW_PACK;     /* doubles per stage */
    constexpr int E_IN = DIR > 0 ? 0 : 31;
    constexpr int E_OUT = DIR > 0 ? 31 : 0;
    static_assert(NG == 2, "the group ping-pong below assumes two groups per block");
    extern __shared__ __align__(128) unsigned char wf_smem[];
    __shared__ double s_inbox[WF_WARPS][SW_RB];          /* edge-row rings of this block's strips */
    __shared__ volatile int s_credit[WF_WARPS];          /* steps of MINE the strip after me has consumed */
    __shared__ int s_ticket;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    if (check_done && sc->done) return;                  /* uniform over the whole grid */
    const long long tr_enter = trace ? (long long)globaltimer_ns() : 0;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = lane; i < SW_RB; i += 32) s_inbox[warp][i] = sentinel();
    if (lane == 0) s_credit[warp] = 0;
    /* one ticket per cluster: a waiting cluster's predecessor has started (no reliance on launch order) */
    if (crank == 0 && threadIdx.x == 0) {
        int tk = (int)atomicAdd(DIR > 0 ? &sc->ticket_f : &sc->ticket_b, 1u);
        for (int r = 0; r < csize; r++) *cluster.map_shared_rank(&s_ticket, r) = tk;
    }
    cluster.sync();
    const int nown = g.s1 - g.s0;
    const int cslot = crank * WF_WARPS + warp;                        /* position inside the cluster */
    const int slot = s_ticket * csize * WF_WARPS + cslot;
    const bool live = slot < nown;
    const int k = DIR > 0 ? g.s0 + slot : g.s1 - 1 - slot;
    const int w = g.w;
    double acc = 0.0;
    if (live) {
        const int y = k * 32 + lane;
        const bool yok = y < g.h;
        const int nblk = (w + 31 + U - 1) / U;
        const int tfirst = DIR > 0 ? 0 : nblk * U - 1;
        const bool has_prev = DIR > 0 ? (k > 0) : (k < g.nstrips - 1);

        /* per-warp operand ring: S stages of [src tile | coefficient pack | rr tile], then S mbarriers */
        double* ring = reinterpret_cast<double*>(wf_smem) + (size_t)warp * S * STAGE;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(
            wf_smem + (size_t)WF_WARPS * S * STAGE * sizeof(double)) + warp * S;
        const double* gsrc = src + (long long)k * g.ss;
        const double* grr = WITH_DOT ? rr + (long long)k * g.ss : nullptr;
        const double* gpack = pack + (long long)k * (g.tl / U) * SW_PACK;
        auto issue = [&](int b) {
            const int st = b % S;
            const int gb = DIR > 0 ? b : nblk - 1 - b;
            double* sdst = ring + (size_t)st * STAGE;
            mbar_expect_tx(&bars[st], STAGE * sizeof(double));
            bulk_g2s(sdst, gsrc + (long long)gb * SW_TILE, SW_TILE * sizeof(double), &bars[st]);
            bulk_g2s(sdst + SW_TILE, gpack + (long long)gb * SW_PACK, SW_PACK * sizeof(double), &bars[st]);
            if (WITH_DOT) bulk_g2s(sdst + SW_TILE + SW_PACK, grr + (long long)gb * SW_TILE, SW_TILE * sizeof(double), &bars[st]);
        };
        if (lane == 0) {
            for (int i = 0; i < S; i++) mbar_init(&bars[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int b = 0; b < S && b < nblk; b++) issue(b);
        }
        __syncwarp();

        /* incoming edge: my own ring unless I am the first strip of the cluster (then a global boundary row) */
        const bool in_ring = cslot > 0;
        EdgeSrc feed;
        feed.mode = !has_prev ? 0 : (in_ring ? 2 : 1);
        feed.src = in_ring ? &s_inbox[warp][0] : (has_prev ? bnd + (long long)(k - DIR) * w : nullptr);
        feed.ring = in_ring ? &s_inbox[warp][0] : nullptr;
        feed.w = w;
        feed.feeder = (lane == E_OUT) && feed.mode != 0;
        const bool use_feed = feed.mode != 0;                             /* warp uniform */
        /* outgoing edge: the next strip's ring (mapped shared memory of its block) unless I am the last strip of
         * the cluster / of my slab (then the global boundary row, possibly in the neighbour GPU's memory) */
        const bool out_ring = (cslot < csize * WF_WARPS - 1) && (slot + 1 < nown);
        const bool to_peer = bnd_peer != nullptr && (DIR > 0 ? k == g.s1 - 1 : k == g.s0);
        double* out_base;
        volatile int* credit_of_prev = nullptr;
        if (out_ring) out_base = cluster.map_shared_rank(&s_inbox[(cslot + 1) % WF_WARPS][0], (cslot + 1) / WF_WARPS);
        else out_base = (to_peer ? bnd_peer : bnd) + (long long)k * w;
        if (in_ring) credit_of_prev = cluster.map_shared_rank((int*)&s_credit[(cslot - 1) % WF_WARPS], (cslot - 1) / WF_WARPS);
        /* per step the edge lane's store moves by +1 slot in a ring, by DIR columns in a global row */
        const int ostep = out_ring ? 1 : DIR;

        double eb[2][SW_FG];                                              /* edge values: two groups, ping-pong */
#pragma unroll
        for (int j = 0; j < SW_FG; j++) { eb[0][j] = 0.0; eb[1][j] = 0.0; }
        if (use_feed) feed.fetch(eb[0], tfirst - E_IN, 0, DIR);

        double* pdst = dst + (long long)k * g.ss + lane + (long long)tfirst * 32;
        double dprev = 0.0, cxprev = 0.0;
        long long tr_first = 0, wait_cyc = 0;
        for (int b = 0; b < nblk; b++) {
            const int t0 = tfirst + DIR * b * U;
            const int st = b % S;
            if (trace && b == 1) tr_first = (long long)globaltimer_ns();
            const int tlo = DIR > 0 ? t0 : t0 - (U - 1);
            const bool interior = (tlo - 31 >= 0) && (tlo + U - 1 < w);
            if (out_ring) {
                /* credit: the last step of this block must not lap the consumer's ring */
                for (unsigned spins = 0; (b * U + U - 1) - s_credit[warp] >= SW_RB && spins < (1u << 24); spins++) {}
            }
            const long long wc0 = trace ? clock64() : 0;
            mbar_wait(&bars[st], (unsigned)((b / S) & 1));
            if (trace) wait_cyc += clock64() - wc0;
            const double* tile_a = ring + (size_t)st * STAGE + lane;
            const double* tile_c = tile_a + SW_TILE;
            const double* tile_r = tile_a + SW_TILE + SW_PACK;
            StepOps<WITH_DOT> ops[2];
            ops[0].template load<DIR>(tile_a, tile_c, tile_r, 0);
#pragma unroll
            for (int gi = 0; gi < NG; gi++) {
                const int sc0 = b * U + gi * SW_FG;                        /* my step count at the group's first step */
                const int tg = t0 + DIR * gi * SW_FG;
                const int xin = tg - E_IN;
                if (gi + 1 < NG) ops[(gi + 1) & 1].template load<DIR>(tile_a, tile_c, tile_r, (gi + 1) * SW_FG);
                if (use_feed) {
                    feed.fetch(eb[(gi + 1) & 1], xin + DIR * SW_FG, sc0 + SW_FG, DIR);        /* next group */
                    if (!feed.ready(eb[gi & 1])) feed.wait(eb[gi & 1], xin, sc0, DIR);
                    feed.rearm(xin, sc0, DIR);
                }
                double* gd = pdst + DIR * gi * SW_FG * 32;
                double* po = out_ring ? out_base + (sc0 & (SW_RB - 1)) : out_base + (tg - E_OUT);
                if (interior)
                    sweep_steps<DIR, WITH_DOT, false, MASKED>(ops[gi & 1], eb[gi & 1], tg, lane, w, yok, gd, po, ostep, dprev, cxprev, acc);
                else
                    sweep_steps<DIR, WITH_DOT, true, MASKED>(ops[gi & 1], eb[gi & 1], tg, lane, w, yok, gd, po, ostep, dprev, cxprev, acc);
            }
            pdst += DIR * U * 32;
            __syncwarp();
            if (lane == 0) {
                if (b + S < nblk) issue(b + S);
                /* I have consumed my producer's steps up to (my step count + 31) */
                if (credit_of_prev) *credit_of_prev = (b + 1) * U + 31;
            }
        }
        if (trace && lane == 0) {
            long long* tr = trace + (long long)k * 8;
            tr[0] = tr_enter; tr[1] = tr_first; tr[2] = 0; tr[3] = (long long)globaltimer_ns();
            tr[4] = wait_cyc; tr[5] = nblk; tr[6] = slot; tr[7] = (long long)get_smid();
        }
    }
    if (WITH_DOT && live) {
        double ws = warp_sum(acc);
        if (lane == 0) {
            strip_part[k] = ws;
            __threadfence();
            unsigned prev = atomicAdd(&sc->ctr_b, 1u);
            if (prev == (unsigned)nown - 1u) {
                __threadfence();
                if (nown == g.nstrips) {      /* single GPU: strips in order; otherwise k_apply_dot after the all-gather */
                    double total = 0.0;
                    for (int kk = 0; kk < g.nstrips; kk++) total += ld_cg(strip_part + kk);
                    if (dot_mode != DOT_NONE) apply_dot_result(sc, dot_mode, total);
                }
                sc->ctr_b = 0;
            }
        }
    }
    /* nobody may leave while a neighbour can still store into its shared memory (edge values, credits) */
    cluster.sync();
}

/* buildPreconditioner F3:464 / masked F4:744 (=F5:858, F6:918, F7:951) -- MIC(0) with tau/sigma,
 * lexicographic recurrence on precon.  fl = skewed fluid flags (F4+: non-fluid cells are skipped and
 * never used as neighbours) or nullptr. */
__global__ void __launch_bounds__(WF_WARPS * 32)
k_build_precon(SkewGeom g, const double* __restrict__ adiag, const double* __restrict__ ax,
               const double* __restrict__ ay, double* pre, double* bnd, SolveScalars* sc,
               double tau, double sigma, const unsigned char* __restrict__ fl) {
    constexpr int U = WF_U;
    const int slot = claim_slot(&sc->ticket_f, g.nstrips);
    if (slot < 0) return;
    const int k = slot;
    const int lane = threadIdx.x & 31;
    const int y = k * 32 + lane;
    const bool yok = y < g.h;
    const int w = g.w;
    const long long base = (long long)k * g.ss + lane;
    const int nblk = (w + 31 + U - 1) / U;
    long long upbase = base - 33;
    if (lane == 0) upbase = (long long)(k - 1) * g.ss + 31LL * 32 + 31;
    const bool has_up = y > 0;
    const bool masked = fl != nullptr;

    double AD[U], AX[U], AY[U], AXU[U], AYU[U];
    unsigned char FL[U], FLU[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        int t = j;
        AD[j] = adiag[base + (long long)t * 32];
        AX[j] = ax[base + (long long)t * 32];
        AY[j] = ay[base + (long long)t * 32];
        bool okup = has_up && (t - lane) >= 0 && (t - lane) < w;
        AXU[j] = okup ? ax[upbase + (long long)t * 32] : 0.0;
        AYU[j] = okup ? ay[upbase + (long long)t * 32] : 0.0;
        FL[j] = masked ? fl[base + (long long)t * 32] : 1;
        FLU[j] = (masked && okup) ? fl[upbase + (long long)t * 32] : 1;
    }
    EdgeFeed<U> feed;
    feed.init(k > 0 ? bnd + (long long)(k - 1) * w : nullptr, w);
    feed.prefetch(lane, lane);
    feed.advance();

    double pprev = 0.0, axprev = 0.0, ayprev = 0.0;
    bool flprev = false;
    for (int b = 0; b < nblk; b++) {
        const int t0 = b * U;
        feed.prefetch(t0 + U + lane, lane);
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int t = t0 + j;
            const int x = t - lane;
            double pu = __shfl_up_sync(0xffffffffu, pprev, 1);
            const double bv = feed.get(j, lane);
            if (lane == 0) pu = bv;
            const double ad = AD[j];
            const bool fself = FL[j] != 0;
            double e = ad;
            if (x > 0 && flprev) {
                double px = axprev * pprev;
                double py = ayprev * pprev;
                e -= (px * px + tau * px * py);
            }
            if (y > 0 && FLU[j]) {
                double px = AXU[j] * pu;
                double py = AYU[j] * pu;
                e -= (py * py + tau * px * py);
            }
            if (e < sigma * ad) e = ad;
            const double pc = 1.0 / sqrt(e);
            const bool active = yok && x >= 0 && x < w;
            if (active) {
                if (fself) pre[base + (long long)t * 32] = pc;
                if (lane == 31) bnd[(long long)k * w + x] = fself ? pc : 0.0;   /* flag-in-data: always publish */
            }
            pprev = pc; axprev = AX[j]; ayprev = AY[j]; flprev = fself;
            const int tn = t + U;
            if (tn < g.tl) {
                AD[j] = adiag[base + (long long)tn * 32];
                AX[j] = ax[base + (long long)tn * 32];
                AY[j] = ay[base + (long long)tn * 32];
                bool okup = has_up && (tn - lane) >= 0 && (tn - lane) < w;
                AXU[j] = okup ? ax[upbase + (long long)tn * 32] : 0.0;
                AYU[j] = okup ? ay[upbase + (long long)tn * 32] : 0.0;
                FL[j] = masked ? fl[base + (long long)tn * 32] : 1;
                FLU[j] = (masked && okup) ? fl[upbase + (long long)tn * 32] : 1;
            }
        }
        feed.advance();
    }
}

/* One in-place lexicographic Gauss-Seidel sweep of project()  F1:380 (=F2:395).
 * `parity` selects the boundary-row set of this sweep; the other set is re-armed here
 * for the next sweep.  The last strip to finish decides convergence (F1:420). */
__global__ void __launch_bounds__(WF_WARPS * 32)
k_gs_sweep(SkewGeom g, const double* __restrict__ r, double* p, double* bnd_use, double* bnd_rearm,
           SolveScalars* sc, double scale, double tol) {
    constexpr int U = WF_U;
    if (sc->done) return;
    const int slot = claim_slot(&sc->ticket_f, g.nstrips);
    if (slot < 0) return;
    const int k = slot;
    const int lane = threadIdx.x & 31;
    const int y = k * 32 + lane;
    const bool yok = y < g.h;
    const int w = g.w, h = g.h;
    const long long base = (long long)k * g.ss + lane;
    const int nblk = (w + 31 + U - 1) / U;
    for (int x = lane; x < w; x += 32) bnd_rearm[(long long)k * w + x] = sentinel();

    /* old value of the cell below: (x, y+1) is at i+33, or in the next strip's lane 0 */
    long long dnbase = base + 33;
    if (lane == 31) dnbase = (long long)(k + 1) * g.ss - 31LL * 32;
    const bool has_dn = y < h - 1;

    double R[U], PO[U], PD[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        int t = j;
        R[j] = r[base + (long long)t * 32];
        PO[j] = __ldcg(p + base + (long long)t * 32);
        bool okdn = has_dn && (t - lane) >= 0 && (t - lane) < w;
        PD[j] = okdn ? __ldcg(p + dnbase + (long long)t * 32) : 0.0;
    }
    EdgeFeed<U> feed;
    feed.init(k > 0 ? bnd_use + (long long)(k - 1) * w : nullptr, w);
    feed.prefetch(lane, lane);
    feed.advance();

    double pprev = 0.0;
    unsigned long long mbits = 0ull;
    for (int b = 0; b < nblk; b++) {
        const int t0 = b * U;
        feed.prefetch(t0 + U + lane, lane);
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int t = t0 + j;
            const int x = t - lane;
            double pu = __shfl_up_sync(0xffffffffu, pprev, 1);
            const double bv = feed.get(j, lane);
            if (lane == 0) pu = bv;
            const double pold = PO[j];
            /* fetch p(x+1, y) = own old value of the next step before it is needed; the
             * register of step j+1 still holds it (rotation happens after use) */
            const int tn = t + U;
            const double pright = PO[(j + 1) % U];
            double diag = 0.0, offDiag = 0.0;
            if (x > 0)     { diag += scale; offDiag -= scale * pprev; }
            if (y > 0)     { diag += scale; offDiag -= scale * pu; }
            if (x < w - 1) { diag += scale; offDiag -= scale * pright; }
            if (y < h - 1) { diag += scale; offDiag -= scale * PD[j]; }
            const double newP = (R[j] - offDiag) / diag;
            const bool active = yok && x >= 0 && x < w;
            if (active) {
                /* F1:414: old p rounded through float, difference taken in double */
                double delta = fabs(jfloat(pold) - newP);
                unsigned long long db = (unsigned long long)__double_as_longlong(delta);
                mbits = db > mbits ? db : mbits;
                p[base + (long long)t * 32] = newP;
                if (lane == 31) bnd_use[(long long)k * w + x] = newP;
            }
            pprev = newP;
            if (tn < g.tl) {
                R[j] = r[base + (long long)tn * 32];
                PO[j] = __ldcg(p + base + (long long)tn * 32);
                bool okdn = has_dn && (tn - lane) >= 0 && (tn - lane) < w;
                PD[j] = okdn ? __ldcg(p + dnbase + (long long)tn * 32) : 0.0;
            }
        }
        feed.advance();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, mbits, o);
        mbits = other > mbits ? other : mbits;
    }
    if (lane == 0) {
        atomicMax(&sc->max_bits, mbits);
        __threadfence();
        unsigned prev = atomicAdd(&sc->ctr_b, 1u);
        if (prev == (unsigned)g.nstrips - 1u) {
            __threadfence();
            unsigned long long mb = atomicMax(&sc->max_bits, 0ull);
            double md = __longlong_as_double((long long)mb);
            sc->max_error = md;
            sc->iterations = sc->iterations + 1;
            if (md < tol) sc->done = 1;
            sc->max_bits = 0ull;
            sc->ctr_b = 0;
            sc->ticket_f = 0;
        }
    }
}

/* slab decomposition: boundary rows of a skewed vector <-> packed rows.  rowbuf = [send up | send down | recv up | recv down] */
__global__ void k_pack_rows(SkewGeom g, const double* __restrict__ vec, double* __restrict__ rowbuf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    rowbuf[x] = vec[skew_index(g, x, g.s0 * 32)];                                       /* my first row -> rank-1 */
    int ylast = g.s1 * 32 - 1;
    if (ylast > g.h - 1) ylast = g.h - 1;
    rowbuf[g.w + x] = vec[skew_index(g, x, ylast)];                                     /* my last row -> rank+1 */
}
__global__ void k_unpack_rows(SkewGeom g, double* __restrict__ vec, const double* __restrict__ rowbuf, int has_up, int has_down) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    if (has_up) vec[skew_index(g, x, g.s0 * 32 - 1)] = rowbuf[2 * g.w + x];             /* row above my slab */
    if (has_down) vec[skew_index(g, x, g.s1 * 32)] = rowbuf[3 * g.w + x];               /* row below my slab */
}

/* ---------------------------------------------------------------------------------
 * host-side enqueue helpers
 * ------------------------------------------------------------------------------- */
static inline int wf_blocks(const SkewGeom& g) { return (g.nstrips + WF_WARPS - 1) / WF_WARPS; }
static inline int rearm_blocks(const SkewGeom& g) {
    long long n = (long long)g.nstrips * g.w;
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));
}

void launch_build_precon(SolveCtx& c) {
    k_sweep_rearm<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.sc);
    k_build_precon<<<wf_blocks(c.g), WF_WARPS * 32, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.precon,
                                                                  c.bnd_f, c.sc, c.mic_tau, c.mic_sigma, c.fl);
    { SkewGeom fg = full_geom(c.g);   /* built redundantly for the whole grid on every rank */
      k_precon_coeffs<<<ew_grid(fg), EW_THREADS, 0, c.stream>>>(fg, c.aplusx, c.aplusy, c.precon, c.pack_f, c.pack_b, c.fl); }
    count(c, 3);
}

static size_t sweep_smem(bool with_dot) {
    size_t stage = ((with_dot ? 2 : 1) * SW_TILE + SW_PACK) * sizeof(double);
    return (size_t)WF_WARPS * SW_STAGES * stage + (size_t)WF_WARPS * SW_STAGES * 8;
}
template <int DIR, bool WITH_DOT, bool MASKED>
static void launch_sweep(SolveCtx& c, const double* src, double* dst, const double* pack, const double* rr, double* bnd,
                         int dot_mode, int check_done, long long* trace) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_precon_sweep<DIR, WITH_DOT, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sweep_smem(WITH_DOT));
        attr_done = true;
    }
    const int nown = c.g.s1 - c.g.s0;
    double* peer = nullptr;
    if (c.comm) peer = DIR > 0 ? c.peer_bnd_f_next : c.peer_bnd_b_prev;
    const int nblocks = (nown + WF_WARPS - 1) / WF_WARPS;
    int cl = SW_CL_MAX;
    while (cl > 1 && cl > nblocks) cl >>= 1;                       /* cluster of 8 blocks = 32 consecutive strips */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((nblocks + cl - 1) / cl) * cl);
    cfg.blockDim = dim3(WF_WARPS * 32);
    cfg.dynamicSmemBytes = sweep_smem(WITH_DOT);
    cfg.stream = c.stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_precon_sweep<DIR, WITH_DOT, MASKED>, c.g, src, dst, pack, rr, bnd, peer, c.strip_part, c.sc,
                       dot_mode, check_done, trace);
}

/* dst = M^-1 src: forward then backward sweep.  Boundary rows must be armed. */
static inline void prof_mark(SolveCtx& c, int slot, int e);
static void enqueue_precon(SolveCtx& c, const double* src, double* dst, int dot_mode, int check_done, int ps = -1) {
    const bool masked = c.fl != nullptr;
    if (masked) launch_sweep<1, false, true>(c, src, dst, c.pack_f, nullptr, c.bnd_f, DOT_NONE, check_done, c.trace);
    else        launch_sweep<1, false, false>(c, src, dst, c.pack_f, nullptr, c.bnd_f, DOT_NONE, check_done, c.trace);
    prof_mark(c, ps, 3);
    int fused_mode = c.sequential ? DOT_NONE : dot_mode;
    long long* tr2 = c.trace ? c.trace + 8 * c.g.nstrips : nullptr;
    if (dot_mode != DOT_NONE) {
        if (masked) launch_sweep<-1, true, true>(c, dst, dst, c.pack_b, src, c.bnd_b, fused_mode, check_done, tr2);
        else        launch_sweep<-1, true, false>(c, dst, dst, c.pack_b, src, c.bnd_b, fused_mode, check_done, tr2);
    } else {
        if (masked) launch_sweep<-1, false, true>(c, dst, dst, c.pack_b, nullptr, c.bnd_b, DOT_NONE, check_done, tr2);
        else        launch_sweep<-1, false, false>(c, dst, dst, c.pack_b, nullptr, c.bnd_b, DOT_NONE, check_done, tr2);
    }
    prof_mark(c, ps, 4);
    count(c, 2);
    if (c.sequential && dot_mode != DOT_NONE) {
        k_dot_sequential<<<1, 32, 0, c.stream>>>(c.g, dst, src, c.sc, dot_mode, check_done, c.mask_vector_ops ? c.fl : nullptr);
        count(c);
    }
}

void launch_apply_precon(SolveCtx& c, const double* src, double* dst, int /*with_dot_mode*/) {
    k_sweep_rearm<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.sc);
    count(c);
    enqueue_precon(c, src, dst, DOT_NONE, 0);
}

void launch_matvec(SolveCtx& c, const double* b, double* dst) {
    k_matvec_dot<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, b, dst,
                                                           c.partials, c.strip_part, c.sc, DOT_NONE, 0, nullptr, 1);
    count(c);
}

/* ---- distributed pieces (no-ops on a single GPU) ---- */
__global__ void k_apply_dot(SkewGeom g, const double* __restrict__ strip_part, SolveScalars* sc, int mode, int check_done);
__global__ void k_apply_max(const unsigned long long* __restrict__ rank_max, int world, SolveScalars* sc, double tol, int init);
static void dist_dot(SolveCtx& c, int mode, int check_done) {
    if (single_gpu(c) || mode == DOT_NONE) return;
    const int chunk = c.strips_alloc / c.comm->world;
    comm_all_gather(*c.comm, c.strip_part + (size_t)c.comm->rank * chunk, c.strip_part, sizeof(double) * chunk);
    k_apply_dot<<<1, 32, 0, c.stream>>>(c.g, c.strip_part, c.sc, mode, check_done);
    count(c);
}
static void dist_max(SolveCtx& c, int init) {
    if (single_gpu(c)) return;
    comm_all_gather(*c.comm, c.rank_max + c.comm->rank, c.rank_max, sizeof(unsigned long long));
    k_apply_max<<<1, 32, 0, c.stream>>>(c.rank_max, c.comm->world, c.sc, c.tolerance, init);
    count(c);
}
/* the mat-vec of the next iteration reads s one row above and one row below the slab */
static void dist_halo_rows(SolveCtx& c, double* vec) {
    if (single_gpu(c)) return;
    const int w = c.g.w;
    k_pack_rows<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, vec, c.rowbuf);
    comm_exchange_rows(*c.comm, c.rowbuf, c.rowbuf + 2 * w, c.rowbuf + w, c.rowbuf + 3 * w, (size_t)w);
    k_unpack_rows<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, vec, c.rowbuf, c.comm->rank > 0, c.comm->rank < c.comm->world - 1);
    count(c, 2);
}

/* project() F3:590-602: p = 0; z = M^-1 r; s = z; maxError; sigma = z.r */
void pcg_enqueue_init(SolveCtx& c) {
    cudaMemsetAsync(c.p, 0, sizeof(double) * (size_t)c.g.total, c.stream);
    k_solve_reset<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.sc);
    count(c);
    /* every rank's inboxes must be re-armed before any neighbour starts writing into them */
    if (!single_gpu(c)) comm_all_gather(*c.comm, c.rank_max + c.comm->rank, c.rank_max, sizeof(unsigned long long));
    enqueue_precon(c, c.r, c.z, DOT_SIGMA_INIT, 0);
    dist_dot(c, DOT_SIGMA_INIT, 0);
    k_init_s<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.s, c.z, c.r, c.sc, c.tolerance, c.mask_vector_ops ? c.fl : nullptr,
                                                       c.rank_max, c.comm ? c.comm->rank : 0, single_gpu(c));
    count(c);
    dist_max(c, 1);
    dist_halo_rows(c, c.s);
}

__global__ void k_record_done(const SolveScalars* sc, int* out) { *out = sc->done; }

/* profiling: returns the sample slot to fill for this chunk, or -1 */
static int prof_begin(SolveCtx& c, int kind) {
    Profiler* p = c.prof;
    if (!p || !p->enabled || p->count >= Profiler::NS) return -1;
    int slot = p->count++;
    p->kind[slot] = kind;
    k_record_done<<<1, 1, 0, c.stream>>>(c.sc, p->d_done + slot);
    cudaEventRecord(p->ev[slot][0], c.stream);
    return slot;
}
static inline void prof_mark(SolveCtx& c, int slot, int e) {
    if (slot >= 0) cudaEventRecord(c.prof->ev[slot][e], c.stream);
}

/* n iterations of the loop body F3:603-628; every kernel returns at once after convergence */
void pcg_enqueue_iterations(SolveCtx& c, int n) {
    dim3 eg = ew_grid(c.g);
    const unsigned char* mfl = c.mask_vector_ops ? c.fl : nullptr;
    const int rank = c.comm ? c.comm->rank : 0;
    for (int it = 0; it < n; it++) {
        const int ps = (it == 0 && !c.sequential) ? prof_begin(c, 0) : -1;
        k_matvec_dot<<<eg, EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.s, c.z, c.partials, c.strip_part,
                                                      c.sc, c.sequential ? DOT_NONE : DOT_ZS, 1, mfl, single_gpu(c));
        count(c);
        if (c.sequential) {
            k_dot_sequential<<<1, 32, 0, c.stream>>>(c.g, c.z, c.s, c.sc, DOT_ZS, 1, mfl);
            count(c);
        }
        dist_dot(c, DOT_ZS, 1);
        prof_mark(c, ps, 1);
        k_update_pr<<<eg, EW_THREADS, 0, c.stream>>>(c.g, c.p, c.s, c.r, c.z, c.bnd_f, c.bnd_b, c.sc, c.tolerance, mfl,
                                                     c.rank_max, rank, single_gpu(c));
        count(c);
        dist_max(c, 0);
        prof_mark(c, ps, 2);
        enqueue_precon(c, c.r, c.z, DOT_SIGMA_NEW, 1, ps);
        dist_dot(c, DOT_SIGMA_NEW, 1);
        k_update_s<<<eg, EW_THREADS, 0, c.stream>>>(c.g, c.s, c.z, c.sc, mfl);
        count(c);
        dist_halo_rows(c, c.s);
        prof_mark(c, ps, 5);
    }
}

void pcg_enqueue_finish(SolveCtx& c, int limit) {
    k_solve_finish<<<1, 1, 0, c.stream>>>(c.sc, limit);
    count(c);
    if (!single_gpu(c)) {
        /* every rank needs the whole pressure field for applyPressure: in-place all-gather of the slabs */
        const size_t chunk = (size_t)(c.strips_alloc / c.comm->world) * (size_t)c.g.ss;
        comm_all_gather(*c.comm, c.p + (size_t)c.comm->rank * chunk, c.p, sizeof(double) * chunk);
    }
}

/* Gauss-Seidel project() F1:380: _p is warm-started, so only the control state is reset */
void gs_enqueue_begin(SolveCtx& c) {
    k_solve_reset<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.sc);
    count(c);
}
void gs_enqueue_iterations(SolveCtx& c, int first_iter, int n, double scale) {
    for (int it = first_iter; it < first_iter + n; it++) {
        double* use = (it & 1) ? c.bnd_b : c.bnd_f;
        double* rearm = (it & 1) ? c.bnd_f : c.bnd_b;
        const int ps = (it == first_iter) ? prof_begin(c, 1) : -1;
        k_gs_sweep<<<wf_blocks(c.g), WF_WARPS * 32, 0, c.stream>>>(c.g, c.r, c.p, use, rearm, c.sc, scale, c.tolerance);
        count(c);
        prof_mark(c, ps, 1);
    }
}
void gs_enqueue_finish(SolveCtx& c, int limit) { pcg_enqueue_finish(c, limit); }

} // namespace ifl
