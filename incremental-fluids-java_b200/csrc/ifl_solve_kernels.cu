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
#include <type_traits>
#include <cooperative_groups.h>
#include "ifl_internal.h"

namespace ifl {

static constexpr int EW_THREADS = 256;
static constexpr int EW_TROWS = 64 / IFL_SKB;   /* m-rows per element-wise block: 2048 elements, 8 per thread */
static constexpr int EW_RPP = 8 / IFL_SKB;       /* m-rows covered by the 256 threads in one pass */
static constexpr int WF_WARPS = 4;     /* strips (warps) per wavefront block */
static constexpr int WF_U = 4;         /* register prefetch distance in macro-steps (build / Gauss-Seidel kernels) */
static constexpr int SKB = IFL_SKB;    /* cells per lane per wavefront macro-step */

static dim3 ew_grid(const SkewGeom& g) { return dim3((g.tl + EW_TROWS - 1) / EW_TROWS, g.s1 - g.s0); }
int solve_partials_needed(const SkewGeom& g) { return (int)(((g.tl + EW_TROWS - 1) / EW_TROWS) * g.nstrips); }
static inline void count(SolveCtx& c, int n = 1) { if (c.launches) *c.launches += n; }
static inline int single_gpu(const SolveCtx& c) { return c.comm == nullptr; }

/* iteration helper of the element-wise kernels: a thread handles 8 elements; the 256 threads of a block cover
 * EW_RPP consecutive m-rows per pass (element q = tid % (32*SKB) of its m-row: lane l_ = q / SKB, sub-column
 * j_ = q % SKB), fully coalesced.  Visits in a fixed order (e ascending).  Neighbour indices of element i_:
 * nl_/nr_ = (x-1,y)/(x+1,y), nu_/nd_ = (x,y-1)/(x,y+1) (the latter two may lie in the neighbouring strip). */
#define IFL_EW_LOOP_BEGIN(g) IFL_EW_LOOP_BEGIN_AT(g, blockIdx.x, blockIdx.y)
#define IFL_EW_LOOP_BEGIN_AT(g, bx_, by_)                                               \
    const int k_ = (by_) + (g).s0;                                                      \
    const long long base_ = (long long)k_ * (g).ss;                                     \
    const int q_ = threadIdx.x % (32 * SKB);                                            \
    const int l_ = q_ / SKB;                                                            \
    const int j_ = q_ % SKB;                                                            \
    const int y_ = k_ * 32 + l_;                                                        \
    _Pragma("unroll") for (int e_ = 0; e_ < 8; e_++) {                                  \
        const int t_ = (bx_) * EW_TROWS + e_ * EW_RPP + threadIdx.x / (32 * SKB);       \
        const int x_ = (t_ - l_) * SKB + j_;                                            \
        const long long i_ = base_ + (long long)t_ * (32 * SKB) + q_;                   \
        const bool ok_ = (t_ < (g).tl) && x_ >= 0 && x_ < (g).w && y_ < (g).h;
#define IFL_EW_NEIGHBOURS(g)                                                            \
        const long long nl_ = (j_ > 0) ? (i_ - 1) : (i_ - 32 * SKB + (SKB - 1));        \
        const long long nr_ = (j_ < SKB - 1) ? (i_ + 1) : (i_ + 32 * SKB - (SKB - 1));  \
        const long long nu_ = (l_ > 0) ? (i_ - 33 * SKB)                                \
            : (base_ - (g).ss + ((long long)(t_ + 31) * 32 + 31) * SKB + j_);           \
        const long long nd_ = (l_ < 31) ? (i_ + 33 * SKB)                               \
            : (base_ + (g).ss + ((long long)(t_ - 31) * 32) * SKB + j_);
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

/* Sum of the per-strip partial sums in global strip order by the whole block: the values are staged through shared
 * memory 256 at a time and added by thread 0 in ascending strip order (the result, valid in thread 0, does not depend on
 * how many GPUs or blocks produced the partials). */
__device__ __forceinline__ double ordered_strip_total(const double* strip_part, int nstrips, double* s_buf) {
    double total = 0.0;
    for (int base = 0; base < nstrips; base += EW_THREADS) {
        const int kk = base + (int)threadIdx.x;
        s_buf[threadIdx.x] = kk < nstrips ? ld_cg(strip_part + kk) : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = nstrips - base < EW_THREADS ? nstrips - base : EW_THREADS;
            for (int i = 0; i < n; i++) total += s_buf[i];
        }
        __syncthreads();
    }
    return total;
}

/* Two-level "last block" reduction, fixed order and independent of the slab decomposition: the partials of the blocks
 * of one strip are added in block order by the last block of THAT strip (one counter per strip, so no address is hit by
 * more than gridDim.x atomics), then the last strip to finish adds the strips in global strip order.  On several GPUs
 * the second level runs in k_apply_dot after the per-strip sums of all ranks have been all-gathered.  Returns true in
 * every thread of the one block that finishes last (it owns the kernel's epilogue). */
__device__ __forceinline__ bool finish_partials(double blocksum, double* partials, const SkewGeom& g, double* strip_part,
                                                unsigned* strip_ctr, SolveScalars* sc, int mode, int single_gpu,
                                                int* s_flag, double* s_buf) {
    const int gx = (int)gridDim.x, by = (int)blockIdx.y;
    if (threadIdx.x == 0) {
        int flag = 0;
        partials[(long long)by * gx + blockIdx.x] = blocksum;
        __threadfence();
        const unsigned prev = atomicAdd(&strip_ctr[by], 1u);
        if (prev == (unsigned)gx - 1u) {
            __threadfence();
            double v = 0.0;
            for (int j = 0; j < gx; j++) v += ld_cg(partials + (long long)by * gx + j);
            strip_part[g.s0 + by] = v;
            strip_ctr[by] = 0u;
            __threadfence();
            const unsigned prev2 = atomicAdd(&sc->ctr_a, 1u);
            if (prev2 == gridDim.y - 1u) flag = 1;
        }
        *s_flag = flag;
    }
    __syncthreads();
    if (!*s_flag) return false;
    __threadfence();
    if (single_gpu && mode != DOT_NONE) {
        const double total = ordered_strip_total(strip_part, g.nstrips, s_buf);
        if (threadIdx.x == 0) apply_dot_result(sc, mode, total);
    }
    if (threadIdx.x == 0) sc->ctr_a = 0;
    return true;
}
/* second level of a distributed dot product: strips in global order (every rank computes the same bits) */
__global__ void __launch_bounds__(EW_THREADS)
k_apply_dot(SkewGeom g, const double* __restrict__ strip_part, SolveScalars* sc, int mode, int check_done) {
    __shared__ double s_buf[EW_THREADS];
    if (check_done && sc->done) return;
    const double total = ordered_strip_total(strip_part, g.nstrips, s_buf);
    if (threadIdx.x == 0) apply_dot_result(sc, mode, total);
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

/* matrixVectorProduct F3:544 on its own (IFL_STAGE_MATVEC; the solver loop uses k_step_a).
 * CONSTM (F3, whose matrix buildPressureMatrix F3:435 fills from the grid position alone): the coefficients are
 * re-derived per cell instead of read -- aDiag is `scale` added once per in-grid neighbour (the same additions, in the
 * same order, as k_build_matrix_f3), aPlusX / aPlusY are -scale wherever the guarded term exists -- so the kernel moves
 * 16 instead of 40 bytes per cell and produces the same bits. */
template <bool CONSTM>
__global__ void __launch_bounds__(EW_THREADS)
k_matvec(SkewGeom g, const double* __restrict__ adiag, const double* __restrict__ ax,
         const double* __restrict__ ay, const double* __restrict__ b, double* __restrict__ dst, double cscale) {
    const double ms = -cscale;
    IFL_EW_LOOP_BEGIN(g)
        IFL_EW_NEIGHBOURS(g)
        if (ok_) {
            double bc = b[i_];
            double tv;
            if (CONSTM) {
                double dg = 0.0;
                if (y_ > 0) dg += cscale;
                if (x_ > 0) dg += cscale;
                if (x_ < g.w - 1) dg += cscale;
                if (y_ < g.h - 1) dg += cscale;
                tv = dg * bc;
                if (x_ > 0) tv += ms * b[nl_];
                if (y_ > 0) tv += ms * b[nu_];
                if (x_ < g.w - 1) tv += ms * b[nr_];
                if (y_ < g.h - 1) tv += ms * b[nd_];
            } else {
                tv = adiag[i_] * bc;
                if (x_ > 0) tv += ax[nl_] * b[nl_];
                if (y_ > 0) tv += ay[nu_] * b[nu_];
                if (x_ < g.w - 1) tv += ax[i_] * b[nr_];
                if (y_ < g.h - 1) tv += ay[i_] * b[nd_];
            }
            dst[i_] = tv;                                   /* matrixVectorProduct is never masked (F6:1011) */
        }
    IFL_EW_LOOP_END
}

/* First kernel of a PCG iteration: everything between two global reductions that is element-wise or a 5-point
 * stencil, in ONE pass over the vectors --
 *     p += alpha*s            scaledAdd(_p, _p, _s, alpha)   F3:607 of the PREVIOUS iteration (deferred: nothing in
 *                             the loop reads p, so the update only has to happen before s is overwritten; the last
 *                             iteration's update is applied by k_update_p)
 *     s  = z + beta*s         scaledAdd(_s, _z, _s, beta)    F3:626 of the previous iteration
 *     z  = A s                matrixVectorProduct            F3:604
 *     z.s                     dotProduct                     F3:605  -> alpha = sigma / (z.s)
 * The mat-vec needs the new s of the four neighbours: it is recomputed from z and the old s with the very same two
 * operations (z[n] + s[n]*beta), so every s value has the bits the reference's separate loop produces.  Because cells
 * read their neighbours' OLD z and s, the kernel cannot work in place: z and s live in two buffers each and the kernel
 * reads generation `zpar`, writes generation `zpar ^ 1` and flips `zpar` (a device scalar: the host enqueues blindly).
 * Masked scaledAdd (F6:1041) leaves s unchanged outside the fluid -- copied here; the mat-vec is never masked.
 * In the first iteration (iterations == 0) s already is the copy of z made by k_init_s (F3:596): s is carried over
 * unchanged and nothing is added to p.
 * Moves 48 B/cell for F3 (R z,s,p  W s,p,z), against 16 + 40 for separate mat-vec and update kernels. */
/* resident blocks per SM the register allocation aims at.  F3 (matrix re-derived from the grid position): 5 (48 registers, a
 * few spilled values) measured 4 % faster than 4 (64 registers), 6 is 35 % slower.  The array-reading instance (F4+):
 * 4 (64 registers) measured 11 % faster than unconstrained (96 registers, 2 blocks) and 16 % faster than 5 */
template <bool CONSTM>
__global__ void __launch_bounds__(EW_THREADS, CONSTM ? 5 : 4)
k_step_a(SkewGeom g, const double* __restrict__ adiag, const double* __restrict__ ax, const double* __restrict__ ay,
         double* z0, double* z1, double* s0, double* s1, double* __restrict__ p,
         double* partials, double* strip_part, unsigned* strip_ctr, SolveScalars* sc, int dot_mode,
         const unsigned char* __restrict__ fl, int single_gpu, double cscale) {
    static_assert(SKB == 2 && EW_THREADS == 256, "a thread owns the two cells of one lane and m-row");
    constexpr int TR = EW_TROWS;                    /* m-rows per block */
    constexpr int RS = 32 * SKB;                    /* doubles per m-row */
    if (sc->done) return;
    /* the new s of the block's m-rows plus one m-row before and after (left / right / up / down neighbours inside the
     * strip are at +-1 m-row), and of the two rows of the neighbouring strips that lanes 0 and 31 look at */
    __shared__ __align__(16) double S[(TR + 2) * RS];
    __shared__ __align__(16) double UP[TR * SKB], DN[TR * SKB];
    __shared__ double s_buf[EW_THREADS];
    __shared__ double smem8[8];
    __shared__ int s_flag;
    const int zp = sc->zpar;
    const bool first = sc->iterations == 0;
    const double* __restrict__ zin = zp ? z1 : z0;
    const double* __restrict__ sin_ = zp ? s1 : s0;
    double* __restrict__ zout = zp ? z0 : z1;
    double* __restrict__ sout = zp ? s0 : s1;
    const double beta = sc->beta, alpha = sc->alpha;
    const double ms = -cscale;
    const int k = (int)blockIdx.y + g.s0;
    const int t0 = (int)blockIdx.x * TR;
    const long long base = (long long)k * g.ss;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int y = k * 32 + lane;
    const bool yok = y < g.h;
    const int w = g.w;
    /* s of cell i after F3:626 (masked F6:1041): the reference's two operations, or the untouched old value */
    auto s_new = [&](long long i) -> double {
        const double sv = sin_[i];
        if (first || (fl != nullptr && !fl[i])) return sv;
        return zin[i] + sv * beta;
    };

    /* ---- phase 1: p and s of the own cells (registers + shared memory), s of the halo cells (shared memory) ---- */
    double sown[TR / 8][SKB];
#pragma unroll
    for (int ps = 0; ps < TR / 8; ps++) {
        const int r = ps * 8 + wrp;                 /* m-row inside the block */
        const int t = t0 + r;
        const int x0 = (t - lane) * SKB;
        const long long i = base + (long long)t * RS + lane * SKB;
        double sn[SKB] = {0.0, 0.0};
        if (t < g.tl) {
            const double2 so = *reinterpret_cast<const double2*>(sin_ + i);
            double zv[SKB] = {0.0, 0.0}, pv[SKB] = {0.0, 0.0};
            if (!first) {
                const double2 zz = *reinterpret_cast<const double2*>(zin + i);
                const double2 pp = *reinterpret_cast<const double2*>(p + i);
                zv[0] = zz.x; zv[1] = zz.y; pv[0] = pp.x; pv[1] = pp.y;
            }
            const double sold[SKB] = {so.x, so.y};
            bool ok[SKB];
#pragma unroll
            for (int j = 0; j < SKB; j++) {
                ok[j] = yok && x0 + j >= 0 && x0 + j < w;
                const bool act = ok[j] && !first && (fl == nullptr || fl[i + j]);
                sn[j] = sold[j];
                if (act) {
                    pv[j] = pv[j] + sold[j] * alpha;        /* scaledAdd(_p, _p, _s, alpha)   F3:607 */
                    sn[j] = zv[j] + sold[j] * beta;         /* scaledAdd(_s, _z, _s, beta)    F3:626 */
                }
            }
            if (ok[0] && ok[1]) {
                *reinterpret_cast<double2*>(sout + i) = make_double2(sn[0], sn[1]);
                if (!first) *reinterpret_cast<double2*>(p + i) = make_double2(pv[0], pv[1]);
            } else {
#pragma unroll
                for (int j = 0; j < SKB; j++) if (ok[j]) { sout[i + j] = sn[j]; if (!first) p[i + j] = pv[j]; }
            }
        }
        sown[ps][0] = sn[0]; sown[ps][1] = sn[1];
        *reinterpret_cast<double2*>(&S[(r + 1) * RS + lane * SKB]) = make_double2(sn[0], sn[1]);
    }
    if (wrp < 2) {                                          /* m-rows t0 - 1 and t0 + TR of this strip */
        const int t = wrp == 0 ? t0 - 1 : t0 + TR;
        double sn[SKB] = {0.0, 0.0};
        if (t >= 0 && t < g.tl) {
            const long long i = base + (long long)t * RS + lane * SKB;
            sn[0] = s_new(i); sn[1] = s_new(i + 1);
        }
        *reinterpret_cast<double2*>(&S[(wrp == 0 ? 0 : TR + 1) * RS + lane * SKB]) = make_double2(sn[0], sn[1]);
    } else if (wrp < 4) {                                   /* the cells above lane 0 / below lane 31 */
        const int r = lane;                                 /* TR == 32 m-rows, one per thread */
        const int t = t0 + r;
        double sn[SKB] = {0.0, 0.0};
        if (wrp == 2) {                                     /* (x, 32k - 1): strip k-1, lane 31, m-row t + 31 */
            if (t + 31 < g.tl) {
                const long long i = base - g.ss + ((long long)(t + 31) * 32 + 31) * SKB;
                sn[0] = s_new(i); sn[1] = s_new(i + 1);
            }
            UP[r * SKB] = sn[0]; UP[r * SKB + 1] = sn[1];
        } else {                                            /* (x, 32k + 32): strip k+1, lane 0, m-row t - 31 */
            if (t - 31 >= 0 && t - 31 < g.tl) {
                const long long i = base + g.ss + (long long)(t - 31) * 32 * SKB;
                sn[0] = s_new(i); sn[1] = s_new(i + 1);
            }
            DN[r * SKB] = sn[0]; DN[r * SKB + 1] = sn[1];
        }
    }
    __syncthreads();

    /* ---- phase 2: z = A s (matrixVectorProduct F3:544-565, never masked) and the partial sums of z.s ---- */
    double acc = 0.0;
#pragma unroll
    for (int ps = 0; ps < TR / 8; ps++) {
        const int r = ps * 8 + wrp;
        const int t = t0 + r;
        if (t >= g.tl) continue;
        const int x0 = (t - lane) * SKB;
        const long long i = base + (long long)t * RS + lane * SKB;
        const double* Sr = &S[(r + 1) * RS + lane * SKB];
        const double sl0 = Sr[-RS + 1];                     /* (x0 - 1, y): previous m-row, second cell */
        const double sr1 = Sr[RS];                          /* (x0 + 2, y): next m-row, first cell */
        const double2 up = lane > 0 ? *reinterpret_cast<const double2*>(Sr - RS - SKB)
                                    : *reinterpret_cast<const double2*>(&UP[r * SKB]);
        const double2 dn = lane < 31 ? *reinterpret_cast<const double2*>(Sr + RS + SKB)
                                     : *reinterpret_cast<const double2*>(&DN[r * SKB]);
        const double sc_[SKB] = {sown[ps][0], sown[ps][1]};
        const double sl[SKB] = {sl0, sc_[0]}, sr[SKB] = {sc_[1], sr1};
        const double su[SKB] = {up.x, up.y}, sd[SKB] = {dn.x, dn.y};
        double cd[SKB], cl[SKB], cu[SKB], cr[SKB], cdn[SKB];
        if (!CONSTM) {
            const double2 ad = *reinterpret_cast<const double2*>(adiag + i);
            const double2 axc = *reinterpret_cast<const double2*>(ax + i);
            const double2 ayc = *reinterpret_cast<const double2*>(ay + i);
            const long long iu = lane > 0 ? i - RS - SKB : base - g.ss + ((long long)(t + 31) * 32 + 31) * SKB;
            const double2 ayu = *reinterpret_cast<const double2*>(ay + iu);
            cd[0] = ad.x; cd[1] = ad.y;
            cl[0] = ax[i - RS + 1]; cl[1] = axc.x;          /* aPlusX of the left neighbour */
            cu[0] = ayu.x; cu[1] = ayu.y;                   /* aPlusY of the upper neighbour */
            cr[0] = axc.x; cr[1] = axc.y;
            cdn[0] = ayc.x; cdn[1] = ayc.y;
        }
        double tv[SKB];
        bool ok[SKB];
#pragma unroll
        for (int j = 0; j < SKB; j++) {
            const int x = x0 + j;
            ok[j] = yok && x >= 0 && x < w;
            if (CONSTM) {
                double dg = 0.0;
                if (y > 0) dg += cscale;
                if (x > 0) dg += cscale;
                if (x < w - 1) dg += cscale;
                if (y < g.h - 1) dg += cscale;
                double v = dg * sc_[j];
                if (x > 0) v += ms * sl[j];
                if (y > 0) v += ms * su[j];
                if (x < w - 1) v += ms * sr[j];
                if (y < g.h - 1) v += ms * sd[j];
                tv[j] = v;
            } else {
                double v = cd[j] * sc_[j];
                if (x > 0) v += cl[j] * sl[j];
                if (y > 0) v += cu[j] * su[j];
                if (x < w - 1) v += cr[j] * sr[j];
                if (y < g.h - 1) v += cdn[j] * sd[j];
                tv[j] = v;
            }
            /* dotProduct: fluid cells only from F6 on (F6:999) */
            if (ok[j] && (fl == nullptr || fl[i + j])) acc += tv[j] * sc_[j];
        }
        if (ok[0] && ok[1]) *reinterpret_cast<double2*>(zout + i) = make_double2(tv[0], tv[1]);
        else {
#pragma unroll
            for (int j = 0; j < SKB; j++) if (ok[j]) zout[i + j] = tv[j];
        }
    }
    const double bs = block_sum_256(acc, smem8);
    if (finish_partials(bs, partials, g, strip_part, strip_ctr, sc, dot_mode, single_gpu, &s_flag, s_buf) && threadIdx.x == 0) {
        sc->zpar = zp ^ 1;
        sc->ctr_b = 0; sc->ctr_c = 0;
        sc->ticket_f = 0; sc->ticket_b = 0;
    }
}

/* r -= alpha z ; maxError = infinityNorm(r)   (F3:608, scaledAdd F3:570, infinityNorm F3:579) as a kernel of its own:
 * used instead of the r update fused into the forward sweep when IFL_FLAG_SEPARATE_R_UPDATE is set (same operations,
 * same bits).  The last block decides convergence (F3:609-619). */
__global__ void __launch_bounds__(EW_THREADS)
k_update_pr(SkewGeom g, double* __restrict__ r, const double* z0, const double* z1,
            SolveScalars* sc, double tol, const unsigned char* __restrict__ fl,
            unsigned long long* rank_max, int rank, int single_gpu) {
    if (sc->done) return;
    __shared__ unsigned s_max[8];
    const double* __restrict__ z = sc->zpar ? z1 : z0;
    const double nalpha = sc->neg_alpha;
    unsigned m = 0;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_ && (fl == nullptr || fl[i_])) {
            double rv = r[i_] + z[i_] * nalpha;
            r[i_] = rv;
            unsigned bits = absf_bits(rv);
            m = bits > m ? bits : m;
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
                sc->iterations = sc->iterations + 1;
                if (err < tol) sc->done = 1;
                if (err != err) { sc->done = 1; sc->nan = 1; }
            } else {
                rank_max[rank] = mb;          /* decision after the all-gather, in k_apply_max */
            }
            sc->max_bits = 0ull;
            sc->ctr_a = 0;
            sc->p_pending = 1;
        }
    }
}

/* scaledAdd(_p, _p, _s, alpha) F3:607 (masked F6:1041) of the LAST iteration of a solve: inside the loop the update of p
 * rides on the next iteration's k_step_a, which never runs for the iteration that converges or hits the limit. */
__global__ void __launch_bounds__(EW_THREADS)
k_update_p(SkewGeom g, double* __restrict__ p, const double* s0, const double* s1, const SolveScalars* sc,
           const unsigned char* __restrict__ fl) {
    if (!sc->p_pending) return;
    const double alpha = sc->alpha;
    const double* __restrict__ s = sc->zpar ? s1 : s0;
    IFL_EW_LOOP_BEGIN(g)
        if (ok_ && (fl == nullptr || fl[i_])) p[i_] = p[i_] + s[i_] * alpha;
    IFL_EW_LOOP_END
}

/* start of project(): s = z (arraycopy F3:596); maxError = infinityNorm(r) (F3:598) */
__global__ void __launch_bounds__(EW_THREADS)
k_init_s(SkewGeom g, double* s0, double* s1, const double* z0, const double* z1, const double* __restrict__ r,
         SolveScalars* sc, double tol, const unsigned char* __restrict__ fl,
         unsigned long long* rank_max, int rank, int single_gpu) {
    __shared__ unsigned s_max[8];
    double* __restrict__ s = sc->zpar ? s1 : s0;
    const double* __restrict__ z = sc->zpar ? z1 : z0;
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

/* (re)arm scalars, tickets, the boundary rows of the MIC(0) build / Gauss-Seidel wavefronts ([nstrips][w]) and the
 * global inbox rows of the triangular sweeps ([nstrips][inbox_len]) before a solve.  (Inside a solve the sweeps re-arm
 * what they consume.) */
__device__ __forceinline__ void arm_rows(SkewGeom g, double* bnd_f, double* bnd_b, double* inbox_f, double* inbox_b, int inbox_len) {
    const long long n = (long long)g.nstrips * g.w, ni = (long long)g.nstrips * inbox_len;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) { bnd_f[i] = sentinel(); bnd_b[i] = sentinel(); }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ni; i += stride) { inbox_f[i] = sentinel(); inbox_b[i] = sentinel(); }
}
__global__ void k_solve_reset(SkewGeom g, double* bnd_f, double* bnd_b, double* inbox_f, double* inbox_b, int inbox_len, SolveScalars* sc) {
    arm_rows(g, bnd_f, bnd_b, inbox_f, inbox_b, inbox_len);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->sigma = 0.0; sc->alpha = 0.0; sc->neg_alpha = 0.0; sc->beta = 0.0; sc->zs = 0.0;
        sc->max_error = 0.0; sc->max_bits = 0ull;
        sc->iterations = 0; sc->done = 0; sc->nan = 0; sc->exceeded = 0; sc->p_pending = 0;
        sc->ctr_a = 0; sc->ctr_b = 0; sc->ctr_c = 0; sc->ticket_f = 0; sc->ticket_b = 0;
        /* zpar is NOT reset: which z / s generation is current survives from solve to solve, like the reference's
         * persistent _z / _s arrays (their stale values on masked cells) */
    }
}

/* re-arm only the rows / tickets (before a stand-alone wavefront) */
__global__ void k_sweep_rearm(SkewGeom g, double* bnd_f, double* bnd_b, double* inbox_f, double* inbox_b, int inbox_len, SolveScalars* sc) {
    arm_rows(g, bnd_f, bnd_b, inbox_f, inbox_b, inbox_len);
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->ticket_f = 0; sc->ticket_b = 0; sc->ctr_b = 0; sc->ctr_c = 0; }
}

__global__ void k_solve_finish(SolveScalars* sc, int limit) {
    sc->p_pending = 0;
    if (!sc->done && sc->iterations >= limit) sc->exceeded = 1;
}

/* Reductions in the reference's serial order (bit-exact mode): dotProduct F3:532, `result += a[i]*b[i]` with i ascending.
 * One block: the 256 threads form the products of 256 consecutive cells of a row (the loads run in parallel), thread 0
 * then adds them strictly in the reference's order.  A cell the masked dotProduct skips (F6:1002) contributes +0.0, which
 * leaves the running sum unchanged bit for bit (the sum starts at +0.0 and can never become -0.0).
 * a = current z; b = b_fixed (r) or the current s. */
__global__ void __launch_bounds__(EW_THREADS)
k_dot_sequential(SkewGeom g, const double* z0, const double* z1, const double* b_fixed,
                 const double* s0, const double* s1,
                 SolveScalars* sc, int mode, int check_done, const unsigned char* __restrict__ fl) {
    if (check_done && sc->done) return;
    __shared__ double s_prod[2][EW_THREADS];
    const double* __restrict__ a = sc->zpar ? z1 : z0;
    const double* __restrict__ b = b_fixed != nullptr ? b_fixed : (sc->zpar ? s1 : s0);
    const long long n = (long long)g.w * g.h;
    double result = 0.0;
    int buf = 0;
    for (long long base = 0; base < n; base += EW_THREADS, buf ^= 1) {
        const long long c = base + threadIdx.x;
        double pr = 0.0;
        if (c < n) {
            const int y = (int)(c / g.w), x = (int)(c - (long long)y * g.w);
            const long long i = skew_index(g, x, y);
            if (fl == nullptr || fl[i]) pr = a[i] * b[i];
        }
        s_prod[buf][threadIdx.x] = pr;
        __syncthreads();                       /* two buffers: thread 0 may still be adding the previous chunk */
        if (threadIdx.x == 0) {
            const int m = n - base < EW_THREADS ? (int)(n - base) : EW_THREADS;
            for (int i = 0; i < m; i++) result += s_prod[buf][i];
        }
    }
    if (threadIdx.x == 0) apply_dot_result(sc, mode, result);
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

/* Flag-in-data feed of the neighbouring strip's edge row.  Lane q (< U*SKB) owns the value the edge lane needs
 * for sub-column q % SKB of macro-step q / SKB of the current block; `get(q)` spins (warp-uniformly) until the
 * producer has replaced the sentinel. */
template <int U>
struct EdgeFeed {
    static_assert(U * SKB <= 32, "one value per lane");
    const double* row;   /* neighbour strip's boundary row or nullptr */
    int w;
    double cur, nxt;
    const double* acur;
    const double* anxt;
    __device__ __forceinline__ void init(const double* r, int w_) {
        row = r; w = w_; cur = 0.0; nxt = 0.0; acur = nullptr; anxt = nullptr;
    }
    /* start fetching the values of the block whose first macro-step needs column block c0 */
    __device__ __forceinline__ void prefetch(int c0, int lane) {
        anxt = nullptr; nxt = 0.0;
        const int x = c0 * SKB + lane;           /* (c0 + lane / SKB) * SKB + lane % SKB */
        if (row != nullptr && lane < U * SKB && x >= 0 && x < w) {
            anxt = row + x;
            nxt = ld_volatile(anxt);
        }
    }
    __device__ __forceinline__ void advance() { cur = nxt; acur = anxt; }
    __device__ __forceinline__ double get(int q, int lane) {
        double v = __shfl_sync(0xffffffffu, cur, q);
        while (is_sentinel(v)) {
            /* one round trip re-reads every value of the block that is still missing */
            if (lane >= q && lane < U * SKB && acur != nullptr && is_sentinel(cur)) cur = ld_volatile(acur);
            v = __shfl_sync(0xffffffffu, cur, q);
        }
        return v;
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

#ifndef IFL_SW_U
#define IFL_SW_U 8
#endif
#ifndef IFL_SW_STAGES
#define IFL_SW_STAGES 5
#endif
#ifndef IFL_SW_WARPS
#define IFL_SW_WARPS 2
#endif
static constexpr int SW_U = IFL_SW_U;               /* macro-steps (m-rows) per staged block of the triangular sweeps */
static constexpr int SW_STAGES = IFL_SW_STAGES;     /* shared-memory ring depth per strip */
static constexpr int SW_WARPS = IFL_SW_WARPS;       /* strips per block of the sweep kernel: few, so that each strip's operand ring
                                                     * (~100 KB) covers the latency of the bulk copies; a wavefront over a w-wide
                                                     * grid only ever has ~ (w/SKB + 31) / 40 strips in flight at a time */
static constexpr int SW_TILE = SW_U * 32 * SKB;     /* doubles per array per block (2 KB) */
static constexpr int SW_PACK = 3 * SW_TILE;         /* packed coefficient block: 3 arrays (6 KB) */
static constexpr int SW_THREADS = (2 * SW_WARPS + 1) * 32;   /* strip warps, their loader warps, one edge-mover warp */
static constexpr int SW_FG = 8;                     /* macro-steps per block of the sweep loop (edge-feed group, hand-over granularity) */
static_assert(SW_FG % SW_U == 0 && SW_FG % 2 == 0, "a block is a whole, even number of staged tiles");

/* Packed sweep coefficients.  applyPreconditioner multiplies `aPlusX[i]*precon[i]` and
 * `aPlusY[i]*precon[i]` before it touches dst (F3:500,503,514,517); those products are the same in
 * every PCG iteration, so they are formed once per solve -- same operands, same rounding -- and
 * stored per strip in blocks of SW_U m-rows, [3][SW_U][32][SKB] doubles, so that ONE bulk copy fetches
 * every static operand of SW_U wavefront macro-steps:
 *     forward pack : cx = aPlusX*precon (own cell), cy = (aPlusY*precon) of the cell above, precon
 *     backward pack: cx = aPlusX*precon,            cy = aPlusY*precon (own cell),          precon */
__global__ void __launch_bounds__(EW_THREADS)
k_precon_coeffs(SkewGeom g, const double* __restrict__ ax, const double* __restrict__ ay,
                const double* __restrict__ pre, double* __restrict__ pack_f, double* __restrict__ pack_b,
                const unsigned char* __restrict__ fl) {
    const long long pstrip = (long long)(g.tl / SW_U) * SW_PACK;
    IFL_EW_LOOP_BEGIN(g)
        IFL_EW_NEIGHBOURS(g)
        if (ok_) {
            double pr = pre[i_];
            double cx = ax[i_] * pr;
            double cy = ay[i_] * pr;
            double cyu = ay[nu_] * pre[nu_];
            if (fl != nullptr) {
                /* F4+: non-fluid cells are skipped by applyPreconditioner and never used as neighbours
                 * (F4:783-807).  They are tagged with a NaN precon (the sweep then neither computes nor
                 * stores them and hands +0.0 to its neighbours) and exact +0.0 products. */
                if (!fl[nu_]) cyu = 0.0;
                if (!fl[i_]) { cx = 0.0; cy = 0.0; pr = __longlong_as_double(0x7ff8000000000b20ll); }
            }
            const long long o = (long long)k_ * pstrip + (long long)(t_ / SW_U) * SW_PACK + (t_ % SW_U) * (32 * SKB) + q_;
            pack_f[o] = cx; pack_f[o + SW_TILE] = cyu; pack_f[o + 2 * SW_TILE] = pr;
            pack_b[o] = cx; pack_b[o + SW_TILE] = cy;  pack_b[o + 2 * SW_TILE] = pr;
            (void)nl_; (void)nr_; (void)nd_;
        }
    IFL_EW_LOOP_END
}

/* ---- strip hand-over inside a thread-block cluster ---------------------------------------------------
 * Consecutive strips run on consecutive warps of a cluster (SW_CL blocks x SW_WARPS warps).  A strip hands its
 * edge row to the next one through a small ring in the CONSUMER's shared memory, written with distributed-
 * shared-memory stores (same block: plain shared stores through the same mapped address):
 *   - flag-in-data again: a slot holds the sentinel until the n-th edge value is stored; the consumer polls its
 *     own shared memory (~30 cycles a poll instead of ~300-700 for a global boundary row) and re-arms the slot;
 *   - credit: the consumer publishes how many values it has consumed into the producer's shared memory once per
 *     block; the producer only checks that word (local) before a block, so it can never lap the ring.
 * Only the first strip of a cluster takes its input from a global boundary row (previous cluster / previous GPU)
 * and only the last one writes one. */
#ifndef IFL_SW_CL
#define IFL_SW_CL 8
#endif
static constexpr int SW_CL_MAX = IFL_SW_CL;        /* blocks per cluster (8 = portable maximum) */
static constexpr int SW_RB = 128;                  /* ring slots per strip inbox */

/* sentinel test on the high word only (0xFFF8B200: a NaN payload no arithmetic produces) */
__device__ __forceinline__ bool is_sentinel_hi(double v) { return __double2hiint(v) == (int)0xFFF8B200; }

__device__ __forceinline__ void ld_volatile_v2(const double* p, double& a, double& b) {      /* generic address */
    asm volatile("ld.volatile.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_v2(double* p, double a, double b) {                      /* generic address */
    asm volatile("st.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
/* predicated 16-byte store without a branch (written as `if`, a store under a lane predicate becomes a divergent branch) */
__device__ __forceinline__ void st_v2_if(bool p, double* addr, double a, double b) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.v2.f64 [%1], {%2, %3};\n}" ::"r"((unsigned)p), "l"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_int_if(bool p, volatile int* addr, int v) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.volatile.s32 [%1], %2;\n}" ::"r"((unsigned)p), "l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void lds_cells(unsigned addr, double* out) {
    static_assert(SKB == 2, "one 16-byte shared-memory load per array and macro-step");
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[1]) : "r"(addr));
}
__device__ __forceinline__ void mbar_arrive_if(bool p, unsigned bar) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q mbarrier.arrive.shared::cta.b64 _, [%1];\n}" ::"r"((unsigned)p), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "IFL_WAITU:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra IFL_DONEU;\n"
        "bra IFL_WAITU;\n"
        "IFL_DONEU:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
/* keeps a value in a register: the compiler would otherwise re-derive shared-memory addresses from the thread id on
 * every use, which costs issue slots on a one-warp-per-scheduler critical path */
__device__ __forceinline__ unsigned pin_u32(unsigned v) { asm volatile("mov.u32 %0, %0;" : "+r"(v)); return v; }

/* applyPreconditioner F3:495: DIR=+1 forward substitution (first loop nest), DIR=-1 backward substitution (second loop
 * nest).  WITH_DOT (backward only) accumulates dotProduct(dst, rr) per strip (F3:602 / F3:625).
 *
 * A warp owns a strip of 32 rows (lane l = row l) and at macro-step m works on the two cells of column block c = m - l
 * (one 512-byte m-row of the skewed layout): the left neighbour of its first cell is its own previous result (a
 * register), the upper neighbours are lane l-1's results of its previous macro-step (two 64-bit shuffles issued at the
 * start of the macro-step, off the dependency chain).  Any schedule that respects the left/up dependencies performs
 * exactly the reference's per-cell operations, so the results are bit-identical to the lexicographic loops.  A macro-step
 * is nothing but the recurrence (10 FP64 operations), the shuffles, five 16-byte operand loads for the NEXT macro-step and
 * two 16-byte stores.
 *
 * The reference guards the two neighbour terms with x>0 / y>0 (x<w-1 / y<h-1).  Here the guards are implicit: every
 * skewed array is zero outside the grid (padding cells, plus one all-zero strip before and after), inactive lanes
 * provably carry +0.0, so the guarded products are +0.0 and `v - (+0.0)` returns v bit-for-bit (also for v = -0.0); for
 * the backward sweep aPlusX(w-1,y) and aPlusY(x,h-1) are +0.0 by construction (F3:447,454).  Storing the +0.0 of an
 * inactive lane into its padding cell keeps the padding zero.  Hence ONE code path, without a per-cell predicate, for
 * every block of every strip.
 *
 * Strips run on consecutive warps of a thread-block cluster (SW_CL blocks x SW_WARPS strip warps; every strip warp has
 * a loader warp that streams its operands with bulk copies into a shared-memory ring, full/empty mbarrier pair per
 * stage).  Strip k+1 consumes the edge row of strip k -- the results of its last lane -- through an INBOX indexed by
 * the producer's macro-step + 1: consumer macro-step c needs the pair of producer macro-step c + 31.  The inbox every
 * strip warp reads is a 128-pair ring in its own shared memory.  Inside a cluster the producer writes it directly
 * (distributed-shared-memory stores, one 16-byte store per macro-step).  Between clusters and between GPUs the producer
 * stores into the consumer's row of global memory (CUDA-IPC-mapped peer memory over NVLink on the next GPU), and the
 * last warp of the consumer's block, the EDGE MOVER, polls that row and forwards every pair into the ring -- so a strip
 * warp never waits for an L2 round trip (a strip fed from global memory directly ran at 1200 instead of 720 us at
 * 16384^2 and paced every strip after it).
 * Flag-in-data: a slot holds a NaN sentinel until the value is stored; the consumer takes the 8 pairs of a block of 8
 * macro-steps at the block's start (8 broadcast loads, repeated while a pair is missing), re-arms them with one store
 * and returns a credit word to the producer (or the mover), which keeps it from lapping the ring.  A strip that has
 * caught up with its producer advances block by block in lockstep with it, 31 + 8 macro-steps behind.  No fences on
 * the critical path, every spin is bounded.  Clusters take an atomic ticket, so a waiting cluster's predecessor has
 * always started (deadlock-free without a cooperative launch).
 *
 * FUSE (forward sweep inside the PCG loop only): the loader warp of every strip also performs, tile by tile and before
 * the strip warp sees the operands,
 *     r += (-alpha) * z            scaledAdd(_r, _r, _z, -alpha)  F3:608   (FUSE = 2: fluid cells only, F6:1041)
 *     maxError = infinityNorm(r)   F3:609 / F6:1054, and the convergence decision F3:611-619 when the last strip is done,
 * on the r and z (= A s) tiles the bulk copies have just staged, writes the new r back to global memory and hands the
 * tile to the strip warp, which runs the triangular solve on it: one pass over r and z instead of a separate 24 B/cell
 * kernel, on a warp that is otherwise idle.  z is swept in place (its tile is read before the strip warp overwrites it).
 * If the solve turns out to have converged the sweep's output is simply not used (the reference returns before
 * applyPreconditioner; z is dead until the next solve overwrites it). */
template <int DIR, bool WITH_DOT, bool MASKED, int FUSE>
__global__ void __launch_bounds__(SW_THREADS)
k_precon_sweep(SkewGeom g, const double* src_fixed, double* dst0, double* dst1, const double* __restrict__ pack,
               const double* __restrict__ rr, double* inbox, double* inbox_peer, int inbox_len, double* strip_part,
               SolveScalars* sc, int dot_mode, int check_done, long long* trace,
               double tol, unsigned long long* rank_max, int rank, int single_gpu) {
    static_assert(FUSE == 0 || (DIR > 0 && !WITH_DOT), "the r update rides on the forward sweep");
    static_assert(SKB == 2 && SW_U == SW_FG && SW_FG == 8, "a block of the strip loop is one staged tile of 8 macro-steps");
    namespace cg = cooperative_groups;
    constexpr int U = SW_U;                                           /* macro-steps per staged tile */
    constexpr int UB = SW_FG;                                         /* macro-steps per block of the loop */
    constexpr int S = SW_STAGES;
    constexpr int STAGE = ((WITH_DOT || FUSE) ? 2 : 1) * SW_TILE + SW_PACK;   /* doubles per stage */
#ifndef IFL_SW_LAG
#define IFL_SW_LAG 2
#endif
#ifndef IFL_SW_G
#define IFL_SW_G 8                                                     /* edge pairs taken at once (divides 8) */
#endif
    constexpr int LAG = IFL_SW_LAG;                                   /* FUSE: tiles between a bulk copy and its r update (< S - 1) */
    static_assert(LAG < SW_STAGES - 1, "the strip warp needs tile b+1 before it releases tile b");
    constexpr int E_IN = DIR > 0 ? 0 : 31;
    constexpr int E_OUT = DIR > 0 ? 31 : 0;
    extern __shared__ __align__(128) unsigned char wf_smem[];
    __shared__ __align__(16) double s_inbox[SW_WARPS][SW_RB * SKB];   /* edge-row rings of this block's strips */
    __shared__ volatile int s_credit[SW_WARPS];          /* macro-steps of MINE the strip after me has consumed */
    __shared__ volatile int s_mcredit;                   /* the same, for the edge mover of this block's strip 0 */
    __shared__ __align__(16) double s_dummy[SW_WARPS][32][SKB];           /* where a non-fluid cell (MASKED) stores */
    __shared__ int s_ticket;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    if (check_done && sc->done) return;                  /* uniform over the whole grid */
    double* const dst = sc->zpar ? dst1 : dst0;          /* the current generation of z */
    const double* const src = src_fixed != nullptr ? src_fixed : dst;
    const long long tr_enter = trace ? (long long)globaltimer_ns() : 0;
    const int lane = threadIdx.x & 31;
    const bool mover = threadIdx.x >= 2 * SW_WARPS * 32;              /* last warp: feeds strip warp 0 from its global inbox row */
    const bool loader = !mover && threadIdx.x >= SW_WARPS * 32;
    const int warp = mover ? 0 : (threadIdx.x >> 5) - (loader ? SW_WARPS : 0);    /* strip warp index (own, or the one served) */
    /* per-strip operand ring: S stages of [src tile | coefficient pack | rr tile]; then per strip S full + S empty mbarriers */
    double* ring = reinterpret_cast<double*>(wf_smem) + (size_t)warp * S * STAGE;
    unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(
        wf_smem + (size_t)SW_WARPS * S * STAGE * sizeof(double)) + warp * 3 * S;
    unsigned long long* bar_empty = bar_full + S;
    unsigned long long* bar_ready = bar_empty + S;       /* FUSE: the loader warp has updated r in the staged tile */
    if (mover && lane == 0) s_mcredit = 0;
    if (!loader && !mover) {
        s_dummy[warp][lane][0] = 0.0; s_dummy[warp][lane][SKB - 1] = 0.0;
        for (int i = lane; i < SW_RB * SKB; i += 32) s_inbox[warp][i] = sentinel();
        if (lane == 0) {
            s_credit[warp] = 0;
            for (int i = 0; i < S; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); mbar_init(&bar_ready[i], 1); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    /* one ticket per cluster: a waiting cluster's predecessor has started (no reliance on launch order) */
    if (crank == 0 && threadIdx.x == 0) {
        int tk = (int)atomicAdd(DIR > 0 ? &sc->ticket_f : &sc->ticket_b, 1u);
        for (int r = 0; r < csize; r++) *cluster.map_shared_rank(&s_ticket, r) = tk;
    }
    cluster.sync();
    const int nown = g.s1 - g.s0;
    const int cslot = crank * SW_WARPS + warp;                        /* position inside the cluster */
    const int slot = s_ticket * csize * SW_WARPS + cslot;
    const bool live = slot < nown;
    const int k = DIR > 0 ? g.s0 + slot : g.s1 - 1 - slot;
    const int w = g.w;
    const int cw = (w + SKB - 1) / SKB;                               /* column blocks per row */
    const int nblk = (cw + 31 + UB - 1) / UB;
    const int ntile = nblk * (UB / U);
    double acc = 0.0;
    if (live && loader) {
        /* running pointers / shared addresses: this loop has to issue a tile faster than the strip warp consumes one */
        const long long dtile = DIR > 0 ? SW_TILE : -SW_TILE;
        const long long dpack = DIR > 0 ? SW_PACK : -SW_PACK;
        const long long first = DIR > 0 ? 0 : ntile - 1;
        const double* gsrc = src + (long long)k * g.ss + first * SW_TILE;
        /* second streamed vector: rr (dot product fused into the backward sweep) or z = A s (FUSE: the r update) */
        const double* grr = WITH_DOT ? rr + (long long)k * g.ss + first * SW_TILE
                                     : (FUSE ? dst + (long long)k * g.ss + first * SW_TILE : nullptr);
        const double* gpack = pack + (long long)k * (g.tl / U) * SW_PACK + first * SW_PACK;
        const unsigned ring0 = smem_u32(ring), full0 = smem_u32(bar_full), empty0 = smem_u32(bar_empty);
        unsigned sdst = ring0, fb = full0, eb = empty0;
        int st = 0;
        unsigned ph = 1;                             /* parity of the completion that frees a stage; first lap: free */
        auto issue_tile = [&](const int b) {         /* lane 0 only */
            if (b >= S) mbar_wait_u32(eb, ph);
            if (b == ntile && DIR < 0) { gsrc -= dtile; gpack -= dpack; if (WITH_DOT) grr -= dtile; }   /* stay inside */
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((unsigned)(STAGE * sizeof(double))) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sdst), "l"(gsrc), "r"((unsigned)(SW_TILE * sizeof(double))), "r"(fb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sdst + (unsigned)(SW_TILE * sizeof(double))), "l"(gpack), "r"((unsigned)(SW_PACK * sizeof(double))), "r"(fb) : "memory");
            if (WITH_DOT || FUSE)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sdst + (unsigned)((SW_TILE + SW_PACK) * sizeof(double))), "l"(grr), "r"((unsigned)(SW_TILE * sizeof(double))), "r"(fb) : "memory");
            gsrc += dtile; gpack += dpack; if (WITH_DOT || FUSE) grr += dtile;
            sdst += (unsigned)(STAGE * sizeof(double)); fb += 8u; eb += 8u;
            if (++st == S) { st = 0; ph ^= 1u; sdst = ring0; fb = full0; eb = empty0; }
        };
        if (!FUSE) {
            /* one tile more than the strip has: the strip warp prefetches across every tile boundary without a test */
            if (lane == 0) for (int b = 0; b <= ntile; b++) issue_tile(b);
        } else {
            /* Every round the whole warp first updates r in the tile whose bulk copies were issued LAG rounds ago (lane l
             * owns the SKB cells of row l in every m-row, like the strip warp) and releases it to the strip warp through
             * the `ready` barrier; lane 0 then issues the next tile's copies as soon as the strip warp has handed a stage
             * back.  The update therefore runs 2-3 tiles ahead of the strip warp and never on its critical path.  Only
             * cells inside the grid are touched: the padding of r and z stays zero whatever alpha is. */
            const int y = k * 32 + lane;
            const bool yok = y < g.h;
            const double nalpha = sc->neg_alpha;
            double* const rout = const_cast<double*>(src) + (long long)k * g.ss + lane * SKB;
            /* running max of |r| as a 64-bit pattern (non-negative doubles order like unsigned integers, NaN sorts above
             * +inf); abs((float) .) of F3:582 is monotonic, so it is applied once, to the maximum (the double -> float
             * conversion is a slow instruction: 16 of them per tile cost 0.4 ms per sweep at 16384^2) */
            unsigned long long mx0 = 0ull, mx1 = 0ull;
#ifdef IFL_SW_TRACE_WAITS
            long long hw_full = 0, hw_proc = 0, hw_issue = 0;
            const long long hw_start = clock64();
#define IFL_HW(acc, stmt) { const long long c0_ = clock64(); stmt; acc += clock64() - c0_; }
#else
#define IFL_HW(acc, stmt) { stmt; }
#endif
            for (int b = 0; b <= ntile + LAG; b++) {
                const int pb = b - LAG;
                if (pb >= 0) {
                    const int sp = pb % S;
                    IFL_HW(hw_full, mbar_wait_u32(full0 + 8u * (unsigned)sp, (unsigned)(pb / S) & 1u))
#ifdef IFL_SW_TRACE_WAITS
                    const long long hw_p0 = clock64();
#endif
                    if (pb < ntile) {
                        double* const tl_r = ring + (size_t)sp * STAGE + lane * SKB;      /* r tile, this lane's cells */
                        const double* const tl_z = tl_r + SW_TILE + SW_PACK;              /* z = A s */
                        const double* const tl_p = tl_r + 3 * SW_TILE;                    /* precon (NaN: not fluid) */
                        const int tbase = pb * U;
                        /* every cell of the tile inside the grid: plain 16-byte accesses, no predicates */
                        const bool inner = yok && (tbase - lane) >= 0 && ((tbase + U - 1 - lane) * SKB + SKB - 1) < w;
                        if (__all_sync(0xffffffffu, inner) && FUSE == 1) {
#pragma unroll
                            for (int e = 0; e < U; e++) {
                                const double2 rv = *reinterpret_cast<const double2*>(tl_r + e * 32 * SKB);
                                const double2 zv = *reinterpret_cast<const double2*>(tl_z + e * 32 * SKB);
                                const double n0 = rv.x + zv.x * nalpha, n1 = rv.y + zv.y * nalpha;    /* a[i] + b[i]*s  F3:573 */
                                const unsigned long long b0 = (unsigned long long)__double_as_longlong(n0) & 0x7fffffffffffffffull;
                                const unsigned long long b1 = (unsigned long long)__double_as_longlong(n1) & 0x7fffffffffffffffull;
                                mx0 = b0 > mx0 ? b0 : mx0;
                                mx1 = b1 > mx1 ? b1 : mx1;
                                *reinterpret_cast<double2*>(tl_r + e * 32 * SKB) = make_double2(n0, n1);
                                *reinterpret_cast<double2*>(rout + (long long)(tbase + e) * (32 * SKB)) = make_double2(n0, n1);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < U; e++) {
                                const int t = tbase + e;
                                const int x0 = (t - lane) * SKB;
                                const double2 rv2 = *reinterpret_cast<const double2*>(tl_r + e * 32 * SKB);
                                const double2 zv2 = *reinterpret_cast<const double2*>(tl_z + e * 32 * SKB);
                                double rv[SKB] = {rv2.x, rv2.y};
                                const double zv[SKB] = {zv2.x, zv2.y};
                                double pv[SKB] = {0.0, 0.0};
                                if (FUSE == 2) { const double2 q = *reinterpret_cast<const double2*>(tl_p + e * 32 * SKB); pv[0] = q.x; pv[1] = q.y; }
                                bool ok[SKB];
#pragma unroll
                                for (int j = 0; j < SKB; j++) {
                                    ok[j] = yok && x0 + j >= 0 && x0 + j < w;
                                    if (FUSE == 2) ok[j] = ok[j] && (pv[j] == pv[j]);     /* NaN precon tags a non-fluid cell */
                                    const double nv = rv[j] + zv[j] * nalpha;
                                    if (ok[j]) {
                                        rv[j] = nv;
                                        const unsigned long long bits = (unsigned long long)__double_as_longlong(nv) & 0x7fffffffffffffffull;
                                        mx0 = bits > mx0 ? bits : mx0;
                                    }
                                }
                                *reinterpret_cast<double2*>(tl_r + e * 32 * SKB) = make_double2(rv[0], rv[1]);
                                double* gp = rout + (long long)t * (32 * SKB);
                                if (ok[0] && ok[1]) *reinterpret_cast<double2*>(gp) = make_double2(rv[0], rv[1]);
                                else {
#pragma unroll
                                    for (int j = 0; j < SKB; j++) if (ok[j]) gp[j] = rv[j];
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_ready[sp]);
#ifdef IFL_SW_TRACE_WAITS
                    hw_proc += clock64() - hw_p0;
#endif
                }
                IFL_HW(hw_issue, if (b <= ntile && lane == 0) issue_tile(b); __syncwarp())
            }
#ifdef IFL_SW_TRACE_WAITS
            if (trace && lane == 0) {
                long long* tr = trace + (long long)k * 16;
                tr[9] = hw_full; tr[10] = hw_proc; tr[11] = hw_issue; tr[12] = clock64() - hw_start; tr[13] = ntile;
            }
#endif
            /* maxError and the convergence decision (F3:609-619) once every strip of this rank has updated its r */
            unsigned mloc = absf_bits(__longlong_as_double((long long)(mx0 > mx1 ? mx0 : mx1)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned other = __shfl_down_sync(0xffffffffu, mloc, o);
                mloc = other > mloc ? other : mloc;
            }
            if (lane == 0) {
                atomicMax(&sc->max_bits, (unsigned long long)mloc);
                __threadfence();
                const unsigned prev = atomicAdd(&sc->ctr_c, 1u);
                if (prev == (unsigned)nown - 1u) {
                    __threadfence();
                    const unsigned long long mb = atomicMax(&sc->max_bits, 0ull);
                    if (single_gpu) {
                        const double err = (double)__uint_as_float((unsigned)mb);
                        sc->max_error = err;
                        sc->iterations = sc->iterations + 1;
                        if (err < tol) sc->done = 1;
                        if (err != err) { sc->done = 1; sc->nan = 1; }
                    } else {
                        rank_max[rank] = mb;          /* decision after the all-gather, in k_apply_max */
                    }
                    sc->max_bits = 0ull;
                    sc->ctr_c = 0;
                    sc->p_pending = 1;                /* r carries this iteration's alpha, p does not yet */
                }
            }
        }
    } else if (live && mover) {
        /* The first strip of a cluster (or of a GPU's slab) has its predecessor in another cluster: that one stores its edge
         * pairs into this strip's row of global memory.  This warp moves them into the strip's shared-memory ring, so that
         * every strip warp consumes its edge values the same way, from shared memory: lane i polls pair slots i, i+32, ...
         * (one L2 round trip per poll, 32 polls in flight), hands a pair over as soon as it is there and the ring has room
         * (the strip warp's credit), and re-arms the slot in the row for the next sweep. */
        const bool has_prev = DIR > 0 ? (k > 0) : (k < g.nstrips - 1);
        if (crank == 0 && has_prev) {
            const int last = nblk * UB + 7;                          /* the strip consumes slots 25 .. nsteps + 7 */
            double* const row = inbox + (long long)k * inbox_len;
            const unsigned ring_addr = smem_u32(&s_inbox[0][0]);
            const double sv = sentinel();
            int slot = 25 + lane;
            for (unsigned spins = 0; spins < (1u << 24) && __any_sync(0xffffffffu, slot <= last); spins++) {
                if (slot <= last) {
                    double v0, v1;
                    ld_volatile_v2(row + (size_t)slot * SKB, v0, v1);
                    if (!is_sentinel_hi(v0) && slot - s_mcredit <= SW_RB) {
                        asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};"
                                     ::"r"(ring_addr + ((unsigned)slot & (unsigned)(SW_RB - 1)) * (unsigned)(SKB * sizeof(double))), "d"(v0), "d"(v1) : "memory");
                        st_v2(row + (size_t)slot * SKB, sv, sv);
                        slot += 32;
                    }
                }
            }
        }
    } else if (live) {
        const int nsteps = nblk * UB;                                     /* macro-steps of a strip (>= column blocks + 31) */
        const int tfirst = DIR > 0 ? 0 : nsteps - 1;
        const bool has_prev = DIR > 0 ? (k > 0) : (k < g.nstrips - 1);
        const bool has_next = DIR > 0 ? (k < g.nstrips - 1) : (k > 0);
        const bool is_out = lane == E_OUT;
        const bool is_in = lane == E_IN;
        const bool lane0 = lane == 0;
        /* incoming edge values: my shared-memory ring if my predecessor runs in this cluster, else my global inbox row */
        const double* const in_base = !has_prev ? nullptr : &s_inbox[warp][0];
        const unsigned in_mask = (unsigned)(SW_RB - 1);
        const unsigned in_ring_addr = smem_u32(&s_inbox[warp][0]);
        /* outgoing edge values: the next strip's ring (mapped shared memory of its block), or its global inbox row --
         * in the neighbour GPU's memory when the next strip belongs to the next rank */
        const bool out_ring = (cslot < csize * SW_WARPS - 1) && (slot + 1 < nown);
        double* out_base = nullptr;
        unsigned out_mask = 0xffffffffu;
        if (out_ring) {
            out_base = cluster.map_shared_rank(&s_inbox[(cslot + 1) % SW_WARPS][0], (cslot + 1) / SW_WARPS);
            out_mask = (unsigned)(SW_RB - 1);
        } else if (has_next) {
            const bool to_peer = inbox_peer != nullptr && (DIR > 0 ? k == g.s1 - 1 : k == g.s0);
            out_base = (to_peer ? inbox_peer : inbox) + (long long)(k + DIR) * inbox_len;
        }
        const bool can_pub = out_base != nullptr;
        volatile int* const credit_of_prev = !has_prev ? nullptr :
            (cslot > 0 ? cluster.map_shared_rank((int*)&s_credit[(cslot - 1) % SW_WARPS], (cslot - 1) / SW_WARPS) : (int*)&s_mcredit);
        const bool pub_credit = lane0 && credit_of_prev != nullptr;
        double* const dummy = &s_dummy[warp][lane][0];
        const int last_edge_block = nblk - 4;           /* the producer's last real macro-step is taken in this block */

        /* Incoming edge values: consumer macro-step c needs the pair of producer macro-step c + 31 = pair slot c + 32 of
         * the ring in shared memory (written by the strip before, or by the edge mover).  Every lane reads the same
         * addresses (a broadcast), so the flag-in-data test is warp uniform.  The 8 pairs of a block are taken at its
         * start; while one is missing the whole fetch is repeated, so a strip that has caught up with its producer
         * advances block by block with it, 39 macro-steps behind.  (A pair is stored with ONE 16-byte store, so testing
         * its first value is enough.) */
        /* credit: the last macro-step of the next two blocks must not lap the consumer's ring */
        auto credit_wait = [&](const int last_step) {
            if (!out_ring) return;
            for (unsigned spins = 0; last_step - s_credit[warp] >= SW_RB && spins < (1u << 24); spins++) {}
        };

        /* operands of one macro-step: the two adjacent cells of every staged array (one 16-byte shared-memory load per
         * array), fetched one macro-step ahead of their use */
        struct Ops { double a[SKB], cx[SKB], cy[SKB], pr[SKB], rr[SKB]; };
        auto load_step = [&](Ops& o, const unsigned tile, const int row) {
            const unsigned p = tile + (unsigned)((DIR > 0 ? row : U - 1 - row) * (32 * SKB) * (int)sizeof(double));
            lds_cells(p, o.a);
            lds_cells(p + SW_TILE * (unsigned)sizeof(double), o.cx);
            lds_cells(p + 2 * SW_TILE * (unsigned)sizeof(double), o.cy);
            lds_cells(p + 3 * SW_TILE * (unsigned)sizeof(double), o.pr);
            if (WITH_DOT) lds_cells(p + (SW_TILE + SW_PACK) * (unsigned)sizeof(double), o.rr);
        };
        double d0 = 0.0, d1 = 0.0, cxp = 0.0;
        double sh0 = 0.0, sh1 = 0.0;                      /* the adjacent lane's results of its previous macro-step */
        /* running pointer to this lane's cells of the current block's first macro-step */
        double* pd = dst + (long long)k * g.ss + ((long long)tfirst * 32 + lane) * SKB;

        const unsigned ring_lane = pin_u32(smem_u32(ring + lane * SKB));
        const unsigned full0 = pin_u32(smem_u32(FUSE ? bar_ready : bar_full));   /* what tells this warp that a tile is usable */
        const unsigned empty0 = pin_u32(smem_u32(bar_empty));
        int st = 0;
        unsigned ph = 0;
        Ops ops[2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < SKB; j++) { ops[i].a[j] = 0.0; ops[i].cx[j] = 0.0; ops[i].cy[j] = 0.0; ops[i].pr[j] = 0.0; ops[i].rr[j] = 0.0; }
        double e[UB][SKB];                                /* incoming edge pairs, indexed by macro-step & 7 */
#pragma unroll
        for (int i = 0; i < UB; i++) { e[i][0] = 0.0; e[i][1] = 0.0; }
        long long tr_first = 0;

        /* One block = one staged tile = 8 macro-steps, fully unrolled: the recurrence, two shuffles, the operand loads of the
         * next macro-step (from the next tile at the last one, which is awaited first; the loader provides one tile more
         * than the strip has) and two stores per macro-step.  The edge lane's pairs go to slots s0+1 .. s0+8 of the next
         * strip's inbox: the first seven are consecutive (`pf`), only the last one can wrap around a ring (`pl`).
         * EDGE = false: no incoming edge (first strip; columns right / left of the grid). */
        auto run_block = [&](auto edge_c, const int b) {
            constexpr bool EDGE = decltype(edge_c)::value;
            const int s0 = b * UB;
            const bool pub = is_out && can_pub && b >= 3;       /* only the lane at the outgoing edge stores (predicated) */
            double* pf = dummy;
            double* pl = dummy;
            if (pub) {
                pf = out_base + (size_t)((unsigned)(s0 + 1) & out_mask) * SKB;
                pl = out_base + (size_t)((unsigned)(s0 + UB) & out_mask) * SKB;
            }
            const unsigned tile = ring_lane + (unsigned)(st * STAGE * (int)sizeof(double));
            unsigned done_bar = 0, tnext = tile;
            /* the pairs of `G` macro-steps at once (slots s0+32+u0 ..); while one is missing the fetch is repeated */
            auto edge_take = [&](const int u0) {
                const unsigned q = in_ring_addr + ((unsigned)(s0 + 32) & in_mask) * (unsigned)(SKB * sizeof(double));
                for (unsigned spins = 0;; spins++) {
                    bool ok = true;
#pragma unroll
                    for (int i = u0; i < u0 + IFL_SW_G; i++) {
                        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e[i][0]), "=d"(e[i][1]) : "r"(q + 16u * i) : "memory");
                        ok = ok && !is_sentinel_hi(e[i][0]);
                    }
                    if (ok) break;
                    if (spins > (1u << 25)) {                  /* dead producer: give up instead of hanging the GPU */
#pragma unroll
                        for (int i = u0; i < u0 + IFL_SW_G; i++) if (is_sentinel_hi(e[i][0])) { e[i][0] = 0.0; e[i][1] = 0.0; }
                        break;
                    }
                }
            };
#pragma unroll
            for (int u = 0; u < UB; u++) {
                /* the upper neighbours: shuffled out of the adjacent lane as soon as it had them (below), i.e. half a
                 * macro-step / a whole macro-step before they are needed */
                double up0 = sh0, up1 = sh1;
                if (EDGE && u % IFL_SW_G == 0) edge_take(u);
                if (u == UB - 1) {
                    done_bar = empty0 + 8u * (unsigned)st;
                    const bool wrap = st == S - 1;
                    st = wrap ? 0 : st + 1;
                    ph ^= wrap ? 1u : 0u;
                    tnext = ring_lane + (unsigned)(st * STAGE * (int)sizeof(double));
                    mbar_wait_u32(full0 + 8u * (unsigned)st, ph);
                }
                load_step(ops[(u + 1) & 1], tnext, (u + 1) % U);
                const Ops& o = ops[u & 1];
                if (is_in) { up0 = EDGE ? e[u][0] : 0.0; up1 = EDGE ? e[u][1] : 0.0; }
                double n0, n1;
                bool f0 = true, f1 = true;
                if (DIR > 0) {
                    double v = o.a[0] - cxp * d1;       /* (aPlusX[i-1]*precon[i-1]) * dst[i-1]   F3:500 */
                    v = v - o.cy[0] * up0;              /* (aPlusY[i-w]*precon[i-w]) * dst[i-w]   F3:503 */
                    n0 = v * o.pr[0];
                    if (MASKED) { f0 = o.pr[0] == o.pr[0]; n0 = f0 ? n0 : 0.0; }   /* NaN precon tags a non-fluid cell: dst stays, */
                    sh0 = __shfl_up_sync(0xffffffffu, n0, 1);                      /* neighbours see +0.0 (k_precon_coeffs)        */
                    v = o.a[1] - o.cx[0] * n0;
                    v = v - o.cy[1] * up1;
                    n1 = v * o.pr[1];
                    if (MASKED) { f1 = o.pr[1] == o.pr[1]; n1 = f1 ? n1 : 0.0; }
                    sh1 = __shfl_up_sync(0xffffffffu, n1, 1);
                    cxp = o.cx[1];
                } else {
                    double v = o.a[1] - o.cx[1] * d0;   /* (aPlusX[i]*precon[i]) * dst[i+1]       F3:514 */
                    v = v - o.cy[1] * up1;              /* (aPlusY[i]*precon[i]) * dst[i+w]       F3:517 */
                    n1 = v * o.pr[1];
                    if (MASKED) { f1 = o.pr[1] == o.pr[1]; n1 = f1 ? n1 : 0.0; }
                    sh1 = __shfl_down_sync(0xffffffffu, n1, 1);
                    v = o.a[0] - o.cx[0] * n1;
                    v = v - o.cy[0] * up0;
                    n0 = v * o.pr[0];
                    if (MASKED) { f0 = o.pr[0] == o.pr[0]; n0 = f0 ? n0 : 0.0; }
                    sh0 = __shfl_down_sync(0xffffffffu, n0, 1);
                }
                d0 = n0; d1 = n1;
                /* unconditional stores: inactive lanes store their +0.0 into padding; a non-fluid cell (MASKED) stores into
                 * a per-lane dummy slot (a predicated store becomes a divergent branch) */
                double* ps = pd + DIR * u * (32 * SKB);
                if (MASKED) { *(f0 ? ps : dummy) = d0; *(f1 ? ps + 1 : dummy + 1) = d1; }
                else *reinterpret_cast<double2*>(ps) = make_double2(d0, d1);
                st_v2_if(pub, u < UB - 1 ? pf + u * SKB : pl, d0, d1);
                if (WITH_DOT) {                          /* inactive lanes hold +0.0, rr is zero in the padding: adding them is exact */
                    if (DIR > 0) { acc += d0 * o.rr[0]; acc += d1 * o.rr[1]; }
                    else         { acc += d1 * o.rr[1]; acc += d0 * o.rr[0]; }
                }
                /* every row of the finished tile is in registers (the loads are warp-wide instructions issued, in order,
                 * before this arrive): hand the stage back to the loader warp */
                if (u == UB - 1) mbar_arrive_if(lane0, done_bar);
            }
            pd += DIR * UB * (32 * SKB);
        };

        /* prologue: the producer publishes from its block 3 on (macro-steps 24..); its macro-steps 24..30 (pair slots
         * 25..31) are needed by nobody -- they are awaited and re-armed so that a ring slot is always armed when its turn
         * comes */
        if (in_base != nullptr) {
            for (int i = 25; i <= 31; i++) {
                double v0, v1;
                for (unsigned spins = 0; spins < (1u << 26); spins++) {
                    ld_volatile_v2(in_base + i * SKB, v0, v1);
                    if (!is_sentinel_hi(v0) && !is_sentinel_hi(v1)) break;
                }
            }
            if (lane0) {
                double* q = const_cast<double*>(in_base);
                for (int i = 25 * SKB; i < 32 * SKB; i++) q[i] = sentinel();
            }
            __syncwarp();
        }
        mbar_wait_u32(full0, 0);
        load_step(ops[0], ring_lane, 0);
        /* the blocks that take edge values: 0 .. last_edge_block (the producer's last real macro-step is in the last one);
         * after each, lanes 0..7 re-arm its 8 slots with one store and the credit goes back to the producer */
        int b_rest = 0;
        if (in_base != nullptr) {
            const double sv = sentinel();
            for (int b = 0; b <= last_edge_block; b++) {
                if ((b & 1) == 0) credit_wait(b * UB + 2 * UB - 1);
                run_block(std::true_type(), b);
                if (lane < UB) st_v2(const_cast<double*>(in_base) + (size_t)(((unsigned)(b * UB + 32) & in_mask) + (unsigned)lane) * SKB, sv, sv);
                __syncwarp();
                if (pub_credit) *credit_of_prev = (b + 1) * UB + 31;                        /* producer macro-steps I have consumed */
                if (trace && b == 0) tr_first = (long long)globaltimer_ns();
            }
            b_rest = last_edge_block + 1;
        }
        for (int b = b_rest; b < nblk; b++) {
            if ((b & 1) == 0) credit_wait(b * UB + 2 * UB - 1);
            run_block(std::false_type(), b);
            if (trace && b == 0) tr_first = (long long)globaltimer_ns();
        }
        if (can_pub) {
            /* 8 more zero pairs: the consumer takes 8 pairs per block through its block nblk-4 */
            credit_wait(nsteps + UB - 1);
            if (is_out) {
#pragma unroll
                for (int i = 0; i < UB; i++) st_v2(out_base + (size_t)((unsigned)(nsteps + 1 + i) & out_mask) * SKB, 0.0, 0.0);
            }
        }
        if (trace && lane == 0) {
            long long* tr = trace + (long long)k * 16;
            tr[0] = tr_enter; tr[1] = tr_first; tr[2] = 0; tr[3] = (long long)globaltimer_ns();
            tr[4] = 0; tr[5] = nblk; tr[6] = slot; tr[7] = (long long)get_smid();
        }
    }
    if (WITH_DOT && live && !loader && !mover) {
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
    constexpr int RS = 32 * SKB;                     /* doubles per m-row */
    const int slot = claim_slot(&sc->ticket_f, g.nstrips);
    if (slot < 0) return;
    const int k = slot;
    const int lane = threadIdx.x & 31;
    const int y = k * 32 + lane;
    const bool yok = y < g.h;
    const int w = g.w;
    const int cw = (w + SKB - 1) / SKB;
    const long long base = (long long)k * g.ss + lane * SKB;
    const int nblk = (cw + 31 + U - 1) / U;
    long long upbase = base - 33 * SKB;
    if (lane == 0) upbase = (long long)(k - 1) * g.ss + (31LL * 32 + 31) * SKB;
    const bool has_up = y > 0;
    const bool masked = fl != nullptr;

    double AD[U][SKB], AX[U][SKB], AY[U][SKB], AXU[U][SKB], AYU[U][SKB];
    unsigned char FL[U][SKB], FLU[U][SKB];
    auto fetch = [&](int u, int m) {
#pragma unroll
        for (int j = 0; j < SKB; j++) {
            const long long i = base + (long long)m * RS + j;
            const long long iu = upbase + (long long)m * RS + j;
            const int x = (m - lane) * SKB + j;
            AD[u][j] = adiag[i];
            AX[u][j] = ax[i];
            AY[u][j] = ay[i];
            const bool okup = has_up && x >= 0 && x < w;
            AXU[u][j] = okup ? ax[iu] : 0.0;
            AYU[u][j] = okup ? ay[iu] : 0.0;
            FL[u][j] = masked ? fl[i] : 1;
            FLU[u][j] = (masked && okup) ? fl[iu] : 1;
        }
    };
#pragma unroll
    for (int u = 0; u < U; u++) fetch(u, u);
    EdgeFeed<U> feed;
    feed.init(k > 0 ? bnd + (long long)(k - 1) * w : nullptr, w);
    feed.prefetch(0, lane);
    feed.advance();

    double pprev[SKB];
#pragma unroll
    for (int j = 0; j < SKB; j++) pprev[j] = 0.0;
    double pl = 0.0, axprev = 0.0, ayprev = 0.0;
    bool flprev = false;
    for (int b = 0; b < nblk; b++) {
        const int m0 = b * U;
        feed.prefetch(m0 + U, lane);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            double pu[SKB];
#pragma unroll
            for (int j = 0; j < SKB; j++) {
                pu[j] = __shfl_up_sync(0xffffffffu, pprev[j], 1);
                const double bv = feed.get(u * SKB + j, lane);
                if (lane == 0) pu[j] = bv;
            }
#pragma unroll
            for (int j = 0; j < SKB; j++) {
                const int x = (m - lane) * SKB + j;
                const double ad = AD[u][j];
                const bool fself = FL[u][j] != 0;
                double e = ad;
                if (x > 0 && flprev) {
                    double px = axprev * pl;
                    double py = ayprev * pl;
                    e -= (px * px + tau * px * py);
                }
                if (y > 0 && FLU[u][j]) {
                    double px = AXU[u][j] * pu[j];
                    double py = AYU[u][j] * pu[j];
                    e -= (py * py + tau * px * py);
                }
                if (e < sigma * ad) e = ad;
                const double pc = 1.0 / sqrt(e);
                const bool active = yok && x >= 0 && x < w;
                if (active) {
                    if (fself) pre[base + (long long)m * RS + j] = pc;
                    if (lane == 31) bnd[(long long)k * w + x] = fself ? pc : 0.0;   /* flag-in-data: always publish */
                }
                pprev[j] = pc; pl = pc; axprev = AX[u][j]; ayprev = AY[u][j]; flprev = fself;
            }
            if (m + U < g.tl) fetch(u, m + U);
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
    constexpr int RS = 32 * SKB;
    if (sc->done) return;
    const int slot = claim_slot(&sc->ticket_f, g.nstrips);
    if (slot < 0) return;
    const int k = slot;
    const int lane = threadIdx.x & 31;
    const int y = k * 32 + lane;
    const bool yok = y < g.h;
    const int w = g.w, h = g.h;
    const int cw = (w + SKB - 1) / SKB;
    const long long base = (long long)k * g.ss + lane * SKB;
    const int nblk = (cw + 31 + U - 1) / U;
    for (int x = lane; x < w; x += 32) bnd_rearm[(long long)k * w + x] = sentinel();

    /* old value of the cell below: (x, y+1) is 33 lanes-slots further, or in the next strip's lane 0 */
    long long dnbase = base + 33 * SKB;
    if (lane == 31) dnbase = (long long)(k + 1) * g.ss - 31LL * RS;
    const bool has_dn = y < h - 1;

    double R[U][SKB], PO[U][SKB], PD[U][SKB];
    auto fetch = [&](int u, int m) {
#pragma unroll
        for (int j = 0; j < SKB; j++) {
            const int x = (m - lane) * SKB + j;
            R[u][j] = r[base + (long long)m * RS + j];
            PO[u][j] = __ldcg(p + base + (long long)m * RS + j);
            const bool okdn = has_dn && x >= 0 && x < w;
            PD[u][j] = okdn ? __ldcg(p + dnbase + (long long)m * RS + j) : 0.0;
        }
    };
#pragma unroll
    for (int u = 0; u < U; u++) fetch(u, u);
    EdgeFeed<U> feed;
    feed.init(k > 0 ? bnd_use + (long long)(k - 1) * w : nullptr, w);
    feed.prefetch(0, lane);
    feed.advance();

    double pprev[SKB];
#pragma unroll
    for (int j = 0; j < SKB; j++) pprev[j] = 0.0;
    double pl = 0.0;
    unsigned long long mbits = 0ull;
    for (int b = 0; b < nblk; b++) {
        const int m0 = b * U;
        feed.prefetch(m0 + U, lane);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            double pu[SKB];
#pragma unroll
            for (int j = 0; j < SKB; j++) {
                pu[j] = __shfl_up_sync(0xffffffffu, pprev[j], 1);
                const double bv = feed.get(u * SKB + j, lane);
                if (lane == 0) pu[j] = bv;
            }
#pragma unroll
            for (int j = 0; j < SKB; j++) {
                const int x = (m - lane) * SKB + j;
                const double pold = PO[u][j];
                /* p(x+1, y) = own old value of the next cell; the registers of the next macro-step still
                 * hold it (they are refilled only after their own use) */
                const double pright = (j < SKB - 1) ? PO[u][(j + 1) % SKB] : PO[(u + 1) % U][0];
                double diag = 0.0, offDiag = 0.0;
                if (x > 0)     { diag += scale; offDiag -= scale * pl; }
                if (y > 0)     { diag += scale; offDiag -= scale * pu[j]; }
                if (x < w - 1) { diag += scale; offDiag -= scale * pright; }
                if (y < h - 1) { diag += scale; offDiag -= scale * PD[u][j]; }
                const double newP = (R[u][j] - offDiag) / diag;
                const bool active = yok && x >= 0 && x < w;
                if (active) {
                    /* F1:414: old p rounded through float, difference taken in double */
                    double delta = fabs(jfloat(pold) - newP);
                    unsigned long long db = (unsigned long long)__double_as_longlong(delta);
                    mbits = db > mbits ? db : mbits;
                    p[base + (long long)m * RS + j] = newP;
                    if (lane == 31) bnd_use[(long long)k * w + x] = newP;
                }
                pprev[j] = newP; pl = newP;
            }
            if (m + U < g.tl) fetch(u, m + U);
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

/* project() F1:380 (=F2:395) for grids whose pressure fits in shared memory (config 1: 128x128): ALL Gauss-Seidel
 * iterations of a solve in ONE launch of ONE block.  p lives in shared memory for the whole solve; warp k owns the strip of
 * rows 32k..32k+31 (lane l = row l) and walks the anti-diagonals x = m - l with a one-cell skew, so that the cell to the
 * left is its own previous result, the cell above the previous result of lane l-1 (one shuffle) and the cells to the right
 * and below still hold the values of the previous iteration -- exactly the operands of the reference's in-place
 * lexicographic loop, hence the same bits.  Strip k+1 trails strip k by 32 cells (a progress counter in shared memory),
 * which is what makes the row above its first row new and the row below the last row of strip k old.  After every sweep
 * the block reduces maxDelta (F1:414, a max: order independent) and decides like F1:420.
 * A sweep costs (w + 31 + 32*(strips-1)) dependent cell steps -- about 20 us at 128x128 instead of one 100-180 us kernel
 * launch per sweep. */
__global__ void __launch_bounds__(1024)
k_gs_resident(SkewGeom g, const double* __restrict__ r, double* __restrict__ p, SolveScalars* sc, double scale, double tol, int limit) {
    extern __shared__ double sp[];                      /* p, row-major [h][w] */
    __shared__ volatile int s_prog[32];                 /* columns a strip has finished in the current sweep */
    __shared__ unsigned long long s_max[32];
    __shared__ int s_done;
    const int w = g.w, h = g.h;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5, nst = blockDim.x >> 5;
    const int y = k * 32 + lane;
    const bool yok = y < h;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) sp[i] = p[skew_index(g, i % w, i / w)];
    if (threadIdx.x < 32) s_prog[threadIdx.x] = 0;
    __syncthreads();
    /* r of my row in the skewed layout: (x, y) sits at strip*ss + (((x >> 1) + l) * 32 + l) * 2 + (x & 1) */
    static_assert(IFL_SKB == 2, "index arithmetic of the skewed layout");
    const double* const rrow = r + (long long)k * g.ss + (long long)(lane * 32 + lane) * 2;
    auto r_at = [&](int x) { return rrow[((x >> 1) << 6) + (x & 1)]; };
    volatile double* const vsp = sp;                    /* rows other warps write are read through volatile accesses */
    int iterations = 0;
    double max_error = 0.0;
    bool done = false;
    const int nsteps = w + 31;
    constexpr int PF = 4;                               /* r is fetched PF steps ahead (it stays in L1 / L2) */
    const int yc = yok ? y : h - 1;                     /* rows below the grid compute on a copy of the last row; nothing is stored */
    const bool has_up = yc > 0, has_dn = yc < h - 1;
    for (int it = 0; it < limit && !done; it++) {
        double pl = 0.0, pnew = 0.0;
        unsigned long long mbits = 0ull;
        double rq[PF];
#pragma unroll
        for (int i = 0; i < PF; i++) { const int xq = i - lane; rq[i] = (yok && xq >= 0 && xq < w) ? r_at(xq) : 0.0; }
        int seen = 0;                                   /* columns of the strip above known to be finished */
        for (int m0 = 0; m0 < nsteps; m0 += PF) {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const int m = m0 + i;
                const int x = m - lane;
                const bool act = yok && x >= 0 && x < w && m < nsteps;
                /* the row above my strip must be finished up to column m (lane 0 needs p(x = m, 32k - 1) of this sweep) */
                if (k > 0 && seen <= m && m < w) { do { seen = s_prog[k - 1]; } while (seen <= m); }
                double pu = __shfl_up_sync(0xffffffffu, pnew, 1);
                const double rv = rq[i];
                { const int xq = m + PF - lane; rq[i] = (yok && xq >= 0 && xq < w) ? r_at(xq) : 0.0; }
                /* branch-free: out-of-grid lanes work on a clamped cell and store nothing; a term the reference skips
                 * (x > 0 ... F1:399-410) enters as +0.0, which changes neither diag (>= +0.0) nor offDiag (never -0.0) */
                const int xc = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
                const int idx = yc * w + xc;
                const double pup = lane == 0 ? (has_up ? vsp[idx - (has_up ? w : 0)] : 0.0) : pu;
                const double pold = sp[idx];
                const double prt = sp[idx + (xc < w - 1 ? 1 : 0)];
                const double pdn = lane == 31 ? vsp[idx + (has_dn ? w : 0)] : sp[idx + (has_dn ? w : 0)];
                const bool cl = xc > 0, cu = has_up, cr = xc < w - 1, cd = has_dn;
                double diag = 0.0, offDiag = 0.0;
                diag += cl ? scale : 0.0; offDiag -= cl ? scale * pl : 0.0;
                diag += cu ? scale : 0.0; offDiag -= cu ? scale * pup : 0.0;
                diag += cr ? scale : 0.0; offDiag -= cr ? scale * prt : 0.0;
                diag += cd ? scale : 0.0; offDiag -= cd ? scale * pdn : 0.0;
                const double newP = (rv - offDiag) / diag;
                /* F1:414: old p rounded through float, difference taken in double */
                const double delta = fabs(jfloat(pold) - newP);
                const unsigned long long db = (unsigned long long)__double_as_longlong(delta);
                if (act) {
                    mbits = db > mbits ? db : mbits;
                    sp[idx] = newP;
                    pl = newP; pnew = newP;
                }
                __syncwarp();
                /* lane 31 has finished the columns up to m - 31: the strip below may use them (and their old values are no
                 * longer needed here).  Published every 8 steps: one fence per 8 cells instead of per cell. */
                if (lane == 31 && m >= 31 && m < nsteps && ((m - 31) & 7) == 7) { __threadfence_block(); s_prog[k] = m - 30; }
            }
        }
        if (lane == 31) { __threadfence_block(); s_prog[k] = w; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_down_sync(0xffffffffu, mbits, o);
            mbits = other > mbits ? other : mbits;
        }
        if (lane == 0) s_max[k] = mbits;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long mb = 0ull;
            for (int i = 0; i < nst; i++) mb = s_max[i] > mb ? s_max[i] : mb;
            const double md = __longlong_as_double((long long)mb);
            s_done = md < tol ? 1 : 0;
            *(double*)&s_max[0] = md;
        }
        if (threadIdx.x < 32) s_prog[threadIdx.x] = 0;
        __syncthreads();
        iterations++;
        max_error = *(double*)&s_max[0];
        done = s_done != 0;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) p[skew_index(g, i % w, i / w)] = sp[i];
    if (threadIdx.x == 0) {
        sc->iterations = iterations;
        sc->max_error = max_error;
        sc->done = done ? 1 : 0;
    }
}
/* does the resident Gauss-Seidel kernel apply?  (p in shared memory, one warp per strip) */
static inline bool gs_fits_resident(const SkewGeom& g) {
    return (size_t)g.w * g.h * sizeof(double) <= 200 * 1024 && g.nstrips <= 32;
}

/* slab decomposition: boundary rows of the current z and s <-> packed rows.
 * rowbuf = [z first row | s first row | z last row | s last row | from above: z, s | from below: z, s], w doubles each */
__global__ void k_pack_rows(SkewGeom g, const double* z0, const double* z1, const double* s0, const double* s1,
                            const SolveScalars* sc, double* __restrict__ rowbuf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    const double* __restrict__ z = sc->zpar ? z1 : z0;
    const double* __restrict__ s = sc->zpar ? s1 : s0;
    const long long ifirst = skew_index(g, x, g.s0 * 32);                               /* my first row -> rank-1 */
    int ylast = g.s1 * 32 - 1;
    if (ylast > g.h - 1) ylast = g.h - 1;
    const long long ilast = skew_index(g, x, ylast);                                    /* my last row -> rank+1 */
    rowbuf[x] = z[ifirst];
    rowbuf[g.w + x] = s[ifirst];
    rowbuf[2 * g.w + x] = z[ilast];
    rowbuf[3 * g.w + x] = s[ilast];
}
__global__ void k_unpack_rows(SkewGeom g, double* z0, double* z1, double* s0, double* s1, const SolveScalars* sc,
                              const double* __restrict__ rowbuf, int has_up, int has_down) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    double* __restrict__ z = sc->zpar ? z1 : z0;
    double* __restrict__ s = sc->zpar ? s1 : s0;
    if (has_up) {                                                                       /* row above my slab */
        const long long i = skew_index(g, x, g.s0 * 32 - 1);
        z[i] = rowbuf[4 * g.w + x]; s[i] = rowbuf[5 * g.w + x];
    }
    if (has_down) {                                                                     /* row below my slab */
        const long long i = skew_index(g, x, g.s1 * 32);
        z[i] = rowbuf[6 * g.w + x]; s[i] = rowbuf[7 * g.w + x];
    }
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

/* every enqueue helper returns 0 or -1 (communication failure, comm_last_error()); launch errors are picked up by the
 * caller's cudaGetLastError() */
#define IFL_TRY(call) do { if ((call) != 0) return -1; } while (0)

int launch_build_precon(SolveCtx& c) {
    k_sweep_rearm<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.inbox_f, c.inbox_b, c.inbox_len, c.sc);
    k_build_precon<<<wf_blocks(c.g), WF_WARPS * 32, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.precon,
                                                                  c.bnd_f, c.sc, c.mic_tau, c.mic_sigma, c.fl);
    /* the MIC(0) factor is built for the whole grid on every rank (a recurrence from strip 0); the packed products only
     * for the strips this rank sweeps */
    k_precon_coeffs<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.aplusx, c.aplusy, c.precon, c.pack_f, c.pack_b, c.fl);
    count(c, 3);
    return 0;
}

static size_t sweep_smem(bool two_streams) {
    size_t stage = ((two_streams ? 2 : 1) * SW_TILE + SW_PACK) * sizeof(double);
    return (size_t)SW_WARPS * SW_STAGES * stage + (size_t)SW_WARPS * 3 * SW_STAGES * 8;
}
/* src == nullptr: in place on the current z generation */
template <int DIR, bool WITH_DOT, bool MASKED, int FUSE>
static void launch_sweep(SolveCtx& c, const double* src, const double* pack, const double* rr, double* inbox,
                         int dot_mode, int check_done, long long* trace) {
    /* the opt-in to > 48 KB of dynamic shared memory is per device and per kernel instance */
    static unsigned long long attr_devices = 0ull;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((attr_devices >> dev) & 1ull)) {
        cudaFuncSetAttribute(k_precon_sweep<DIR, WITH_DOT, MASKED, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sweep_smem(WITH_DOT || FUSE));
        if (SW_CL_MAX > 8)
            cudaFuncSetAttribute(k_precon_sweep<DIR, WITH_DOT, MASKED, FUSE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (dev < 64) __atomic_fetch_or(&attr_devices, 1ull << dev, __ATOMIC_RELAXED);
    }
    const int nown = c.g.s1 - c.g.s0;
    double* peer = nullptr;
    if (c.comm) peer = DIR > 0 ? c.peer_inbox_f_next : c.peer_inbox_b_prev;
    const int nblocks = (nown + SW_WARPS - 1) / SW_WARPS;
    int cl = SW_CL_MAX;
    while (cl > 1 && cl > nblocks) cl >>= 1;                       /* cluster of 8 blocks = 8 * SW_WARPS consecutive strips */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((nblocks + cl - 1) / cl) * cl);
    cfg.blockDim = dim3(SW_THREADS);                               /* strip warps + loader warps + edge mover */
    cfg.dynamicSmemBytes = sweep_smem(WITH_DOT || FUSE);
    cfg.stream = c.stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_precon_sweep<DIR, WITH_DOT, MASKED, FUSE>, c.g, src, c.zb[0], c.zb[1], pack, rr, inbox, peer,
                       c.inbox_len, c.strip_part, c.sc, dot_mode, check_done, trace, c.tolerance, c.rank_max,
                       c.comm ? c.comm->rank : 0, single_gpu(c));
}

static inline void prof_mark(SolveCtx& c, int slot, int e);
static int dist_max(SolveCtx& c, int init);
/* z = M^-1 r on the current z generation: forward then backward sweep.  Boundary rows must be armed.
 * fuse_r_update: the forward sweep first performs r -= alpha*z (z = A s), the norm and the convergence decision. */
static int enqueue_precon(SolveCtx& c, int fuse_r_update, int dot_mode, int check_done, int ps = -1) {
    const bool masked = c.fl != nullptr;
    if (fuse_r_update && c.separate_r_update) {
        k_update_pr<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.r, c.zb[0], c.zb[1], c.sc, c.tolerance,
                                                              c.mask_vector_ops ? c.fl : nullptr, c.rank_max,
                                                              c.comm ? c.comm->rank : 0, single_gpu(c));
        count(c);
        IFL_TRY(dist_max(c, 0));
        fuse_r_update = 0;
    }
    if (fuse_r_update) {
        if (!masked)                launch_sweep<1, false, false, 1>(c, c.r, c.pack_f, nullptr, c.inbox_f, DOT_NONE, check_done, c.trace);
        else if (c.mask_vector_ops) launch_sweep<1, false, true, 2>(c, c.r, c.pack_f, nullptr, c.inbox_f, DOT_NONE, check_done, c.trace);
        else                        launch_sweep<1, false, true, 1>(c, c.r, c.pack_f, nullptr, c.inbox_f, DOT_NONE, check_done, c.trace);
        count(c);
        IFL_TRY(dist_max(c, 0));
    } else {
        if (masked) launch_sweep<1, false, true, 0>(c, c.r, c.pack_f, nullptr, c.inbox_f, DOT_NONE, check_done, c.trace);
        else        launch_sweep<1, false, false, 0>(c, c.r, c.pack_f, nullptr, c.inbox_f, DOT_NONE, check_done, c.trace);
        count(c);
    }
    prof_mark(c, ps, 2);
    int fused_mode = c.sequential ? DOT_NONE : dot_mode;
    long long* tr2 = c.trace ? c.trace + 16 * c.g.nstrips : nullptr;
    if (dot_mode != DOT_NONE) {
        if (masked) launch_sweep<-1, true, true, 0>(c, nullptr, c.pack_b, c.r, c.inbox_b, fused_mode, check_done, tr2);
        else        launch_sweep<-1, true, false, 0>(c, nullptr, c.pack_b, c.r, c.inbox_b, fused_mode, check_done, tr2);
    } else {
        if (masked) launch_sweep<-1, false, true, 0>(c, nullptr, c.pack_b, nullptr, c.inbox_b, DOT_NONE, check_done, tr2);
        else        launch_sweep<-1, false, false, 0>(c, nullptr, c.pack_b, nullptr, c.inbox_b, DOT_NONE, check_done, tr2);
    }
    count(c);
    if (c.sequential && dot_mode != DOT_NONE) {
        k_dot_sequential<<<1, EW_THREADS, 0, c.stream>>>(c.g, c.zb[0], c.zb[1], c.r, nullptr, nullptr, c.sc, dot_mode, check_done,
                                                 c.mask_vector_ops ? c.fl : nullptr);
        count(c);
    }
    return 0;
}

/* applyPreconditioner(z, r) on its own (IFL_STAGE_APPLY_PRECON, the wavefront trace) */
int launch_apply_precon(SolveCtx& c, int as_in_loop) {
    k_sweep_rearm<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.inbox_f, c.inbox_b, c.inbox_len, c.sc);
    count(c);
    /* slab decomposition: no neighbour may store into an inbox row before its owner has re-armed it */
    if (!single_gpu(c)) IFL_TRY(comm_all_gather(*c.comm, c.rank_max + c.comm->rank, c.rank_max, sizeof(unsigned long long)));
    /* as_in_loop (wavefront trace): the kernels of a PCG iteration -- r update fused into the forward sweep, dot product
     * fused into the backward sweep -- with whatever alpha the last solve left behind */
    return as_in_loop ? enqueue_precon(c, 1, DOT_SIGMA_NEW, 0) : enqueue_precon(c, 0, DOT_NONE, 0);
}

/* z = A s on the current generation (IFL_STAGE_MATVEC) */
int launch_matvec(SolveCtx& c) {
    const int zp = solve_parity(c);
    if (zp < 0) return -1;
    if (c.const_matrix)
        k_matvec<true><<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.sb[zp], c.zb[zp], c.const_scale);
    else
        k_matvec<false><<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.sb[zp], c.zb[zp], 0.0);
    count(c);
    return 0;
}

/* which generation of z / s is current (host side: stages and field access of the tests; synchronises the stream) */
int solve_parity(SolveCtx& c) {
    int zp = 0;
    if (cudaMemcpyAsync(&zp, &c.sc->zpar, sizeof(int), cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c.stream) != cudaSuccess) return -1;
    return zp & 1;
}

/* ---- distributed pieces (no-ops on a single GPU) ---- */
static int dist_dot(SolveCtx& c, int mode, int check_done) {
    if (single_gpu(c) || mode == DOT_NONE) return 0;
    IFL_TRY(comm_all_gather_strips(*c.comm, c.strip_part, c.g.nstrips, 1));
    k_apply_dot<<<1, EW_THREADS, 0, c.stream>>>(c.g, c.strip_part, c.sc, mode, check_done);
    count(c);
    return 0;
}
static int dist_max(SolveCtx& c, int init) {
    if (single_gpu(c)) return 0;
    IFL_TRY(comm_all_gather(*c.comm, c.rank_max + c.comm->rank, c.rank_max, sizeof(unsigned long long)));
    k_apply_max<<<1, 32, 0, c.stream>>>(c.rank_max, c.comm->world, c.sc, c.tolerance, init);
    count(c);
    return 0;
}
/* k_step_a of the next iteration reads z and s one row above and one row below the slab */
static int dist_halo_rows(SolveCtx& c) {
    if (single_gpu(c)) return 0;
    const int w = c.g.w;
    k_pack_rows<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, c.zb[0], c.zb[1], c.sb[0], c.sb[1], c.sc, c.rowbuf);
    /* rowbuf = [z up | s up | z down | s down | recv from above (z, s) | recv from below (z, s)] */
    IFL_TRY(comm_exchange_rows(*c.comm, c.rowbuf, c.rowbuf + 4 * w, c.rowbuf + 2 * w, c.rowbuf + 6 * w, (size_t)2 * w));
    k_unpack_rows<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, c.zb[0], c.zb[1], c.sb[0], c.sb[1], c.sc, c.rowbuf,
                                                         c.comm->rank > 0, c.comm->rank < c.comm->world - 1);
    count(c, 2);
    return 0;
}

/* project() F3:590-602: p = 0; z = M^-1 r; s = z; maxError; sigma = z.r */
int pcg_enqueue_init(SolveCtx& c) {
    cudaMemsetAsync(c.p_first, 0, sizeof(double) * (size_t)(c.g.s1 - c.g.s0) * (size_t)c.g.ss, c.stream);
    k_solve_reset<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.inbox_f, c.inbox_b, c.inbox_len, c.sc);
    count(c);
    /* every rank's inboxes must be re-armed before any neighbour starts writing into them */
    if (!single_gpu(c)) IFL_TRY(comm_all_gather(*c.comm, c.rank_max + c.comm->rank, c.rank_max, sizeof(unsigned long long)));
    IFL_TRY(enqueue_precon(c, 0, DOT_SIGMA_INIT, 0));
    IFL_TRY(dist_dot(c, DOT_SIGMA_INIT, 0));
    k_init_s<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.sb[0], c.sb[1], c.zb[0], c.zb[1], c.r, c.sc, c.tolerance,
                                                       c.mask_vector_ops ? c.fl : nullptr,
                                                       c.rank_max, c.comm ? c.comm->rank : 0, single_gpu(c));
    count(c);
    IFL_TRY(dist_max(c, 1));
    return dist_halo_rows(c);
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

/* n iterations of the loop body F3:603-628 as three kernels; every kernel returns at once after convergence */
int pcg_enqueue_iterations(SolveCtx& c, int n) {
    dim3 eg = ew_grid(c.g);
    const unsigned char* mfl = c.mask_vector_ops ? c.fl : nullptr;
    for (int it = 0; it < n; it++) {
        const int ps = (it == 0 && !c.sequential) ? prof_begin(c, 0) : -1;
        const int dm = c.sequential ? DOT_NONE : DOT_ZS;
        if (c.const_matrix)
            k_step_a<true><<<eg, EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.zb[0], c.zb[1], c.sb[0], c.sb[1], c.p,
                                                            c.partials, c.strip_part, c.strip_ctr, c.sc, dm, mfl, single_gpu(c),
                                                            c.const_scale);
        else
            k_step_a<false><<<eg, EW_THREADS, 0, c.stream>>>(c.g, c.adiag, c.aplusx, c.aplusy, c.zb[0], c.zb[1], c.sb[0], c.sb[1], c.p,
                                                             c.partials, c.strip_part, c.strip_ctr, c.sc, dm, mfl, single_gpu(c),
                                                             0.0);
        count(c);
        if (c.sequential) {
            k_dot_sequential<<<1, EW_THREADS, 0, c.stream>>>(c.g, c.zb[0], c.zb[1], nullptr, c.sb[0], c.sb[1], c.sc, DOT_ZS, 1, mfl);
            count(c);
        }
        IFL_TRY(dist_dot(c, DOT_ZS, 1));
        prof_mark(c, ps, 1);
        IFL_TRY(enqueue_precon(c, 1, DOT_SIGMA_NEW, 1, ps));
        IFL_TRY(dist_dot(c, DOT_SIGMA_NEW, 1));
        IFL_TRY(dist_halo_rows(c));
        prof_mark(c, ps, 3);
    }
    return 0;
}

int pcg_enqueue_finish(SolveCtx& c, int limit) {
    /* the last iteration's p += alpha*s */
    k_update_p<<<ew_grid(c.g), EW_THREADS, 0, c.stream>>>(c.g, c.p, c.sb[0], c.sb[1], c.sc, c.mask_vector_ops ? c.fl : nullptr);
    k_solve_finish<<<1, 1, 0, c.stream>>>(c.sc, limit);
    count(c, 2);
    return 0;
}
/* slab decomposition: applyPressure of the first row of a slab reads p of the row above it (F3:640-643 in gather form);
 * every rank sends the last row of its p to the next rank */
__global__ void k_pack_last_row(SkewGeom g, const double* __restrict__ vec, double* __restrict__ rowbuf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    int ylast = g.s1 * 32 - 1;
    if (ylast > g.h - 1) ylast = g.h - 1;
    rowbuf[x] = vec[skew_index(g, x, ylast)];
}
__global__ void k_unpack_row_above(SkewGeom g, double* __restrict__ vec, const double* __restrict__ rowbuf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    vec[skew_index(g, x, g.s0 * 32 - 1)] = rowbuf[x];
}
int pcg_exchange_pressure_halo(SolveCtx& c) {
    if (single_gpu(c)) return 0;
    const int w = c.g.w, rank = c.comm->rank, world = c.comm->world;
    k_pack_last_row<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, c.p, c.rowbuf);
    CommXfer x[2];
    int n = 0;
    if (rank < world - 1) x[n++] = CommXfer{rank + 1, c.rowbuf, (size_t)w, 1};
    if (rank > 0) x[n++] = CommXfer{rank - 1, c.rowbuf + w, (size_t)w, 0};
    IFL_TRY(comm_transfers(*c.comm, x, n));
    if (rank > 0) k_unpack_row_above<<<(w + 255) / 256, 256, 0, c.stream>>>(c.g, c.p, c.rowbuf + w);
    count(c, 2);
    return 0;
}
/* Gauss-Seidel project() F1:380: _p is warm-started, so only the control state is reset */
void gs_enqueue_begin(SolveCtx& c) {
    k_solve_reset<<<rearm_blocks(c.g), 256, 0, c.stream>>>(c.g, c.bnd_f, c.bnd_b, c.inbox_f, c.inbox_b, c.inbox_len, c.sc);
    count(c);
}
/* returns the number of iterations it has enqueued: `n`, or everything up to `limit` at once when the whole solve runs
 * in the resident kernel */
int gs_enqueue_iterations(SolveCtx& c, int first_iter, int n, double scale, int limit) {
    if (gs_fits_resident(c.g) && first_iter == 0) {
        const size_t smem = (size_t)c.g.w * c.g.h * sizeof(double);
        static unsigned long long attr_devices = 0ull;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 64 || !((attr_devices >> dev) & 1ull)) {
            cudaFuncSetAttribute(k_gs_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (dev < 64) __atomic_fetch_or(&attr_devices, 1ull << dev, __ATOMIC_RELAXED);
        }
        const int ps = prof_begin(c, 1);
        k_gs_resident<<<1, 32 * c.g.nstrips, smem, c.stream>>>(c.g, c.r, c.p, c.sc, scale, c.tolerance, limit);
        count(c);
        prof_mark(c, ps, 1);
        return limit;
    }
    for (int it = first_iter; it < first_iter + n; it++) {
        double* use = (it & 1) ? c.bnd_b : c.bnd_f;
        double* rearm = (it & 1) ? c.bnd_f : c.bnd_b;
        const int ps = (it == first_iter) ? prof_begin(c, 1) : -1;
        k_gs_sweep<<<wf_blocks(c.g), WF_WARPS * 32, 0, c.stream>>>(c.g, c.r, c.p, use, rearm, c.sc, scale, c.tolerance);
        count(c);
        prof_mark(c, ps, 1);
    }
    return n;
}
void gs_enqueue_finish(SolveCtx& c, int limit) {
    k_solve_finish<<<1, 1, 0, c.stream>>>(c.sc, limit);
    count(c);
}

} // namespace ifl
