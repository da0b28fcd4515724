/*
 * ifl_device.cuh -- device-side helpers shared by every kernel of libincfluid_b200.
 *
 * Arithmetic contract: the reference is Java double arithmetic without
 * fused multiply-add, so every translation unit is compiled with -fmad=false and the
 * expressions below keep the reference's left-to-right evaluation order.  IEEE double
 * division and sqrt are correctly rounded on sm_100a by default.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IFL_WARP 32
#define IFL_STRIP 32 /* rows per strip of the skewed layout == warp size */

/* cell types of FluidQuantity._cell (F4:60-63 enum CellType; F8:1795 adds EMPTY) */
#define IFL_CELL_FLUID 0
#define IFL_CELL_SOLID 1
#define IFL_CELL_EMPTY 2

namespace ifl {

/* ---- Java semantics ---------------------------------------------------------- */
/* Math.max / Math.min on doubles.  CUDA fmax/fmin differ from Java only for NaN
 * operands (Java propagates, CUDA suppresses); a NaN residual terminates the
 * reference (F3:616), and the solver reports it through IFL_ERR_NAN. */
__device__ __forceinline__ double jmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double jmin(double a, double b) { return fmin(a, b); }
/* (float) x widened back to double: round-to-nearest-even */
__device__ __forceinline__ double jfloat(double x) { return (double)__double2float_rn(x); }
/* Math.abs((float) x) as double */
__device__ __forceinline__ double jabsf(double x) { return (double)fabsf(__double2float_rn(x)); }
/* (int) x for an already clamped, non-negative finite x */
__device__ __forceinline__ int jint(double x) { return __double2int_rz(x); }

/* math/Interpolation.java:77-86 */
__device__ __forceinline__ double lerp(double a, double b, double x) { return (1.0 - x) * a + x * b; }

/* Catmull-Rom weights of math/Interpolation.java:145-149.  The four polynomials do not
 * depend on the samples, so they are evaluated once per coordinate and reused for every
 * row.  Each is the reference's expression, term by term and left to right, with the
 * operations that cannot change a bit folded away; x is the fractional part of a clamped
 * coordinate, i.e. finite and in [0, 1) (a NaN still propagates through every weight):
 *   1.0*xsq == xsq;   0.0*x == +0.0 and t + (+0.0) == t for the t >= +0.0 it is added to;
 *   0.0 - 0.5*x + xsq  == xsq - 0.5*x   ((-h) + q and q - h are the same IEEE operation; for x == 0 both give +0.0);
 *   0.0 + 0.5*x        == 0.5*x         (0.5*x is +0.0 or positive);
 *   (0.0 + 0.0*x) - 0.5*xsq + 0.5*xcu == 0.5*xcu - 0.5*xsq   (again (-h) + q == q - h; zeros come out +0.0 either way). */
struct CerpW { double a, b, c, d; };
__device__ __forceinline__ CerpW cerp_weights(double x) {
    double xsq = x * x;
    double xcu = xsq * x;
    CerpW w;
    w.a = (xsq - 0.5 * x) - 0.5 * xcu;              /* 0.0 - 0.5*x + 1.0*xsq - 0.5*xcu */
    w.b = (1.0 - 2.5 * xsq) + 1.5 * xcu;            /* 1.0 + 0.0*x - 2.5*xsq + 1.5*xcu */
    w.c = (0.5 * x + 2.0 * xsq) - 1.5 * xcu;        /* 0.0 + 0.5*x + 2.0*xsq - 1.5*xcu */
    w.d = 0.5 * xcu - 0.5 * xsq;                    /* 0.0 + 0.0*x - 0.5*xsq + 0.5*xcu */
    return w;
}
/* math/Interpolation.java:135-152 */
__device__ __forceinline__ double cerp(double a, double b, double c, double d, const CerpW& w) {
    double minV = jmin(a, jmin(b, jmin(c, d)));
    double maxV = jmax(a, jmax(b, jmax(c, d)));
    double t = a * w.a + b * w.b + c * w.c + d * w.d;
    return jmin(jmax(t, minV), maxV);
}

/* ---- strip-skewed layout ------------------------------------------------------
 * Every N-vector of the pressure solve (r p z s aDiag aPlusX aPlusY precon) lives in a
 * layout made for the lexicographic recurrences.  Rows are grouped in strips of 32 and
 * columns in blocks of IFL_SKB cells; in strip k the cell (x, y = 32k + l), x = c*SKB + j,
 * is stored at   k*SS + ((c + l)*32 + l)*SKB + j.
 * One "m-row" m = c + l is what a warp (lane l = row l) processes in one wavefront step:
 * every lane owns SKB consecutive cells of its row (a register-resident chain), the m-row is
 * 32*SKB contiguous doubles, and a strip is one contiguous stream.  tl = m-rows per strip
 * (>= ceil(w/SKB) + 31). */
#define IFL_SKB 2
struct SkewGeom {
    int w, h;
    int nstrips;
    int tl;             /* m-rows per strip */
    long long ss;       /* strip stride in elements = tl * 32 * SKB */
    long long total;    /* allocated strips * ss */
    int s0, s1;         /* strips owned by this rank (slab decomposition); [0, nstrips) on a single GPU */
};
__device__ __host__ __forceinline__ long long skew_index(const SkewGeom& g, int x, int y) {
    int l = y & 31;
    int c = x / IFL_SKB, j = x - c * IFL_SKB;
    return (long long)(y >> 5) * g.ss + ((long long)(c + l) * 32 + l) * IFL_SKB + j;
}

/* non-negative float bit patterns order like unsigned ints; NaN sorts above +inf */
__device__ __forceinline__ unsigned int absf_bits(double x) {
    return __float_as_uint(fabsf(__double2float_rn(x)));
}

/* deterministic block sum for blockDim.x == 256: fixed shuffle tree, then warp totals
 * added in warp order by thread 0.  Result valid in thread 0. */
__device__ __forceinline__ double block_sum_256(double v, double* smem8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem8[wid] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        s = smem8[0];
#pragma unroll
        for (int i = 1; i < 8; i++) s += smem8[i];
    }
    __syncthreads();
    return s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

/* flag-in-data sentinel for the strip-to-strip boundary rows: a NaN payload no
 * arithmetic produces (computed NaNs are the canonical 0x7ff8000000000000 or carry an
 * input payload; no input ever holds this pattern). */
#define IFL_SENTINEL_BITS 0xFFF8B200DEAD0001ull
__device__ __forceinline__ bool is_sentinel(double v) {
    return (unsigned long long)__double_as_longlong(v) == IFL_SENTINEL_BITS;
}
__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long)IFL_SENTINEL_BITS); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned get_smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double ld_volatile(const double* p) {
    return *reinterpret_cast<const volatile double*>(p);
}


/* `v / _hx` (Euler F1:220, RungeKutta3 F2:253, backProject F4:249): when _hx = 1/min(w,h) is a power of two its reciprocal
 * is exact and v * (1/_hx) is bit-identical to the IEEE quotient (scaling by a power of two never rounds; overflow and
 * subnormal results coincide too), which spares the ~25-instruction division sequence.  Any other _hx divides. */
struct HxDiv { double hx, inv; bool pow2; };
__device__ __forceinline__ HxDiv make_hxdiv(double hx) {
    const long long b = __double_as_longlong(hx);
    HxDiv d;
    d.hx = hx;
    d.pow2 = b > 0 && (b & 0x000FFFFFFFFFFFFFll) == 0 && (b >> 52) >= 2 && (b >> 52) <= 0x7FC;
    d.inv = __longlong_as_double((0x7FEll << 52) - b);
    return d;
}
__device__ __forceinline__ double div_hx(double v, const HxDiv& d) { return d.pow2 ? v * d.inv : v / d.hx; }
} // namespace ifl
