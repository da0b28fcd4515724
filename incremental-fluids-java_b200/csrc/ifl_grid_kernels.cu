/*
 * ifl_grid_kernels.cu -- kernels on the row-major MAC-grid fields (FluidQuantity._src/_dst):
 * semi-Lagrangian advection, inflow, divergence RHS, pressure matrix, pressure
 * application, CFL reduction and the row-major <-> strip-skewed conversions.
 *
 * Reference citations are paths under /root/reference/src/physics/ (F1..F8 as in SURVEY.md).
 * Compiled with -fmad=false: no expression below may be contracted into an FMA.
 */
#include "ifl_internal.h"

namespace ifl {

struct FieldView {
    const double* __restrict__ p;
    int w, h;
    double ox, oy;
    double wm, hm;   /* _w - 1.001, _h - 1.001 as the reference evaluates them (int -> double, then subtract) */
};

__host__ static FieldView make_view(const double* p, QGeom q) {
    FieldView f;
    f.p = p; f.w = q.w; f.h = q.h; f.ox = q.ox; f.oy = q.oy;
    f.wm = (double)q.w - 1.001;
    f.hm = (double)q.h - 1.001;
    return f;
}

/* FluidQuantity.lerpGrid  F1:161 (=F2:165, F3:167, F4:207, F5:341, F7:363, F8:365) */
__device__ __forceinline__ double lerp_grid(const FieldView& f, double x, double y) {
    x = jmin(jmax(x - f.ox, 0.0), f.wm);
    y = jmin(jmax(y - f.oy, 0.0), f.hm);
    int ix = jint(x);
    int iy = jint(y);
    x -= ix;
    y -= iy;
    const double* row0 = f.p + (size_t)iy * f.w + ix;
    const double* row1 = row0 + f.w;
    double x00 = __ldg(row0), x10 = __ldg(row0 + 1);
    double x01 = __ldg(row1), x11 = __ldg(row1 + 1);
    return lerp(lerp(x00, x10, x), lerp(x01, x11, x), y);
}

/* FluidQuantity.cerpGrid  F2:183 (=F3:185, F4:225, F5:359, F7:381) */
__device__ __forceinline__ double cerp_grid(const FieldView& f, double x, double y) {
    x = jmin(jmax(x - f.ox, 0.0), f.wm);
    y = jmin(jmax(y - f.oy, 0.0), f.hm);
    int ix = jint(x);
    int iy = jint(y);
    x -= ix;
    y -= iy;
    int x0 = max(ix - 1, 0), x1 = ix, x2 = ix + 1, x3 = min(ix + 2, f.w - 1);
    int y0 = max(iy - 1, 0), y1 = iy, y2 = iy + 1, y3 = min(iy + 2, f.h - 1);
    CerpW wx = cerp_weights(x);
    const double* r0 = f.p + (size_t)y0 * f.w;
    const double* r1 = f.p + (size_t)y1 * f.w;
    const double* r2 = f.p + (size_t)y2 * f.w;
    const double* r3 = f.p + (size_t)y3 * f.w;
    double q0 = cerp(__ldg(r0 + x0), __ldg(r0 + x1), __ldg(r0 + x2), __ldg(r0 + x3), wx);
    double q1 = cerp(__ldg(r1 + x0), __ldg(r1 + x1), __ldg(r1 + x2), __ldg(r1 + x3), wx);
    double q2 = cerp(__ldg(r2 + x0), __ldg(r2 + x1), __ldg(r2 + x2), __ldg(r2 + x3), wx);
    double q3 = cerp(__ldg(r3 + x0), __ldg(r3 + x1), __ldg(r3 + x2), __ldg(r3 + x3), wx);
    CerpW wy = cerp_weights(y);
    return cerp(q0, q1, q2, q3, wy);
}

/* FluidQuantity.advect: F1:178 (Euler F1:220 + lerpGrid) and F2:205 / F3:207
 * (RungeKutta3 F2:253 / F3:255 + cerpGrid).  One thread per destination sample. */
template <bool RK3>
__global__ void __launch_bounds__(256)
k_advect(FieldView q, FieldView u, FieldView v, double* __restrict__ dst, double timestep, double hx, int row0, int row1) {
    const HxDiv hd = make_hxdiv(hx);
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    int iy = row0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= q.w || iy >= row1) return;
    double x = ix + q.ox;
    double y = iy + q.oy;
    double out;
    if (!RK3) {
        double uVel = div_hx(lerp_grid(u, x, y), hd);
        double vVel = div_hx(lerp_grid(v, x, y), hd);
        x -= uVel * timestep;
        y -= vVel * timestep;
        out = lerp_grid(q, x, y);
    } else {
        double firstU = div_hx(lerp_grid(u, x, y), hd);
        double firstV = div_hx(lerp_grid(v, x, y), hd);
        double midX = x - 0.5 * timestep * firstU;
        double midY = y - 0.5 * timestep * firstV;
        double midU = div_hx(lerp_grid(u, midX, midY), hd);
        double midV = div_hx(lerp_grid(v, midX, midY), hd);
        double lastX = x - 0.75 * timestep * midU;
        double lastY = y - 0.75 * timestep * midV;
        /* F2:266-267: the third sample is NOT divided by _hx (port quirk, kept) */
        double lastU = lerp_grid(u, lastX, lastY);
        double lastV = lerp_grid(v, lastX, lastY);
        x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
        y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
        out = cerp_grid(q, x, y);
    }
    dst[(size_t)iy * q.w + ix] = out;
}

/* rows [row0, row1) of the destination (the whole field on a single GPU, the rows a rank owns in a slab decomposition) */
void launch_advect(cudaStream_t st, int variant, QGeom q, const double* src, double* dst,
                   QGeom qu, const double* u, QGeom qv, const double* v, double timestep, double hx, int row0, int row1) {
    if (row1 <= row0) return;
    dim3 block(32, 8);
    dim3 grid((q.w + 31) / 32, (row1 - row0 + 7) / 8);
    FieldView fq = make_view(src, q), fu = make_view(u, qu), fv = make_view(v, qv);
    if (variant == IFL_F1_MATRIXLESS)
        k_advect<false><<<grid, block, 0, st>>>(fq, fu, fv, dst, timestep, hx, row0, row1);
    else
        k_advect<true><<<grid, block, 0, st>>>(fq, fu, fv, dst, timestep, hx, row0, row1);
}

/* FluidQuantity.addInflow: F1:202 (hard rectangle) / F2:230 (=F3:232 ...) smooth pulse.
 * The rectangle bounds are computed on the host exactly as the reference does. */
__device__ __forceinline__ double cubic_pulse(double x) {   /* F2:107 */
    x = jmin(jabsf(x), 1.0);
    return 1.0 - x * x * (3.0 - 2.0 * x);
}
__global__ void k_add_inflow(double* __restrict__ src, int w, size_t total, int xlo, int xhi, int ylo, int yhi, int smooth,
                             double hx, double x0, double y0, double x1, double y1, double v) {
    int x = xlo + blockIdx.x * blockDim.x + threadIdx.x;
    int y = ylo + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= xhi || y >= yhi) return;
    size_t i = (size_t)x + (size_t)y * w;
    /* the x range is clipped with _h (F1:209, kept): on a grid with w < h an index can run into the next row, like in the
     * Java array, but never past the end of the buffer (where the reference throws ArrayIndexOutOfBoundsException) */
    if (i >= total) return;
    double vi = v;
    if (smooth) {
        double lx = (2.0 * (x + 0.5) * hx - (x0 + x1)) / (x1 - x0);
        double ly = (2.0 * (y + 0.5) * hx - (y0 + y1)) / (y1 - y0);
        double l = sqrt(lx * lx + ly * ly);
        vi = cubic_pulse(l) * v;
    }
    if (fabsf(__double2float_rn(src[i])) < fabsf(__double2float_rn(vi))) src[i] = vi;
}

static int jint_host(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int)x;
}

void launch_add_inflow(cudaStream_t st, int variant, QGeom q, double* src, double hx,
                       double x0, double y0, double x1, double y1, double value) {
    int ix0 = jint_host(x0 / hx - q.ox);
    int iy0 = jint_host(y0 / hx - q.oy);
    int ix1 = jint_host(x1 / hx - q.ox);
    int iy1 = jint_host(y1 / hx - q.oy);
    int ylo = iy0 > 0 ? iy0 : 0, yhi = iy1 < q.h ? iy1 : q.h;
    /* F1:209: the x range is clipped with _h, not _w (kept) */
    int xlo = ix0 > 0 ? ix0 : 0, xhi = ix1 < q.h ? ix1 : q.h;
    if (xhi <= xlo || yhi <= ylo) return;
    dim3 block(32, 8);
    dim3 grid((xhi - xlo + 31) / 32, (yhi - ylo + 7) / 8);
    k_add_inflow<<<grid, block, 0, st>>>(src, q.w, (size_t)q.w * q.h, xlo, xhi, ylo, yhi, variant != IFL_F1_MATRIXLESS,
                                         hx, x0, y0, x1, y1, value);
}

/* ---- row-major <-> skewed helpers ------------------------------------------------
 * Kernels that touch both layouts run one thread per cell (x, y) in row-major order and address the
 * skewed side through skew_index(); they run once per step, so their strided skewed accesses are noise. */
struct SkewPos { int x, y; long long i; bool ok; };
__device__ __forceinline__ SkewPos skew_pos(const SkewGeom& g, int row0 = 0, int row1 = 0x7fffffff) {
    SkewPos s;
    s.x = blockIdx.x * blockDim.x + threadIdx.x;
    s.y = row0 + blockIdx.y * blockDim.y + threadIdx.y;
    s.ok = s.x < g.w && s.y < g.h && s.y < row1;
    s.i = s.ok ? skew_index(g, s.x, s.y) : 0;
    return s;
}
static dim3 skew_grid(const SkewGeom& g) { return dim3((g.w + 31) / 32, (g.h + 7) / 8); }
static dim3 rows_grid(const SkewGeom& g, int row0, int row1) { return dim3((g.w + 31) / 32, (row1 - row0 + 7) / 8); }
/* the rows of the strips [s0, s1) a rank owns */
static inline void own_rows(const SkewGeom& g, int* row0, int* row1) {
    *row0 = g.s0 * 32;
    *row1 = g.s1 * 32 < g.h ? g.s1 * 32 : g.h;
}

/* FluidSolver.buildRhs  F1:364 (=F2:379, F3:420) */
__global__ void k_build_rhs(SkewGeom g, const double* __restrict__ u, const double* __restrict__ v,
                            double* __restrict__ r, double scale, int row0, int row1) {
    SkewPos s = skew_pos(g, row0, row1);
    if (!s.ok) return;
    int x = s.x, y = s.y, w = g.w;
    double u1 = u[(size_t)(x + 1) + (size_t)y * (w + 1)], u0 = u[(size_t)x + (size_t)y * (w + 1)];
    double v1 = v[(size_t)x + (size_t)(y + 1) * w], v0 = v[(size_t)x + (size_t)y * w];
    r[s.i] = -scale * (u1 - u0 + v1 - v0);
}
void launch_build_rhs(cudaStream_t st, const SkewGeom& g, const double* u, const double* v, double* r, double hx) {
    int row0, row1;
    own_rows(g, &row0, &row1);          /* the whole grid on a single GPU */
    k_build_rhs<<<rows_grid(g, row0, row1), dim3(32, 8), 0, st>>>(g, u, v, r, 1.0 / hx, row0, row1);
}

/* FluidSolver.buildPressureMatrix F3:435 in gather form.  The reference scatters in
 * lexicographic order, so aDiag[idx] receives its terms in the order up, left, right, down. */
__global__ void k_build_matrix_f3(SkewGeom g, double* __restrict__ adiag, double* __restrict__ aplusx,
                                  double* __restrict__ aplusy, double scale) {
    SkewPos s = skew_pos(g);
    if (!s.ok) return;
    double d = 0.0;
    if (s.y > 0) d += scale;
    if (s.x > 0) d += scale;
    if (s.x < g.w - 1) d += scale;
    if (s.y < g.h - 1) d += scale;
    adiag[s.i] = d;
    aplusx[s.i] = (s.x < g.w - 1) ? -scale : 0.0;
    aplusy[s.i] = (s.y < g.h - 1) ? -scale : 0.0;
}
void launch_build_matrix_f3(cudaStream_t st, const SkewGeom& g, double* adiag, double* aplusx, double* aplusy, double scale) {
    k_build_matrix_f3<<<skew_grid(g), dim3(32, 8), 0, st>>>(g, adiag, aplusx, aplusy, scale);
}

/* FluidSolver.applyPressure F1:432 (=F2:447, F3:634) in gather form: per face the
 * reference first adds scale*p of the cell before it, then subtracts scale*p of its own
 * cell; afterwards the four domain edges are zeroed (F1:444-451). */
__global__ void k_apply_pressure(SkewGeom g, const double* __restrict__ p, double* __restrict__ u,
                                 double* __restrict__ v, double scale, int row0, int row1) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = row0 + blockIdx.y * blockDim.y + threadIdx.y;
    int w = g.w, h = g.h;
    if (x > w || y >= row1) return;
    double pc = 0.0, pl = 0.0, pu = 0.0;
    if (x < w && y < h) pc = p[skew_index(g, x, y)];
    if (y < h) {  /* u face (x, y), x in [0, w] */
        size_t iu = (size_t)x + (size_t)y * (w + 1);
        double val;
        if (x == 0 || x == w) val = 0.0;
        else {
            pl = p[skew_index(g, x - 1, y)];
            val = (u[iu] + scale * pl) - scale * pc;
        }
        u[iu] = val;
    }
    if (x < w) {  /* v face (x, y), y in [0, h] */
        size_t iv = (size_t)x + (size_t)y * w;
        double val;
        if (y == 0 || y == h) val = 0.0;
        else {
            pu = p[skew_index(g, x, y - 1)];
            val = (v[iv] + scale * pu) - scale * pc;
        }
        v[iv] = val;
    }
}
/* faces of the rows a rank owns: u rows [row0, row1), v rows [row0, row1), and v row h on the rank that owns row h-1
 * (needs p of row row0-1: the last row of the previous rank, see pcg_exchange_pressure_halo) */
void launch_apply_pressure(cudaStream_t st, const SkewGeom& g, const double* p, double* u, double* v, double scale) {
    int row0, row1;
    own_rows(g, &row0, &row1);
    if (row1 == g.h) row1 = g.h + 1;
    dim3 block(32, 8);
    dim3 grid((g.w + 1 + 31) / 32, (row1 - row0 + 7) / 8);
    k_apply_pressure<<<grid, block, 0, st>>>(g, p, u, v, scale, row0, row1);
}

/* rows [row0, row1) of a skewed vector <-> a row-major buffer whose first row is row0 */
__global__ void k_skew_to_rows(SkewGeom g, const double* __restrict__ skew, double* __restrict__ rows, int row0, int row1) {
    SkewPos s = skew_pos(g, row0, row1);
    if (!s.ok) return;
    rows[(size_t)s.x + (size_t)(s.y - row0) * g.w] = skew[s.i];
}
__global__ void k_rows_to_skew(SkewGeom g, const double* __restrict__ rows, double* __restrict__ skew, int row0, int row1) {
    SkewPos s = skew_pos(g, row0, row1);
    if (!s.ok) return;
    skew[s.i] = rows[(size_t)s.x + (size_t)(s.y - row0) * g.w];
}
void launch_skew_to_rows(cudaStream_t st, const SkewGeom& g, const double* skew, double* rows, int row0, int row1) {
    if (row1 < 0) { row0 = 0; row1 = g.h; }
    if (row1 <= row0) return;
    k_skew_to_rows<<<rows_grid(g, row0, row1), dim3(32, 8), 0, st>>>(g, skew, rows, row0, row1);
}
void launch_rows_to_skew(cudaStream_t st, const SkewGeom& g, const double* rows, double* skew, int row0, int row1) {
    if (row1 < 0) { row0 = 0; row1 = g.h; }
    if (row1 <= row0) return;
    k_rows_to_skew<<<rows_grid(g, row0, row1), dim3(32, 8), 0, st>>>(g, rows, skew, row0, row1);
}

/* FluidSolver.maxTimestep F1:309: max over cells of |(u,v)| sampled at cell centres.
 * max is order independent, so an atomicMax on the bit pattern is exact. */
__global__ void k_max_velocity(FieldView u, FieldView v, int w, int h, unsigned long long* out_bits) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    double vel = 0.0;
    if (x < w && y < h) {
        double uu = lerp_grid(u, x + 0.5, y + 0.5);
        double vv = lerp_grid(v, x + 0.5, y + 0.5);
        vel = sqrt(uu * uu + vv * vv);
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(vel);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) atomicMax(out_bits, bits);
}
/* slab decomposition: max |v| over the v faces of rows [row0, row1) -- the CFL number that sizes the advection halo
 * (cf. maxTimestep F1:309-333; only the vertical component moves a backtrace across slab boundaries) */
__global__ void k_max_abs_rows(const double* __restrict__ v, int w, int row0, int row1, unsigned long long* out_bits) {
    const long long n = (long long)(row1 - row0) * w;
    const double* p = v + (long long)row0 * w;
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmax(m, fabs(p[i]));
    unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, bits);
}
void launch_max_abs_rows(cudaStream_t st, const double* v, int w, int row0, int row1, unsigned long long* out_bits) {
    if (row1 <= row0) return;
    k_max_abs_rows<<<592, 256, 0, st>>>(v, w, row0, row1, out_bits);
}

/* FluidSolver.toImage: the integers of one frame (see include/incfluid_b200.h).  One thread per cell; with render_heat
 * the frame is 2w wide, heat panel at x, density panel at x + w (idxl / idxr of F6:1263-1264). */
__global__ void k_to_image(int variant, int w, int h, const double* __restrict__ d, const double* __restrict__ t,
                           const double* __restrict__ volume, const unsigned char* __restrict__ cell,
                           double t_ambient, int render_heat, int* __restrict__ rgb) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t i = (size_t)x + (size_t)y * w;
    if (variant < IFL_F6_HEAT) {
        double v = 1.0 - d[i];
        if (variant == IFL_F5_CURVED_BOUNDARIES) v = v * volume[i];          /* F5:1142 */
        int shade = jint(v * 255.0);                                         /* F1:351 */
        shade = max(min(shade, 255), 0);                                     /* F1:352 */
        if (variant == IFL_F4_SOLID_BOUNDARIES && cell[i] == IFL_CELL_SOLID) shade = 0;   /* F4:678 */
        rgb[3 * i + 0] = shade; rgb[3 * i + 1] = shade; rgb[3 * i + 2] = shade;
        return;
    }
    const size_t fw = render_heat ? 2 * (size_t)w : (size_t)w;
    const size_t ir = 3 * ((size_t)x + (size_t)y * fw + (render_heat ? (size_t)w : 0));
    const double vol = volume[i];
    double shade = (1.0 - d[i]) * vol;                                       /* F6:1271 */
    shade = jmin(jmax(shade, 0.0), 1.0);
    int c = jint(shade * 255.0);
    int r = c, g = c, b = c;
    if (variant == IFL_F8_FLIP && cell[i] == IFL_CELL_EMPTY) { r = 0xFF; g = 0; b = 0; }   /* F8:1444 */
    rgb[ir + 0] = r; rgb[ir + 1] = g; rgb[ir + 2] = b;
    if (render_heat) {
        double tt;
        if (variant == IFL_F8_FLIP) tt = fabs(jfloat(t[i]) - t_ambient) / 70.0;      /* F8:1450: (float) binds to the sample only */
        else tt = (t[i] - t_ambient) / 700.0;                                        /* F6:1279 */
        tt = jmin(jmax(tt, 0.0), 1.0);
        const double hr = 1.0 + vol * (jmin(tt * 4.0, 1.0) - 1.0);
        const double hg = 1.0 + vol * (jmin(tt * 2.0, 1.0) - 1.0);
        const double hb = 1.0 + vol * (jmax(jmin(tt * 4.0 - 3.0, 1.0), 0.0) - 1.0);
        const size_t il = 3 * ((size_t)x + (size_t)y * fw);
        rgb[il + 0] = jint(hr * 255.0); rgb[il + 1] = jint(hg * 255.0); rgb[il + 2] = jint(hb * 255.0);
    }
}
void launch_to_image(cudaStream_t st, int variant, int w, int h, const double* d, const double* t, const double* volume,
                     const unsigned char* cell, double t_ambient, int render_heat, int* rgb) {
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    k_to_image<<<grid, block, 0, st>>>(variant, w, h, d, t, volume, cell, t_ambient, render_heat, rgb);
}

void launch_max_velocity(cudaStream_t st, int w, int h, const double* u, const double* v, unsigned long long* out_bits) {
    QGeom qu{w + 1, h, 0.0, 0.5}, qv{w, h + 1, 0.5, 0.0};
    dim3 block(32, 8);
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_max_velocity<<<grid, block, 0, st>>>(make_view(u, qu), make_view(v, qv), w, h, out_bits);
}

} // namespace ifl
