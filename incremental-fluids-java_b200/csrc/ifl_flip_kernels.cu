/*
 * ifl_flip_kernels.cu -- F8 (hybridFluids/ExplicitFluid8flip.java): FLIP/PIC particle <-> grid transfers,
 * particle advection, particle bookkeeping (init / count / prune / seed) and the F8 flavour of
 * FluidQuantity.extrapolate (CELL_EMPTY cells).
 *
 * Order-dependent pieces of the reference and how they are reproduced exactly:
 *   - fromParticles (F8:722) adds samples in particle-index order.  Particles are binned by their base cell with a
 *     STABLE radix sort (cub), and every grid node merges the (index-sorted) lists of the four base cells that
 *     touch it, adding the samples in increasing particle index -- the reference's order, bit for bit.
 *   - Math.random() (F8:1576,1647) is replaced by an injected counter-based stream rng_uniform(seed, k)
 *     (documented in include/incfluid_b200.h), so init and seeding consume "the k-th draw" in the reference's
 *     order while running in parallel (prefix sums give every attempt its draw index and its output slot).
 *   - pruneParticles (F8:1610) is a sequential swap-with-last scan.  Only particles in over-full cells can be
 *     removed, so the scan is replayed by one thread over the compacted list of those particles.
 *   - extrapolate (F8:651) pops a LIFO stack and relabels EMPTY cells FLUID as it goes; that only matters when two
 *     EMPTY cells are 4-adjacent.  Without such a pair a level-parallel sweep is exact; with one, a single-thread
 *     kernel replays the reference's stack literally.
 * Compiled with -fmad=false.
 */
#include <cub/cub.cuh>
#include "ifl_internal.h"

namespace ifl {

__device__ __forceinline__ float rng_uniform(unsigned long long seed, unsigned long long k) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (k + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

struct GridView {
    const double* __restrict__ p;
    int w, h;
    double ox, oy, wm, hm;
};
__host__ static GridView grid_view(const double* p, QGeom q) {
    GridView f; f.p = p; f.w = q.w; f.h = q.h; f.ox = q.ox; f.oy = q.oy;
    f.wm = (double)q.w - 1.001; f.hm = (double)q.h - 1.001;
    return f;
}
__device__ __forceinline__ double g_lerp_grid(const GridView& f, double x, double y) {       /* F8:365 */
    x = jmin(jmax(x - f.ox, 0.0), f.wm);
    y = jmin(jmax(y - f.oy, 0.0), f.hm);
    int ix = jint(x), iy = jint(y);
    x -= ix; y -= iy;
    const double* r0 = f.p + (size_t)iy * f.w + ix;
    const double* r1 = r0 + f.w;
    return lerp(lerp(r0[0], r0[1], x), lerp(r1[0], r1[1], x), y);
}

/* bodies: only distance / closest point / normal are needed here (same arithmetic as ifl_solid_kernels.cu) */
struct V2f { double x, y; };
__device__ __forceinline__ double f_signum(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
__device__ __forceinline__ V2f f_rotate(const ifl_body& b, double x, double y, int dir) {
    const double c = b.cos_theta, sn = dir > 0 ? b.sin_theta : -b.sin_theta;
    V2f r; r.x = c * x + sn * y; r.y = -sn * x + c * y;
    return r;
}
__device__ __forceinline__ double f_nsgn(double v) { double r = f_signum(v); return fabs(r) == 0 ? 0.0 : r; }
__device__ __forceinline__ double f_distance(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {
        x -= b.pos_x; y -= b.pos_y;
        V2f r = f_rotate(b, x, y, -1);
        double dx = fabs(r.x) - b.scale_x * 0.5;
        double dy = fabs(r.y) - b.scale_y * 0.5;
        if (dx >= 0.0 || dy >= 0.0) { double mx = jmax(dx, 0.0), my = jmax(dy, 0.0); return sqrt(mx * mx + my * my); }
        return jmax(dx, dy);
    }
    double ex = x - b.pos_x, ey = y - b.pos_y;
    return sqrt(ex * ex + ey * ey) - b.scale_x * 0.5;
}
__device__ __forceinline__ V2f f_closest(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {
        x -= b.pos_x; y -= b.pos_y;
        V2f r = f_rotate(b, x, y, -1);
        x = r.x; y = r.y;
        double dx = fabs(x) - b.scale_x * 0.5;
        double dy = fabs(y) - b.scale_y * 0.5;
        if (dx > dy) x = f_nsgn(x) * 0.5 * b.scale_x; else y = f_nsgn(y) * 0.5 * b.scale_y;
        r = f_rotate(b, x, y, +1);
        r.x = r.x + b.pos_x; r.y = r.y + b.pos_y;
        return r;
    }
    x -= b.pos_x; y -= b.pos_y;
    V2f r = f_rotate(b, x, y, -1);
    x = r.x / b.scale_x; y = r.y / b.scale_y;
    double rr = sqrt(x * x + y * y);
    if (rr < 1e-4) { x = 0.5; y = 0.0; } else { x /= 2.0 * rr; y /= 2.0 * rr; }
    x *= b.scale_x; y *= b.scale_y;
    r = f_rotate(b, x, y, +1);
    r.x = r.x + b.pos_x; r.y = r.y + b.pos_y;
    return r;
}
__device__ __forceinline__ V2f f_normal(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {
        x -= b.pos_x; y -= b.pos_y;
        V2f r = f_rotate(b, x, y, -1);
        x = r.x; y = r.y;
        double nx, ny;
        if (fabs(x) - b.scale_x * 0.5 > fabs(y) - b.scale_y * 0.5) { nx = f_nsgn(x); ny = 0.0; } else { nx = 0.0; ny = f_nsgn(y); }
        return f_rotate(b, nx, ny, +1);
    }
    x -= b.pos_x; y -= b.pos_y;
    double rr = sqrt(x * x + y * y);
    V2f n;
    if (rr < 1e-4) { n.x = 1.0; n.y = 0.0; } else { n.x = x / rr; n.y = y / rr; }
    return n;
}
__device__ __forceinline__ bool point_in_body(const ifl_body* bodies, int nbodies, double x, double y, double hx) {   /* F8:1564 */
    for (int i = 0; i < nbodies; i++) if (f_distance(bodies[i], x * hx, y * hx) < 0.0) return true;
    return false;
}

/* ---- initParticles F8:1571 ------------------------------------------------------------------ */
__global__ void k_init_accept(int w, int h, int avg, unsigned long long seed, unsigned long long rng0,
                              const ifl_body* __restrict__ bodies, int nbodies, double hx, int* __restrict__ accept) {
    long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)w * h * avg;
    if (a >= total) return;
    int cell = (int)(a / avg);
    int x = cell % w, y = cell / w;
    double px = (double)((float)x + rng_uniform(seed, rng0 + 2 * a));       /* x + (float) random(): a float addition */
    double py = (double)((float)y + rng_uniform(seed, rng0 + 2 * a + 1));
    accept[a] = point_in_body(bodies, nbodies, px, py, hx) ? 0 : 1;
}
__global__ void k_init_write(int w, int h, int avg, unsigned long long seed, unsigned long long rng0,
                             const int* __restrict__ accept, const int* __restrict__ slot,
                             double* __restrict__ posX, double* __restrict__ posY) {
    long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)w * h * avg;
    if (a >= total || !accept[a]) return;
    int cell = (int)(a / avg);
    int x = cell % w, y = cell / w;
    int j = slot[a];
    posX[j] = (double)((float)x + rng_uniform(seed, rng0 + 2 * a));
    posY[j] = (double)((float)y + rng_uniform(seed, rng0 + 2 * a + 1));
}

/* ---- gridToParticles F8:1748 ---------------------------------------------------------------- */
__global__ void k_grid_to_particles(int n, double alpha, const double* __restrict__ posX, const double* __restrict__ posY,
                                    GridView gd, GridView gt, GridView gu, GridView gv,
                                    double* __restrict__ pd, double* __restrict__ pt, double* __restrict__ pu, double* __restrict__ pv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = posX[i], y = posY[i];
    const double keep = 1.0 - alpha;
    double v;
    v = pd[i]; v *= keep; v += g_lerp_grid(gd, x, y); pd[i] = v;
    v = pt[i]; v *= keep; v += g_lerp_grid(gt, x, y); pt[i] = v;
    v = pu[i]; v *= keep; v += g_lerp_grid(gu, x, y); pu[i] = v;
    v = pv[i]; v *= keep; v += g_lerp_grid(gv, x, y); pv[i] = v;
}

/* ---- particle advect F8:1778 (rungeKutta3 F8:1708, backProject F8:1673) ------------------------ */
__global__ void k_advect_particles(int n, double timestep, double hx, int w, int h, GridView gu, GridView gv,
                                   const ifl_body* __restrict__ bodies, int nbodies,
                                   double* __restrict__ posX, double* __restrict__ posY) {
    const HxDiv hd = make_hxdiv(hx);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = posX[i], y = posY[i];
    double firstU = div_hx(g_lerp_grid(gu, x, y), hd);
    double firstV = div_hx(g_lerp_grid(gv, x, y), hd);
    double midX = x + 0.5 * timestep * firstU;
    double midY = y + 0.5 * timestep * firstV;
    double midU = div_hx(g_lerp_grid(gu, midX, midY), hd);
    double midV = div_hx(g_lerp_grid(gv, midX, midY), hd);
    double lastX = x + 0.75 * timestep * midU;
    double lastY = y + 0.75 * timestep * midV;
    double lastU = g_lerp_grid(gu, lastX, lastY);
    double lastV = g_lerp_grid(gv, lastX, lastY);
    x += timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
    y += timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
    double dd = 1e30;
    int closest = -1;
    for (int b = 0; b < nbodies; b++) {
        double id = f_distance(bodies[b], x * hx, y * hx);
        if (id < dd) { dd = id; closest = b; }
    }
    if (dd < -1.0) {                                  /* F8:1685: threshold -1.0 world units (kept) */
        x *= hx; y *= hx;
        V2f r = f_closest(bodies[closest], x, y);
        x = r.x; y = r.y;
        V2f nn = f_normal(bodies[closest], x, y);
        x -= nn.x * hx; y -= nn.y * hx;
        x /= hx; y /= hx;
    }
    posX[i] = jmax(jmin(x, w - 0.001), 0.0);
    posY[i] = jmax(jmin(y, h - 0.001), 0.0);
}

/* ---- fromParticles F8:722 / addSample F8:312 -------------------------------------------------- */
__global__ void k_p2g_keys(int n, const double* __restrict__ posX, const double* __restrict__ posY, int w, int h,
                           double ox, double oy, int* __restrict__ keys, int* __restrict__ vals, int* __restrict__ hist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = jmax(0.5, jmin(w - 1.5, posX[i] - ox));
    double y = jmax(0.5, jmin(h - 1.5, posY[i] - oy));
    int key = jint(x) + jint(y) * w;
    keys[i] = key;
    vals[i] = i;
    atomicAdd(hist + key, 1);
}
/* one thread per grid node: merge the index-sorted particle lists of the four base cells touching the node */
__global__ void k_p2g_gather(int w, int h, double ox, double oy, const int* __restrict__ start, const int* __restrict__ sorted,
                             const double* __restrict__ posX, const double* __restrict__ posY, const double* __restrict__ prop,
                             double* __restrict__ src, unsigned char* __restrict__ cell) {
    int gx = blockIdx.x * blockDim.x + threadIdx.x;
    int gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= w || gy >= h) return;
    int cur[4], end[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int bx = gx - (c & 1), by = gy - (c >> 1);
        if (bx >= 0 && by >= 0 && bx < w && by < h) { int key = bx + by * w; cur[c] = start[key]; end[c] = start[key + 1]; }
        else { cur[c] = 0; end[c] = 0; }
    }
    double weight = 0.0, acc = 0.0;
    const double wlim = w - 1.5, hlim = h - 1.5;
    for (;;) {
        int best = -1, bi = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (cur[c] < end[c]) { int pi = sorted[cur[c]]; if (pi < bi) { bi = pi; best = c; } }
        if (best < 0) break;
#pragma unroll
        for (int c = 0; c < 4; c++) if (c == best) cur[c]++;
        double x = jmax(0.5, jmin(wlim, posX[bi] - ox));
        double y = jmax(0.5, jmin(hlim, posY[bi] - oy));
        double k = (1.0 - fabs(gx - x)) * (1.0 - fabs(gy - y));
        weight += k;
        acc += k * prop[bi];
    }
    size_t idx = (size_t)gx + (size_t)gy * w;
    if (weight != 0.0) src[idx] = acc / weight;
    else {
        src[idx] = acc;                                  /* fresh array: 0.0 */
        if (cell[idx] == IFL_CELL_FLUID) cell[idx] = IFL_CELL_EMPTY;
    }
}

/* ---- copy / diff / undiff F8:339 / 346 / 355 ------------------------------------------------------ */
__global__ void k_diff(double* __restrict__ src, const double* __restrict__ old, size_t n, double keep, int sign) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sign < 0) src[i] -= keep * old[i]; else src[i] += keep * old[i];
}

/* ---- countParticles F8:1595 ------------------------------------------------------------------------ */
__global__ void k_count_particles(int n, const double* __restrict__ posX, const double* __restrict__ posY, int w, int h,
                                  int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ix = jint(posX[i]), iy = jint(posY[i]);
    if (ix >= 0 && iy >= 0 && ix < w && iy < h) atomicAdd(counts + ix + iy * w, 1);
}
/* ---- pruneParticles F8:1610 ------------------------------------------------------------------------- */
__global__ void k_prune_flags(int n, const double* __restrict__ posX, const double* __restrict__ posY, int w,
                              const int* __restrict__ counts, int maxpc, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = counts[jint(posX[i]) + jint(posY[i]) * w] > maxpc ? 1 : 0;
}
__global__ void k_iota(int n, int* v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
/* replay of the reference's scan by ONE thread over the (ascending) list of particles in over-full cells */
__global__ void k_prune_sequential(const int* __restrict__ list, const int* __restrict__ nlist, int* count, int w, int maxpc,
                                   int* counts, double* posX, double* posY, double* pd, double* pt, double* pu, double* pv) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = *count;
    const int m = *nlist;
    for (int li = 0; li < m; li++) {
        const int i = list[li];
        if (i >= n) break;
        while (i < n) {
            int c = jint(posX[i]) + jint(posY[i]) * w;
            if (counts[c] > maxpc) {
                int j = --n;
                posX[i] = posX[j]; posY[i] = posY[j];
                pd[i] = pd[j]; pt[i] = pt[j]; pu[i] = pu[j]; pv[i] = pv[j];
                counts[c]--;
            } else break;
        }
    }
    *count = n;
}
/* ---- seedParticles F8:1637 ---------------------------------------------------------------------------- */
__global__ void k_seed_need(int ncell, const int* __restrict__ counts, int minpc, int* __restrict__ need) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int nd = minpc - counts[c];
    need[c] = nd > 0 ? nd : 0;
}
__global__ void k_seed_accept(int ncell, const int* __restrict__ need, const int* __restrict__ off,
                              const double* __restrict__ posX, const double* __restrict__ posY,
                              const ifl_body* __restrict__ bodies, int nbodies, double hx, int* __restrict__ accept) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell || need[c] == 0) return;
    /* F8:1653 tests the particle whose INDEX is the cell index, not the new particle (kept) */
    int ok = point_in_body(bodies, nbodies, posX[c], posY[c], hx) ? 0 : 1;
    for (int i = 0; i < need[c]; i++) accept[off[c] + i] = ok;
}
__global__ void k_seed_write(int ncell, int w, const int* __restrict__ need, const int* __restrict__ off,
                             const int* __restrict__ accept, const int* __restrict__ slot, int n0, int maxp,
                             unsigned long long seed, unsigned long long rng0,
                             double* posX, double* posY, GridView gd, GridView gt, GridView gu, GridView gv,
                             double* pd, double* pt, double* pu, double* pv, int* executed, int* accepted) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell || need[c] == 0) return;
    int x = c % w, y = c / w;
    for (int i = 0; i < need[c]; i++) {
        int a = off[c] + i;
        int j = n0 + slot[a];
        if (j >= maxp) break;                            /* F8:1642: return once the particle store is full */
        atomicAdd(executed, 1);
        if (!accept[a]) continue;
        double px = (double)((float)x + rng_uniform(seed, rng0 + 2ull * a));
        double py = (double)((float)y + rng_uniform(seed, rng0 + 2ull * a + 1));
        posX[j] = px; posY[j] = py;
        pd[j] = g_lerp_grid(gd, px, py);
        pt[j] = g_lerp_grid(gt, px, py);
        pu[j] = g_lerp_grid(gu, px, py);
        pv[j] = g_lerp_grid(gv, px, py);
        atomicAdd(accepted, 1);
    }
}

/* ---- extrapolate, F8 flavour (F8:483-697) --------------------------------------------------------------- */
__global__ void k_flip_fill_mask(SolidQ q, unsigned char* __restrict__ state, int* adjacent_empty) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= q.w || y >= q.h) return;
    size_t idx = (size_t)x + (size_t)y * q.w;
    const int w = q.w;
    if (x == 0 || y == 0 || x == q.w - 1 || y == q.h - 1) { q.mask[idx] = 0xFF; state[idx] = 0; return; }
    int m = 0;
    unsigned char c = q.cell[idx];
    if (c == IFL_CELL_SOLID) {
        double nx = q.normalX[idx], ny = q.normalY[idx];
        if (nx != 0.0 && q.cell[idx + (int)f_signum(nx)] != IFL_CELL_FLUID) m |= 1;
        if (ny != 0.0 && q.cell[idx + (int)f_signum(ny) * w] != IFL_CELL_FLUID) m |= 2;
    } else if (c == IFL_CELL_EMPTY) {
        unsigned char l = q.cell[idx - 1], r = q.cell[idx + 1], u = q.cell[idx - w], d = q.cell[idx + w];
        m = (l != IFL_CELL_FLUID && r != IFL_CELL_FLUID && u != IFL_CELL_FLUID && d != IFL_CELL_FLUID) ? 1 : 0;
        /* interior EMPTY neighbours make the reference's result depend on its stack order */
        bool adj = (x > 1 && l == IFL_CELL_EMPTY) || (x < q.w - 2 && r == IFL_CELL_EMPTY) ||
                   (y > 1 && u == IFL_CELL_EMPTY) || (y < q.h - 2 && d == IFL_CELL_EMPTY);
        if (adj) *adjacent_empty = 1;
    }
    q.mask[idx] = m;
    state[idx] = c != IFL_CELL_FLUID ? 1 : 0;
}
__device__ __forceinline__ double flip_extrap_normal(const SolidQ& q, const double* src, size_t idx) {   /* F8:530 */
    double nx = q.normalX[idx], ny = q.normalY[idx];
    double srcX = __ldcg(src + idx + (int)f_signum(nx));
    double srcY = __ldcg(src + idx + (int)f_signum(ny) * q.w);
    return (jabsf(nx) * srcX + jabsf(ny) * srcY) / (jabsf(nx) + jabsf(ny));
}
/* level-parallel sweep; exact when no two interior EMPTY cells are adjacent (labels are switched at the end) */
__global__ void k_flip_extrap_sweep(SolidQ q, double* src, unsigned char* state, unsigned int* progress) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    bool did = false;
    const int w = q.w;
    if (x >= 1 && x < q.w - 1 && y >= 1 && y < q.h - 1) {
        size_t idx = (size_t)x + (size_t)y * w;
        if (state[idx] == 1 && *((volatile int*)&q.mask[idx]) == 0) {
            __threadfence();
            double val;
            if (q.cell[idx] == IFL_CELL_EMPTY) {                                   /* extrapolateAverage F8:544 */
                double value = 0.0; int count = 0;
                if (q.cell[idx - 1] == IFL_CELL_FLUID) { value += __ldcg(src + idx - 1); count++; }
                if (q.cell[idx + 1] == IFL_CELL_FLUID) { value += __ldcg(src + idx + 1); count++; }
                if (q.cell[idx - w] == IFL_CELL_FLUID) { value += __ldcg(src + idx - w); count++; }
                if (q.cell[idx + w] == IFL_CELL_FLUID) { value += __ldcg(src + idx + w); count++; }
                val = value / count;
            } else {
                val = flip_extrap_normal(q, src, idx);
            }
            src[idx] = val;
            state[idx] = 2;
            __threadfence();
            if (q.normalX[idx - 1] > 0.0 && q.cell[idx - 1] == IFL_CELL_SOLID) atomicAnd(&q.mask[idx - 1], ~1);
            if (q.normalX[idx + 1] < 0.0 && q.cell[idx + 1] == IFL_CELL_SOLID) atomicAnd(&q.mask[idx + 1], ~1);
            if (q.normalY[idx - w] > 0.0 && q.cell[idx - w] == IFL_CELL_SOLID) atomicAnd(&q.mask[idx - w], ~2);
            if (q.normalY[idx + w] < 0.0 && q.cell[idx + w] == IFL_CELL_SOLID) atomicAnd(&q.mask[idx + w], ~2);
            if (q.cell[idx - 1] == IFL_CELL_EMPTY) atomicCAS(&q.mask[idx - 1], 1, 0);
            if (q.cell[idx + 1] == IFL_CELL_EMPTY) atomicCAS(&q.mask[idx + 1], 1, 0);
            if (q.cell[idx - w] == IFL_CELL_EMPTY) atomicCAS(&q.mask[idx - w], 1, 0);
            if (q.cell[idx + w] == IFL_CELL_EMPTY) atomicCAS(&q.mask[idx + w], 1, 0);
            did = true;
        }
    }
    if (__syncthreads_or(did) && threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(progress, 1u);
}
/* literal replay of the reference's stack by one thread (only used when EMPTY cells touch) */
__global__ void k_flip_extrap_sequential(SolidQ q, double* src, int* stack) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int w = q.w, h = q.h;
    int top = 0;
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            int idx = x + y * w;
            if (q.cell[idx] != IFL_CELL_FLUID && q.mask[idx] == 0) stack[top++] = idx;
        }
    while (top > 0) {
        int idx = stack[--top];
        if (q.cell[idx] == IFL_CELL_EMPTY) {
            double value = 0.0; int count = 0;
            if (q.cell[idx - 1] == IFL_CELL_FLUID) { value += src[idx - 1]; count++; }
            if (q.cell[idx + 1] == IFL_CELL_FLUID) { value += src[idx + 1]; count++; }
            if (q.cell[idx - w] == IFL_CELL_FLUID) { value += src[idx - w]; count++; }
            if (q.cell[idx + w] == IFL_CELL_FLUID) { value += src[idx + w]; count++; }
            src[idx] = value / count;
            q.cell[idx] = IFL_CELL_FLUID;
        } else {
            double nx = q.normalX[idx], ny = q.normalY[idx];
            double srcX = src[idx + (int)f_signum(nx)];
            double srcY = src[idx + (int)f_signum(ny) * w];
            src[idx] = (jabsf(nx) * srcX + jabsf(ny) * srcY) / (jabsf(nx) + jabsf(ny));
        }
        const int nb[4] = {idx - 1, idx + 1, idx - w, idx + w};
        const bool cond[4] = {q.normalX[idx - 1] > 0.0, q.normalX[idx + 1] < 0.0, q.normalY[idx - w] > 0.0, q.normalY[idx + w] < 0.0};
        const int bit[4] = {1, 1, 2, 2};
        for (int k = 0; k < 4; k++)
            if (cond[k] && q.cell[nb[k]] == IFL_CELL_SOLID) {                      /* freeSolidNeighbour F8:567 */
                q.mask[nb[k]] &= ~bit[k];
                if (q.mask[nb[k]] == 0) stack[top++] = nb[k];
            }
        for (int k = 0; k < 4; k++)
            if (q.cell[nb[k]] == IFL_CELL_EMPTY && q.mask[nb[k]] == 1) {           /* freeEmptyNeighbour F8:584 */
                q.mask[nb[k]] = 0;
                stack[top++] = nb[k];
            }
    }
}
/* extrapolateEmptyBorders F8:598: edges (phase 0), corners (phase 1), relabel EMPTY -> FLUID (phase 2) */
__global__ void k_flip_empty_borders(SolidQ q, double* src, int phase) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = q.w, h = q.h;
    if (phase == 0) {
        if (i >= 1 && i < w - 1) {
            int idxT = i, idxB = i + (h - 1) * w;
            if (q.cell[idxT] == IFL_CELL_EMPTY) src[idxT] = src[idxT + w];
            if (q.cell[idxB] == IFL_CELL_EMPTY) src[idxB] = src[idxB - w];
        }
        if (i >= 1 && i < h - 1) {
            int idxL = i * w, idxR = i * w + w - 1;
            if (q.cell[idxL] == IFL_CELL_EMPTY) src[idxL] = src[idxL + 1];
            if (q.cell[idxR] == IFL_CELL_EMPTY) src[idxR] = src[idxR - 1];
        }
    } else if (phase == 1) {
        if (i == 0) {
            int tl = 0, tr = w - 1, bl = (h - 1) * w, br = h * w - 1;
            if (q.cell[tl] == IFL_CELL_EMPTY) src[tl] = 0.5 * (src[tl + 1] + src[tl + w]);
            if (q.cell[tr] == IFL_CELL_EMPTY) src[tr] = 0.5 * (src[tr - 1] + src[tr + w]);
            if (q.cell[bl] == IFL_CELL_EMPTY) src[bl] = 0.5 * (src[bl + 1] + src[bl - w]);
            if (q.cell[br] == IFL_CELL_EMPTY) src[br] = 0.5 * (src[br - 1] + src[br - w]);
        }
    } else {
        if (i < w * h && q.cell[i] == IFL_CELL_EMPTY) q.cell[i] = IFL_CELL_FLUID;
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * host side
 * ------------------------------------------------------------------------------------------------------- */
static inline unsigned nblk(long long n, int b = 256) { return (unsigned)((n + b - 1) / b); }

#define FCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t flip_alloc(FlipCtx& f, int w, int h, int maxpc) {
    f.w = w; f.h = h;
    f.max_particles = w * h * maxpc;                               /* F8:1549 */
    const size_t n = (size_t)f.max_particles;
    FCU(cudaMalloc(&f.posX, sizeof(double) * n)); FCU(cudaMalloc(&f.posY, sizeof(double) * n));
    for (int q = 0; q < 4; q++) { FCU(cudaMalloc(&f.prop[q], sizeof(double) * n)); FCU(cudaMemset(f.prop[q], 0, sizeof(double) * n)); }
    FCU(cudaMemset(f.posX, 0, sizeof(double) * n)); FCU(cudaMemset(f.posY, 0, sizeof(double) * n));
    const size_t ncell = (size_t)(w + 1) * (h + 1) + 1;
    FCU(cudaMalloc(&f.counts, sizeof(int) * ncell));
    FCU(cudaMalloc(&f.cell_start, sizeof(int) * (ncell + 1)));
    FCU(cudaMalloc(&f.cell_tmp, sizeof(int) * (ncell + 1)));
    FCU(cudaMalloc(&f.keys_a, sizeof(int) * n)); FCU(cudaMalloc(&f.keys_b, sizeof(int) * n));
    FCU(cudaMalloc(&f.vals_a, sizeof(int) * n)); FCU(cudaMalloc(&f.vals_b, sizeof(int) * n));
    FCU(cudaMalloc(&f.d_ints, sizeof(int) * 8));
    FCU(cudaMallocHost(&f.h_ints, sizeof(int) * 8));
    /* cub scratch: size for the largest problem (radix sort of max_particles pairs) */
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, f.keys_a, f.keys_b, f.vals_a, f.vals_b, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, f.keys_a, f.keys_b, (int)n);
    cub::DeviceSelect::Flagged(nullptr, b3, f.vals_a, f.keys_a, f.vals_b, f.d_ints, (int)n);
    f.cub_bytes = b1 > b2 ? b1 : b2;
    if (b3 > f.cub_bytes) f.cub_bytes = b3;
    FCU(cudaMalloc(&f.cub_tmp, f.cub_bytes));
    return cudaSuccess;
}
void flip_free(FlipCtx& f) {
    void* ptrs[] = {f.posX, f.posY, f.prop[0], f.prop[1], f.prop[2], f.prop[3], f.counts, f.cell_start, f.cell_tmp,
                    f.keys_a, f.keys_b, f.vals_a, f.vals_b, f.d_ints, f.cub_tmp};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (f.h_ints) cudaFreeHost(f.h_ints);
}

/* initParticles F8:1571 */
cudaError_t flip_init_particles(FlipCtx& f, cudaStream_t st, int avg, unsigned long long seed, const ifl_body* bodies, int nbodies,
                                double hx, unsigned long long* launches) {
    const long long total = (long long)f.w * f.h * avg;
    int* accept = f.keys_a; int* slot = f.keys_b;
    k_init_accept<<<nblk(total), 256, 0, st>>>(f.w, f.h, avg, seed, f.rng_counter, bodies, nbodies, hx, accept);
    size_t bytes = f.cub_bytes;
    FCU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, bytes, accept, slot, (int)total, st));
    k_init_write<<<nblk(total), 256, 0, st>>>(f.w, f.h, avg, seed, f.rng_counter, accept, slot, f.posX, f.posY);
    FCU(cudaMemcpyAsync(f.h_ints, slot + total - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(f.h_ints + 1, accept + total - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCU(cudaStreamSynchronize(st));
    f.count = f.h_ints[0] + f.h_ints[1];
    f.rng_counter += 2ull * (unsigned long long)total;
    *launches += 3;
    return cudaSuccess;
}

void flip_grid_to_particles(FlipCtx& f, cudaStream_t st, double alpha, QGeom qd, const double* d, QGeom qt, const double* t,
                            QGeom qu, const double* u, QGeom qv, const double* v, unsigned long long* launches) {
    if (f.count == 0) return;
    k_grid_to_particles<<<nblk(f.count), 256, 0, st>>>(f.count, alpha, f.posX, f.posY, grid_view(d, qd), grid_view(t, qt),
                                                      grid_view(u, qu), grid_view(v, qv), f.prop[0], f.prop[1], f.prop[2], f.prop[3]);
    *launches += 1;
}

void flip_advect_particles(FlipCtx& f, cudaStream_t st, double timestep, double hx, QGeom qu, const double* u, QGeom qv,
                           const double* v, const ifl_body* bodies, int nbodies, unsigned long long* launches) {
    if (f.count == 0) return;
    k_advect_particles<<<nblk(f.count), 256, 0, st>>>(f.count, timestep, hx, f.w, f.h, grid_view(u, qu), grid_view(v, qv),
                                                     bodies, nbodies, f.posX, f.posY);
    *launches += 1;
}

/* bin + stable sort for a quantity geometry; leaves cell_start / vals_b (sorted particle indices) */
cudaError_t flip_bin(FlipCtx& f, cudaStream_t st, QGeom q, unsigned long long* launches) {
    const int ncell = q.w * q.h;
    FCU(cudaMemsetAsync(f.cell_tmp, 0, sizeof(int) * (ncell + 1), st));
    if (f.count > 0) k_p2g_keys<<<nblk(f.count), 256, 0, st>>>(f.count, f.posX, f.posY, q.w, q.h, q.ox, q.oy, f.keys_a, f.vals_a, f.cell_tmp);
    size_t bytes = f.cub_bytes;
    FCU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, bytes, f.cell_tmp, f.cell_start, ncell + 1, st));
    if (f.count > 0) {
        int bits = 1;
        while ((1 << bits) < ncell) bits++;
        bytes = f.cub_bytes;
        FCU(cub::DeviceRadixSort::SortPairs(f.cub_tmp, bytes, f.keys_a, f.keys_b, f.vals_a, f.vals_b, f.count, 0, bits, st));
    }
    *launches += 3;
    return cudaSuccess;
}
void flip_gather(FlipCtx& f, cudaStream_t st, QGeom q, const double* prop, double* src, unsigned char* cell, unsigned long long* launches) {
    dim3 block(32, 8), grid((q.w + 31) / 32, (q.h + 7) / 8);
    k_p2g_gather<<<grid, block, 0, st>>>(q.w, q.h, q.ox, q.oy, f.cell_start, f.vals_b, f.posX, f.posY, prop, src, cell);
    *launches += 1;
}

void flip_diff(cudaStream_t st, double* src, const double* old, size_t n, double alpha, int sign, unsigned long long* launches) {
    k_diff<<<nblk((long long)n), 256, 0, st>>>(src, old, n, 1.0 - alpha, sign);
    *launches += 1;
}

/* countParticles + pruneParticles + seedParticles (F8:1595-1671) */
cudaError_t flip_count_prune_seed(FlipCtx& f, cudaStream_t st, int maxpc, int minpc, unsigned long long seed,
                                  const ifl_body* bodies, int nbodies, double hx,
                                  QGeom qd, const double* d, QGeom qt, const double* t, QGeom qu, const double* u,
                                  QGeom qv, const double* v, unsigned long long* launches) {
    const int ncell = f.w * f.h;
    FCU(cudaMemsetAsync(f.counts, 0, sizeof(int) * ncell, st));
    if (f.count > 0) k_count_particles<<<nblk(f.count), 256, 0, st>>>(f.count, f.posX, f.posY, f.w, f.h, f.counts);
    *launches += 1;
    /* prune */
    if (f.count > 0) {
        int* flags = f.keys_a; int* iota = f.vals_a; int* list = f.vals_b;
        k_prune_flags<<<nblk(f.count), 256, 0, st>>>(f.count, f.posX, f.posY, f.w, f.counts, maxpc, flags);
        k_iota<<<nblk(f.count), 256, 0, st>>>(f.count, iota);
        size_t bytes = f.cub_bytes;
        FCU(cub::DeviceSelect::Flagged(f.cub_tmp, bytes, iota, flags, list, f.d_ints, f.count, st));
        FCU(cudaMemcpyAsync(f.d_ints + 1, &f.count, sizeof(int), cudaMemcpyHostToDevice, st));
        k_prune_sequential<<<1, 32, 0, st>>>(list, f.d_ints, f.d_ints + 1, f.w, maxpc, f.counts, f.posX, f.posY,
                                             f.prop[0], f.prop[1], f.prop[2], f.prop[3]);
        FCU(cudaMemcpyAsync(f.h_ints, f.d_ints + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        f.count = f.h_ints[0];
        *launches += 4;
    }
    /* seed */
    int* need = f.keys_a; int* off = f.keys_b;
    k_seed_need<<<nblk(ncell), 256, 0, st>>>(ncell, f.counts, minpc, need);
    size_t bytes = f.cub_bytes;
    FCU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, bytes, need, off, ncell, st));
    FCU(cudaMemcpyAsync(f.h_ints, off + ncell - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(f.h_ints + 1, need + ncell - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCU(cudaStreamSynchronize(st));
    const int attempts = f.h_ints[0] + f.h_ints[1];
    *launches += 2;
    if (attempts > 0) {
        int* accept = f.vals_a; int* slot = f.vals_b;
        FCU(cudaMemsetAsync(accept, 0, sizeof(int) * attempts, st));
        k_seed_accept<<<nblk(ncell), 256, 0, st>>>(ncell, need, off, f.posX, f.posY, bodies, nbodies, hx, accept);
        bytes = f.cub_bytes;
        FCU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, bytes, accept, slot, attempts, st));
        FCU(cudaMemsetAsync(f.d_ints + 2, 0, sizeof(int) * 2, st));
        k_seed_write<<<nblk(ncell), 256, 0, st>>>(ncell, f.w, need, off, accept, slot, f.count, f.max_particles, seed, f.rng_counter,
                                                 f.posX, f.posY, grid_view(d, qd), grid_view(t, qt), grid_view(u, qu), grid_view(v, qv),
                                                 f.prop[0], f.prop[1], f.prop[2], f.prop[3], f.d_ints + 2, f.d_ints + 3);
        FCU(cudaMemcpyAsync(f.h_ints, f.d_ints + 2, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        f.rng_counter += 2ull * (unsigned long long)f.h_ints[0];
        f.count += f.h_ints[1];
        *launches += 3;
    }
    return cudaSuccess;
}

/* FluidQuantity.extrapolate, F8 flavour */
cudaError_t flip_extrapolate(FlipCtx& f, cudaStream_t st, const SolidQ& q, double* src, unsigned char* state,
                             unsigned int* d_progress, unsigned int* h_progress, int force_sequential, unsigned long long* launches) {
    dim3 block(32, 8), grid((q.w + 31) / 32, (q.h + 7) / 8);
    FCU(cudaMemsetAsync(f.d_ints + 4, 0, sizeof(int), st));
    k_flip_fill_mask<<<grid, block, 0, st>>>(q, state, f.d_ints + 4);
    FCU(cudaMemcpyAsync(f.h_ints + 4, f.d_ints + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCU(cudaStreamSynchronize(st));
    *launches += 1;
    if (f.h_ints[4] || force_sequential) {
        k_flip_extrap_sequential<<<1, 32, 0, st>>>(q, src, f.vals_a);
        *launches += 1;
        f.sequential_extrapolations++;
    } else {
        const int batch = 4;
        const int max_batches = (q.w + q.h) / batch + 2;
        for (int b = 0; b < max_batches; b++) {
            FCU(cudaMemsetAsync(d_progress, 0, sizeof(unsigned int), st));
            for (int i = 0; i < batch; i++) k_flip_extrap_sweep<<<grid, block, 0, st>>>(q, src, state, d_progress);
            *launches += batch;
            FCU(cudaMemcpyAsync(h_progress, d_progress, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            FCU(cudaStreamSynchronize(st));
            if (*h_progress == 0) break;
        }
    }
    const int nmax = q.w > q.h ? q.w : q.h;
    k_flip_empty_borders<<<nblk(nmax), 256, 0, st>>>(q, src, 0);
    k_flip_empty_borders<<<1, 32, 0, st>>>(q, src, 1);
    k_flip_empty_borders<<<nblk((long long)q.w * q.h), 256, 0, st>>>(q, src, 2);
    *launches += 3;
    return cudaSuccess;
}

} // namespace ifl
