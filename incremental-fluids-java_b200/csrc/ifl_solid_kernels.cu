/*
 * ifl_solid_kernels.cu -- F4-F7 grid kernels: signed-distance bodies, per-quantity solid fields
 * (phi, volume, normal, cell, body), velocity boundary conditions, masked / volume-weighted RHS,
 * face densities, the pressure and heat matrices, masked pressure application, buoyancy, the
 * dependency-ordered extrapolation into solids and the masked advection with back-projection.
 *
 * Citations: F4 = ExplicitFluid4solidBoundaries.java, F5 = ...5curvedBoundaries, F6 = ...6heat,
 * F7 = ...7variableDensity; SB/BOX/SPH = solidBodies/{A_SolidBody,SolidBox,SolidSphere}.java.
 * Scatter loops of the reference are restated as gathers that add the contributions in the
 * reference's (lexicographic) order.  Compiled with -fmad=false.
 */
#include "ifl_internal.h"

namespace ifl {

/* ---- rigid bodies (POD table, cos/sin of theta precomputed by the caller) ------------------- */
struct V2 { double x, y; };
__device__ __forceinline__ double jsignum(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
__device__ __forceinline__ V2 body_rotate(const ifl_body& b, double x, double y, int dir) {     /* SB:65 */
    const double c = b.cos_theta, sn = dir > 0 ? b.sin_theta : -b.sin_theta;
    V2 r; r.x = c * x + sn * y; r.y = -sn * x + c * y;
    return r;
}
__device__ __forceinline__ double body_nsgn(double v) {                                          /* BOX:54 */
    double r = jsignum(v);
    return fabs(r) == 0 ? 0.0 : r;
}
__device__ __forceinline__ double body_distance(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {                                                                /* BOX:80 */
        x -= b.pos_x; y -= b.pos_y;
        V2 r = body_rotate(b, x, y, -1);
        double dx = fabs(r.x) - b.scale_x * 0.5;
        double dy = fabs(r.y) - b.scale_y * 0.5;
        if (dx >= 0.0 || dy >= 0.0) {
            double mx = jmax(dx, 0.0), my = jmax(dy, 0.0);
            return sqrt(mx * mx + my * my);
        }
        return jmax(dx, dy);
    }
    double ex = x - b.pos_x, ey = y - b.pos_y;                                                   /* SPH:62 */
    return sqrt(ex * ex + ey * ey) - b.scale_x * 0.5;
}
__device__ __forceinline__ V2 body_closest_surface_point(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {                                                                /* BOX:95 */
        x -= b.pos_x; y -= b.pos_y;
        V2 r = body_rotate(b, x, y, -1);
        x = r.x; y = r.y;
        double dx = fabs(x) - b.scale_x * 0.5;
        double dy = fabs(y) - b.scale_y * 0.5;
        if (dx > dy) x = body_nsgn(x) * 0.5 * b.scale_x;
        else y = body_nsgn(y) * 0.5 * b.scale_y;
        r = body_rotate(b, x, y, +1);
        r.x = r.x + b.pos_x; r.y = r.y + b.pos_y;
        return r;
    }
    x -= b.pos_x; y -= b.pos_y;                                                                  /* SPH:67, SB:121,138 */
    V2 r = body_rotate(b, x, y, -1);
    x = r.x / b.scale_x; y = r.y / b.scale_y;
    double rr = sqrt(x * x + y * y);
    if (rr < 1e-4) { x = 0.5; y = 0.0; }
    else { x /= 2.0 * rr; y /= 2.0 * rr; }
    x *= b.scale_x; y *= b.scale_y;
    r = body_rotate(b, x, y, +1);
    r.x = r.x + b.pos_x; r.y = r.y + b.pos_y;
    return r;
}
__device__ __forceinline__ V2 body_distance_normal(const ifl_body& b, double x, double y) {
    if (b.type == IFL_BODY_BOX) {                                                                /* BOX:118 */
        x -= b.pos_x; y -= b.pos_y;
        V2 r = body_rotate(b, x, y, -1);
        x = r.x; y = r.y;
        double nx, ny;
        if (fabs(x) - b.scale_x * 0.5 > fabs(y) - b.scale_y * 0.5) { nx = body_nsgn(x); ny = 0.0; }
        else { nx = 0.0; ny = body_nsgn(y); }
        return body_rotate(b, nx, ny, +1);
    }
    x -= b.pos_x; y -= b.pos_y;                                                                  /* SPH:85 */
    double rr = sqrt(x * x + y * y);
    V2 n;
    if (rr < 1e-4) { n.x = 1.0; n.y = 0.0; }
    else { n.x = x / rr; n.y = y / rr; }
    return n;
}
__device__ __forceinline__ double body_velocity_x(const ifl_body& b, double /*x*/, double y) {   /* SB:194 */
    return (b.pos_y - y) * b.vel_theta + b.vel_x;
}
__device__ __forceinline__ double body_velocity_y(const ifl_body& b, double x, double /*y*/) {   /* SB:205 */
    return (x - b.pos_x) * b.vel_theta + b.vel_y;
}

/* F5:135, F5:144, F5:165 */
__device__ __forceinline__ double tri_occ(double out1, double in, double out2) {
    return 0.5 * in * in / ((out1 - in) * (out2 - in));
}
__device__ __forceinline__ double trap_occ(double out1, double out2, double in1, double in2) {
    return 0.5 * (-in1 / (out1 - in1) - in2 / (out2 - in2));
}
__device__ double occupancy(double d11, double d12, double d21, double d22) {
    int b = (d11 < 0.0 ? 1 : 0) | (d12 < 0.0 ? 2 : 0) | (d22 < 0.0 ? 4 : 0) | (d21 < 0.0 ? 8 : 0);
    switch (b) {
        case 0x0: return 0.0;
        case 0x1: return tri_occ(d21, d11, d12);
        case 0x2: return tri_occ(d11, d12, d22);
        case 0x4: return tri_occ(d12, d22, d21);
        case 0x8: return tri_occ(d22, d21, d11);
        case 0xE: return 1.0 - tri_occ(-d21, -d11, -d12);
        case 0xD: return 1.0 - tri_occ(-d11, -d12, -d22);
        case 0xB: return 1.0 - tri_occ(-d12, -d22, -d21);
        case 0x7: return 1.0 - tri_occ(-d22, -d21, -d11);
        case 0x3: return trap_occ(d21, d22, d11, d12);
        case 0x6: return trap_occ(d11, d21, d12, d22);
        case 0x9: return trap_occ(d12, d22, d11, d21);
        case 0xC: return trap_occ(d11, d12, d21, d22);
        case 0x5: return tri_occ(d11, d12, d22) + tri_occ(d22, d21, d11);
        case 0xA: return tri_occ(d21, d11, d12) + tri_occ(d12, d22, d21);
        case 0xF: return 1.0;
    }
    return 0.0;
}

/* ---- FluidQuantity.fillSolidFields  F4:329 / F5:463 (=F6:478, F7:484) ---------------------- */
__global__ void k_fill_phi(SolidQ q, const ifl_body* __restrict__ bodies, int nbodies, double hx) {
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix > q.w || iy > q.h) return;
    double x = (ix + q.ox - 0.5) * hx;
    double y = (iy + q.oy - 0.5) * hx;
    double ph = body_distance(bodies[0], x, y);
    for (int i = 1; i < nbodies; i++) ph = jmin(ph, body_distance(bodies[i], x, y));
    q.phi[(size_t)ix + (size_t)iy * (q.w + 1)] = ph;
}
__global__ void k_fill_solid(SolidQ q, const ifl_body* __restrict__ bodies, int nbodies, double hx, int variant) {
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= q.w || iy >= q.h) return;
    size_t idx = (size_t)ix + (size_t)iy * q.w;
    double x = (ix + q.ox) * hx;
    double y = (iy + q.oy) * hx;
    int bi = 0;
    double d = body_distance(bodies[0], x, y);
    for (int i = 1; i < nbodies; i++) {
        double id = body_distance(bodies[i], x, y);
        if (id < d) { bi = i; d = id; }
    }
    q.body[idx] = bi;
    unsigned char cell;
    if (variant == IFL_F4_SOLID_BOUNDARIES) {
        cell = d < 0.0 ? IFL_CELL_SOLID : IFL_CELL_FLUID;
    } else {
        size_t idxp = (size_t)ix + (size_t)iy * (q.w + 1);
        double vol = 1.0 - occupancy(q.phi[idxp], q.phi[idxp + 1], q.phi[idxp + q.w + 1], q.phi[idxp + q.w + 2]);
        if (vol < 0.01) vol = 0.0;
        q.volume[idx] = vol;
        cell = vol == 0.0 ? IFL_CELL_SOLID : IFL_CELL_FLUID;
    }
    V2 n = body_distance_normal(bodies[bi], x, y);
    q.normalX[idx] = n.x;
    q.normalY[idx] = n.y;
    q.cell[idx] = cell;
}
void launch_fill_solid_fields(cudaStream_t st, const SolidQ& q, const ifl_body* bodies, int nbodies, double hx, int variant) {
    dim3 block(32, 8);
    if (variant >= IFL_F5_CURVED_BOUNDARIES) {
        dim3 g1((q.w + 1 + 31) / 32, (q.h + 1 + 7) / 8);
        k_fill_phi<<<g1, block, 0, st>>>(q, bodies, nbodies, hx);
    }
    dim3 g2((q.w + 31) / 32, (q.h + 7) / 8);
    k_fill_solid<<<g2, block, 0, st>>>(q, bodies, nbodies, hx, variant);
}

/* skewed fluid flags of the d-quantity (1 = CELL_FLUID), padding 0 */
__global__ void k_cell_flags(SkewGeom g, const unsigned char* __restrict__ cell, unsigned char* __restrict__ fl) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;              /* padding stays 0 (allocation is zeroed, never written) */
    fl[skew_index(g, x, y)] = cell[(size_t)x + (size_t)y * g.w] == IFL_CELL_FLUID ? 1 : 0;
}
void launch_cell_flags(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, unsigned char* fl) {
    k_cell_flags<<<dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(32, 8), 0, st>>>(g, cell, fl);
}

/* ---- FluidSolver.setBoundaryCondition F4:948 (=F5:1062, F6:1154, F7:1185) --------------------
 * Gather form.  A face is written by the solid cell before it and by the solid cell after it; the
 * reference loops lexicographically, so the later cell (the face's own cell) wins.  All four faces
 * use velocityX (port quirk, kept).  Then the domain edges are zeroed. */
__global__ void k_set_boundary(int w, int h, const unsigned char* __restrict__ cell, const int* __restrict__ body,
                               const ifl_body* __restrict__ bodies, double* __restrict__ u, double* __restrict__ v, double hx) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x > w || y > h) return;
    if (y < h) {   /* u face (x, y) */
        size_t iu = (size_t)x + (size_t)y * (w + 1);
        if (x == 0 || x == w) u[iu] = 0.0;
        else {
            size_t c = (size_t)x + (size_t)y * w;
            if (cell[c] == IFL_CELL_SOLID) u[iu] = body_velocity_x(bodies[body[c]], x * hx, (y + 0.5) * hx);
            else if (cell[c - 1] == IFL_CELL_SOLID) u[iu] = body_velocity_x(bodies[body[c - 1]], ((x - 1) + 1.0) * hx, (y + 0.5) * hx);
        }
    }
    if (x < w) {   /* v face (x, y) */
        size_t iv = (size_t)x + (size_t)y * w;
        if (y == 0 || y == h) v[iv] = 0.0;
        else {
            size_t c = (size_t)x + (size_t)y * w;
            if (cell[c] == IFL_CELL_SOLID) v[iv] = body_velocity_x(bodies[body[c]], (x + 0.5) * hx, y * hx);
            else if (cell[c - w] == IFL_CELL_SOLID) v[iv] = body_velocity_x(bodies[body[c - w]], (x + 0.5) * hx, ((y - 1) + 1.0) * hx);
        }
    }
}
void launch_set_boundary(cudaStream_t st, int w, int h, const unsigned char* cell, const int* body,
                         const ifl_body* bodies, double* u, double* v, double hx) {
    dim3 block(32, 8), grid((w + 1 + 31) / 32, (h + 1 + 7) / 8);
    k_set_boundary<<<grid, block, 0, st>>>(w, h, cell, body, bodies, u, v, hx);
}

/* ---- buildRhs F4:693 (masked) / F5:785 (=F6:804, F7:813) volume weighted --------------------- */
__global__ void k_build_rhs_solid(SkewGeom g, SolidView d, SolidView uq, SolidView vq, const double* __restrict__ u,
                                  const double* __restrict__ v, const ifl_body* __restrict__ bodies, int nbodies,
                                  double* __restrict__ r, double hx, int variant) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int w = g.w, h = g.h;
    if (x >= w || y >= h) return;
    long long i = skew_index(g, x, y);
    size_t idx = (size_t)x + (size_t)y * w;
    const double scale = 1.0 / hx;
    if (d.cell[idx] != IFL_CELL_FLUID) { r[i] = 0.0; return; }
    size_t iu0 = (size_t)x + (size_t)y * (w + 1), iu1 = iu0 + 1;
    size_t iv0 = idx, iv1 = idx + w;
    if (variant == IFL_F4_SOLID_BOUNDARIES) {
        r[i] = -scale * (u[iu1] - u[iu0] + v[iv1] - v[iv0]);
        return;
    }
    double rv = -scale * (uq.volume[iu1] * u[iu1] - uq.volume[iu0] * u[iu0] + vq.volume[iv1] * v[iv1] - vq.volume[iv0] * v[iv0]);
    double vol = d.volume[idx];
    if (nbodies > 0) {
        if (x > 0)     rv -= (uq.volume[iu0] - vol) * body_velocity_x(bodies[d.body[idx - 1]], x * hx, (y + 0.5) * hx);
        if (y > 0)     rv -= (vq.volume[iv0] - vol) * body_velocity_y(bodies[d.body[idx - w]], (x + 0.5) * hx, y * hx);
        if (x < w - 1) rv += (uq.volume[iu1] - vol) * body_velocity_x(bodies[d.body[idx + 1]], (x + 1.0) * hx, (y + 0.5) * hx);
        if (y < h - 1) rv += (vq.volume[iv1] - vol) * body_velocity_y(bodies[d.body[idx + w]], (x + 0.5) * hx, (y + 1.0) * hx);
    }
    r[i] = rv;
}
void launch_build_rhs_solid(cudaStream_t st, const SkewGeom& g, const SolidView& d, const SolidView& uq, const SolidView& vq,
                            const double* u, const double* v, const ifl_body* bodies, int nbodies, double* r, double hx, int variant) {
    k_build_rhs_solid<<<dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(32, 8), 0, st>>>(g, d, uq, vq, u, v, bodies, nbodies, r, hx, variant);
}

/* ---- computeDensities F7:854: faces receive 0.5*density from the cell before, then from their own cell */
__device__ __forceinline__ double cell_density(const double* __restrict__ dsrc, const double* __restrict__ tsrc, size_t idx,
                                               double rho_air, double t_amb, double alpha) {
    double dens = rho_air * t_amb / tsrc[idx] * (1.0 + alpha * dsrc[idx]);
    return jmax(dens, 0.05 * rho_air);
}
__global__ void k_compute_densities(int w, int h, const double* __restrict__ dsrc, const double* __restrict__ tsrc,
                                    double* __restrict__ ud, double* __restrict__ vd, double rho_air, double t_amb, double alpha) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x > w || y > h) return;
    if (y < h) {
        double acc = 0.0;
        if (x > 0) acc += 0.5 * cell_density(dsrc, tsrc, (size_t)(x - 1) + (size_t)y * w, rho_air, t_amb, alpha);
        if (x < w) acc += 0.5 * cell_density(dsrc, tsrc, (size_t)x + (size_t)y * w, rho_air, t_amb, alpha);
        ud[(size_t)x + (size_t)y * (w + 1)] = acc;
    }
    if (x < w) {
        double acc = 0.0;
        if (y > 0) acc += 0.5 * cell_density(dsrc, tsrc, (size_t)x + (size_t)(y - 1) * w, rho_air, t_amb, alpha);
        if (y < h) acc += 0.5 * cell_density(dsrc, tsrc, (size_t)x + (size_t)y * w, rho_air, t_amb, alpha);
        vd[(size_t)x + (size_t)y * w] = acc;
    }
}
void launch_compute_densities(cudaStream_t st, int w, int h, const double* dsrc, const double* tsrc, double* ud, double* vd,
                              double rho_air, double rho_soot, double t_amb) {
    double alpha = (rho_soot - rho_air) / rho_air;
    dim3 block(32, 8), grid((w + 1 + 31) / 32, (h + 1 + 7) / 8);
    k_compute_densities<<<grid, block, 0, st>>>(w, h, dsrc, tsrc, ud, vd, rho_air, t_amb, alpha);
}

/* ---- buildPressureMatrix F4:713 / F5:825 (=F6:844) / F7:877 and buildHeatDiffusionMatrix F6:880 ----
 * Gather form: aDiag[idx] collects its terms in the order up, left, right, down (lexicographic scatter).
 * mode 0: factor = scale; 1: scale*volume; 2: scale*volume/density (y factor indexes _vDensity with the
 * stride of _u, F7:898, kept); 3: heat (aDiag starts at 1.0 for every cell, factor = scale). */
struct MatrixArgs {
    const unsigned char* cell;
    const double* uvol; const double* vvol;
    const double* ud; const double* vd;
    double scale; int mode;
};
__device__ __forceinline__ double fx(const MatrixArgs& a, int w, int x, int y) {   /* factor between (x,y) and (x+1,y) */
    if (a.mode == 1) return a.scale * a.uvol[(size_t)(x + 1) + (size_t)y * (w + 1)];
    if (a.mode == 2) return a.scale * a.uvol[(size_t)(x + 1) + (size_t)y * (w + 1)] / a.ud[(size_t)(x + 1) + (size_t)y * (w + 1)];
    return a.scale;
}
__device__ __forceinline__ double fy(const MatrixArgs& a, int w, int x, int y) {   /* factor between (x,y) and (x,y+1) */
    if (a.mode == 1) return a.scale * a.vvol[(size_t)x + (size_t)(y + 1) * w];
    if (a.mode == 2) return a.scale * a.vvol[(size_t)x + (size_t)(y + 1) * w] / a.vd[(size_t)x + (size_t)(y + 1) * (w + 1)];
    return a.scale;
}
__global__ void k_build_matrix_solid(SkewGeom g, MatrixArgs a, double* __restrict__ adiag, double* __restrict__ aplusx,
                                     double* __restrict__ aplusy) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int w = g.w, h = g.h;
    if (x >= w || y >= h) return;
    long long i = skew_index(g, x, y);
    size_t idx = (size_t)x + (size_t)y * w;
    const bool fl = a.cell[idx] == IFL_CELL_FLUID;
    double dg = a.mode == 3 ? 1.0 : 0.0, px = 0.0, py = 0.0;
    if (fl) {
        if (y > 0 && a.cell[idx - w] == IFL_CELL_FLUID) dg += fy(a, w, x, y - 1);
        if (x > 0 && a.cell[idx - 1] == IFL_CELL_FLUID) dg += fx(a, w, x - 1, y);
        if (x < w - 1 && a.cell[idx + 1] == IFL_CELL_FLUID) { double f = fx(a, w, x, y); dg += f; px = -f; }
        if (y < h - 1 && a.cell[idx + w] == IFL_CELL_FLUID) { double f = fy(a, w, x, y); dg += f; py = -f; }
    }
    adiag[i] = dg; aplusx[i] = px; aplusy[i] = py;
}
void launch_build_matrix_solid(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, const double* uvol, const double* vvol,
                               const double* ud, const double* vd, double scale, int mode,
                               double* adiag, double* aplusx, double* aplusy) {
    MatrixArgs a{cell, uvol, vvol, ud, vd, scale, mode};
    k_build_matrix_solid<<<dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(32, 8), 0, st>>>(g, a, adiag, aplusx, aplusy);
}

/* ---- applyPressure F4:927 (=F5:1041, F6:1114) / F7:1148: only fluid cells contribute --------- */
__global__ void k_apply_pressure_solid(SkewGeom g, const unsigned char* __restrict__ cell, const double* __restrict__ p,
                                       double* __restrict__ u, double* __restrict__ v, const double* __restrict__ ud,
                                       const double* __restrict__ vd, double scale, int use_density) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int w = g.w, h = g.h;
    if (x > w || y > h) return;
    if (y < h) {
        size_t iu = (size_t)x + (size_t)y * (w + 1);
        double val = u[iu];
        if (x > 0 && cell[(size_t)(x - 1) + (size_t)y * w] == IFL_CELL_FLUID) {
            double sp = scale * p[skew_index(g, x - 1, y)];
            val = val + (use_density ? sp / ud[iu] : sp);
        }
        if (x < w && cell[(size_t)x + (size_t)y * w] == IFL_CELL_FLUID) {
            double sp = scale * p[skew_index(g, x, y)];
            val = val - (use_density ? sp / ud[iu] : sp);
        }
        u[iu] = val;
    }
    if (x < w) {
        size_t iv = (size_t)x + (size_t)y * w;
        double val = v[iv];
        if (y > 0 && cell[(size_t)x + (size_t)(y - 1) * w] == IFL_CELL_FLUID) {
            double sp = scale * p[skew_index(g, x, y - 1)];
            val = val + (use_density ? sp / vd[iv] : sp);
        }
        if (y < h && cell[(size_t)x + (size_t)y * w] == IFL_CELL_FLUID) {
            double sp = scale * p[skew_index(g, x, y)];
            val = val - (use_density ? sp / vd[iv] : sp);
        }
        v[iv] = val;
    }
}
void launch_apply_pressure_solid(cudaStream_t st, const SkewGeom& g, const unsigned char* cell, const double* p, double* u, double* v,
                                 const double* ud, const double* vd, double scale, int use_density) {
    dim3 block(32, 8), grid((g.w + 1 + 31) / 32, (g.h + 1 + 7) / 8);
    k_apply_pressure_solid<<<grid, block, 0, st>>>(g, cell, p, u, v, ud, vd, scale, use_density);
}

/* ---- addBuoyancy F6:1135 (=F7:1169): v(x,y) gets b(x,y-1)*0.5 first, then b(x,y)*0.5 ---------- */
__global__ void k_add_buoyancy(int w, int h, const double* __restrict__ dsrc, const double* __restrict__ tsrc,
                               double* __restrict__ v, double tg, double alpha, double t_amb) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y > h) return;
    size_t iv = (size_t)x + (size_t)y * w;
    double val = v[iv];
    if (y > 0) {
        size_t c = iv - w;
        double b = tg * (alpha * dsrc[c] - (tsrc[c] - t_amb) / t_amb);
        val = val + b * 0.5;
    }
    if (y < h) {
        double b = tg * (alpha * dsrc[iv] - (tsrc[iv] - t_amb) / t_amb);
        val = val + b * 0.5;
    }
    v[iv] = val;
}
void launch_add_buoyancy(cudaStream_t st, int w, int h, const double* dsrc, const double* tsrc, double* v,
                         double timestep, double gravity, double rho_air, double rho_soot, double t_amb) {
    double alpha = (rho_soot - rho_air) / rho_air;
    dim3 block(32, 8), grid((w + 31) / 32, (h + 1 + 7) / 8);
    k_add_buoyancy<<<grid, block, 0, st>>>(w, h, dsrc, tsrc, v, timestep * gravity, alpha, t_amb);
}

/* ---- FluidQuantity.extrapolate F4:452 (fillSolidMask F4:396, extrapolateNormal F4:426, freeNeighbour
 * F4:441).  The reference pops a LIFO stack; every solid cell is filled exactly once, after the solid
 * cells it reads, so any dependency-respecting order gives the same values.  Here: repeated sweeps
 * in which every pending cell whose mask is clear fills itself and then clears its dependants' bits. */
__global__ void k_extrap_fill_mask(SolidQ q, unsigned char* __restrict__ state) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= q.w || y >= q.h) return;
    size_t idx = (size_t)x + (size_t)y * q.w;
    unsigned char stt = 0;
    if (x >= 1 && x < q.w - 1 && y >= 1 && y < q.h - 1 && q.cell[idx] != IFL_CELL_FLUID) {
        double nx = q.normalX[idx], ny = q.normalY[idx];
        int m = 0;
        if (nx != 0.0 && q.cell[idx + (int)jsignum(nx)] != IFL_CELL_FLUID) m |= 1;
        if (ny != 0.0 && q.cell[idx + (int)jsignum(ny) * q.w] != IFL_CELL_FLUID) m |= 2;
        q.mask[idx] = m;
        stt = 1;
    }
    state[idx] = stt;
}
/* Sweep number k of one extrapolate(): fills every pending solid cell whose upstream cells are ready.  Progress is
 * counted on the device in three rotating slots (sweep k adds to slot k%3, reads slot (k-1)%3 and clears slot (k+1)%3):
 * once a sweep has filled nothing, ctl[3] is raised and this and every later sweep of the call returns at once, so the
 * host can enqueue sweeps blindly (ifl_api.cu polls ctl[3] two chunks behind, like the PCG loop). */
__global__ void k_extrap_sweep(SolidQ q, double* src, unsigned char* state, unsigned int* ctl, int k) {
    if (ctl[3]) return;
    if (k > 0 && ctl[(k + 2) % 3] == 0) {                      /* the previous sweep made no progress: finished */
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) ctl[3] = 1;
        return;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) ctl[(k + 1) % 3] = 0;
    unsigned int* progress = ctl + (k % 3);
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    bool did = false;
    if (x >= 1 && x < q.w - 1 && y >= 1 && y < q.h - 1) {
        size_t idx = (size_t)x + (size_t)y * q.w;
        if (state[idx] == 1 && *((volatile int*)&q.mask[idx]) == 0) {
            __threadfence();
            double nx = q.normalX[idx], ny = q.normalY[idx];
            double srcX = __ldcg(src + idx + (int)jsignum(nx));
            double srcY = __ldcg(src + idx + (int)jsignum(ny) * q.w);
            double val = (jabsf(nx) * srcX + jabsf(ny) * srcY) / (jabsf(nx) + jabsf(ny));
            src[idx] = val;
            state[idx] = 2;
            __threadfence();
            if (q.normalX[idx - 1] > 0.0) atomicAnd(&q.mask[idx - 1], ~1);
            if (q.normalX[idx + 1] < 0.0) atomicAnd(&q.mask[idx + 1], ~1);
            if (q.normalY[idx - q.w] > 0.0) atomicAnd(&q.mask[idx - q.w], ~2);
            if (q.normalY[idx + q.w] < 0.0) atomicAnd(&q.mask[idx + q.w], ~2);
            did = true;
        }
    }
    if (__syncthreads_or(did) && threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(progress, 1u);
}
void launch_extrap_fill_mask(cudaStream_t st, const SolidQ& q, unsigned char* state) {
    dim3 block(32, 8), grid((q.w + 31) / 32, (q.h + 7) / 8);
    k_extrap_fill_mask<<<grid, block, 0, st>>>(q, state);
}
void launch_extrap_sweep(cudaStream_t st, const SolidQ& q, double* src, unsigned char* state, unsigned int* progress, int k) {
    dim3 block(32, 8), grid((q.w + 31) / 32, (q.h + 7) / 8);
    k_extrap_sweep<<<grid, block, 0, st>>>(q, src, state, progress, k);
}

/* ---- masked advection F4:268 (=F5:402, F6:417, F7:423) with backProject F4:249 --------------- */
struct AdvView {
    const double* __restrict__ p;
    int w, h;
    double ox, oy, wm, hm;
};
__host__ static AdvView adv_view(const double* p, QGeom q) {
    AdvView f; f.p = p; f.w = q.w; f.h = q.h; f.ox = q.ox; f.oy = q.oy;
    f.wm = (double)q.w - 1.001; f.hm = (double)q.h - 1.001;
    return f;
}
__device__ __forceinline__ double adv_lerp_grid(const AdvView& f, double x, double y) {      /* F4:207 */
    x = jmin(jmax(x - f.ox, 0.0), f.wm);
    y = jmin(jmax(y - f.oy, 0.0), f.hm);
    int ix = jint(x), iy = jint(y);
    x -= ix; y -= iy;
    const double* r0 = f.p + (size_t)iy * f.w + ix;
    const double* r1 = r0 + f.w;
    return lerp(lerp(__ldg(r0), __ldg(r0 + 1), x), lerp(__ldg(r1), __ldg(r1 + 1), x), y);
}
__device__ __forceinline__ double adv_cerp_grid(const AdvView& f, double x, double y) {      /* F4:225 */
    x = jmin(jmax(x - f.ox, 0.0), f.wm);
    y = jmin(jmax(y - f.oy, 0.0), f.hm);
    int ix = jint(x), iy = jint(y);
    x -= ix; y -= iy;
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
__global__ void __launch_bounds__(256)
k_advect_solid(AdvView q, AdvView u, AdvView v, const unsigned char* __restrict__ cell, const int* __restrict__ body,
               const ifl_body* __restrict__ bodies, double* __restrict__ dst, double timestep, double hx) {
    const HxDiv hd = make_hxdiv(hx);
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= q.w || iy >= q.h) return;
    size_t idx = (size_t)iy * q.w + ix;
    if (cell[idx] != IFL_CELL_FLUID) return;      /* stale _dst is kept (F4:273) */
    double x = ix + q.ox;
    double y = iy + q.oy;
    double firstU = div_hx(adv_lerp_grid(u, x, y), hd);
    double firstV = div_hx(adv_lerp_grid(v, x, y), hd);
    double midX = x - 0.5 * timestep * firstU;
    double midY = y - 0.5 * timestep * firstV;
    double midU = div_hx(adv_lerp_grid(u, midX, midY), hd);
    double midV = div_hx(adv_lerp_grid(v, midX, midY), hd);
    double lastX = x - 0.75 * timestep * midU;
    double lastY = y - 0.75 * timestep * midV;
    double lastU = adv_lerp_grid(u, lastX, lastY);
    double lastV = adv_lerp_grid(v, lastX, lastY);
    x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
    y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
    /* backProject F4:249 */
    int rx = min(max(jint(x - q.ox), 0), q.w - 1);
    int ry = min(max(jint(y - q.oy), 0), q.h - 1);
    size_t c = (size_t)rx + (size_t)ry * q.w;
    if (cell[c] != IFL_CELL_FLUID) {
        x = (x - q.ox) * hx;
        y = (y - q.oy) * hx;
        V2 r = body_closest_surface_point(bodies[body[c]], x, y);
        x = div_hx(r.x, hd) + q.ox;
        y = div_hx(r.y, hd) + q.oy;
    }
    dst[idx] = adv_cerp_grid(q, x, y);
}
void launch_advect_solid(cudaStream_t st, QGeom q, const double* src, double* dst, QGeom qu, const double* u, QGeom qv,
                         const double* v, const unsigned char* cell, const int* body, const ifl_body* bodies,
                         double timestep, double hx) {
    dim3 block(32, 8), grid((q.w + 31) / 32, (q.h + 7) / 8);
    k_advect_solid<<<grid, block, 0, st>>>(adv_view(src, q), adv_view(u, qu), adv_view(v, qv), cell, body, bodies, dst, timestep, hx);
}

__global__ void k_fill_value(double* p, size_t n, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill_value(cudaStream_t st, double* p, size_t n, double v) {
    k_fill_value<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
}
__global__ void k_fill_u8(unsigned char* p, size_t n, unsigned char v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill_u8(cudaStream_t st, unsigned char* p, size_t n, unsigned char v) {
    k_fill_u8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
}
/* widen integer fields for ifl_read_field */
__global__ void k_u8_to_double(const unsigned char* a, double* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)a[i];
}
__global__ void k_i32_to_double(const int* a, double* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)a[i];
}
void launch_u8_to_double(cudaStream_t st, const unsigned char* a, double* out, size_t n) {
    k_u8_to_double<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, out, n);
}
void launch_i32_to_double(cudaStream_t st, const int* a, double* out, size_t n) {
    k_i32_to_double<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, out, n);
}

} // namespace ifl
