/*
 * ifl_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A single-threaded restatement of the arithmetic of g-amador/incremental-fluids-java
 * for the timestep hot path, operation for operation and in
 * the reference's evaluation order, with Java semantics made explicit (jmin/jmax,
 * (float) round trips, (int) truncation, no FMA: build with -ffp-contract=off).
 * Every function cites the reference file:line it follows; paths are relative to
 * /root/reference/src/physics/, with F1..F8 = the ExplicitFluid<N>*.java files.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libincfluid_b200.so) never does.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- the reference ships no
 * tests, golden vectors or sample frames and no JVM exists in this environment, so the
 * Java program cannot be executed.  The pins are (tests/test_oracle_*.py):
 *   (i)  line-by-line traceability (citations below),
 *   (ii) analytic invariants of the reference's formulas,
 *   (iii) solver iteration counts measured independently during the survey
 *         which this restatement must reproduce,
 *   (iv) a second restatement made independently from the Java text -- literal Python
 *        transcriptions of ExplicitFluid1 ... ExplicitFluid8 and physics/solidBodies
 *        (tests/java_f*_transcript.py, tests/java_solid_transcript.py) -- that this library
 *        must reproduce bit for bit on whole timesteps (tests/test_oracle_vs_transcript.py),
 *   (v)  committed golden dumps (tests/golden/) that it must keep reproducing.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../include/incfluid_b200.h"

namespace {

/* ---- Java semantics ---------------------------------------------------------------- */
/* Math.max/min(double,double): NaN propagating, max(-0.0,+0.0)=+0.0, min(-0.0,+0.0)=-0.0 */
inline double jmax(double a, double b) {
    if (a != a) return a;
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? b : a;
    return (a >= b) ? a : b;
}
inline double jmin(double a, double b) {
    if (a != a) return a;
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
    return (a <= b) ? a : b;
}
inline int jmaxi(int a, int b) { return a > b ? a : b; }
inline int jmini(int a, int b) { return a < b ? a : b; }
/* (int) double: truncation toward zero, saturating, NaN -> 0 (JLS 5.1.3) */
inline int jint(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int)x;
}
/* (float) x widened back to double */
inline double jfloat(double x) { return (double)(float)x; }
/* Math.abs((float) x) as a double */
inline double jabsf(double x) { return (double)std::fabs((float)x); }
/* Math.signum */
inline double jsignum(double x) {
    if (x != x) return x;
    if (x > 0.0) return 1.0;
    if (x < 0.0) return -1.0;
    return x; /* +-0.0 */
}

/* math/Interpolation.java:77-86 -- the "precise" lerp */
inline double lerp(double a, double b, double x) { return (1 - x) * a + x * b; }

/* math/Interpolation.java:135-152 -- Catmull-Rom, clamped to the sample range */
inline double cerp(double a, double b, double c, double d, double x) {
    double xsq = x * x;
    double xcu = xsq * x;
    double minV = jmin(a, jmin(b, jmin(c, d)));
    double maxV = jmax(a, jmax(b, jmax(c, d)));
    double t = a * (0.0 - 0.5 * x + 1.0 * xsq - 0.5 * xcu)
             + b * (1.0 + 0.0 * x - 2.5 * xsq + 1.5 * xcu)
             + c * (0.0 + 0.5 * x + 2.0 * xsq - 1.5 * xcu)
             + d * (0.0 + 0.0 * x - 0.5 * xsq + 0.5 * xcu);
    return jmin(jmax(t, minV), maxV);
}

/* F2:97 F3:97 */
inline double length(double x, double y) { return std::sqrt(x * x + y * y); }
/* F2:107 F3:109 -- argument rounded through (float) */
inline double cubicPulse(double x) {
    x = jmin(jabsf(x), 1.0);
    return 1.0 - x * x * (3.0 - 2.0 * x);
}

struct Vec2 { double x, y; };

/* Injected replacement of the unseeded Math.random() (F8:1576,1647): the k-th draw of the stream is a pure
 * function of (seed, k) so that a parallel implementation can consume it in the reference's order.
 * Returns (float) in [0, 1) exactly as `(float) random()` would deliver it to the float addition. */
inline float rng_uniform(uint64_t seed, uint64_t k) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (k + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

enum { CELL_FLUID = 0, CELL_SOLID = 1, CELL_EMPTY = 2, CELL_NULL = 3 /* F4: _cell entries stay null until fillSolidFields runs (F4:183) */ };

/* ---- solidBodies/A_SolidBody.java, SolidBox.java, SolidSphere.java --------------------------------
 * cos(theta)/sin(theta) come precomputed in ifl_body (see include/incfluid_b200.h): the reference
 * calls Math.cos/sin inside rotate() (A_SolidBody.java:65-71), which is not bit-specified. */
struct Body {
    ifl_body b;
    /* A_SolidBody.rotate(x, y, phi) with phi = +theta (dir=+1) or -theta (dir=-1): cos is even, sin is odd */
    Vec2 rotate(double x, double y, int dir) const {
        double c = b.cos_theta, sn = dir > 0 ? b.sin_theta : -b.sin_theta;
        return {c * x + sn * y, -sn * x + c * y};
    }
    static double nsgn(double v) {                     /* SolidBox.java:54-60: signum, 0 for 0 */
        double r = jsignum(v);
        if (std::fabs(r) == 0) return 0;
        return r;
    }
    double distance(double x, double y) const {
        if (b.type == IFL_BODY_BOX) {                  /* SolidBox.java:80 */
            x -= b.pos_x; y -= b.pos_y;
            Vec2 r = rotate(x, y, -1);
            double dx = std::fabs(r.x) - b.scale_x * 0.5;
            double dy = std::fabs(r.y) - b.scale_y * 0.5;
            if (dx >= 0.0 || dy >= 0.0) return length(jmax(dx, 0.0), jmax(dy, 0.0));
            return jmax(dx, dy);
        }
        return length(x - b.pos_x, y - b.pos_y) - b.scale_x * 0.5;     /* SolidSphere.java:62 */
    }
    Vec2 closestSurfacePoint(double x, double y) const {
        if (b.type == IFL_BODY_BOX) {                  /* SolidBox.java:95 */
            x -= b.pos_x; y -= b.pos_y;
            Vec2 r = rotate(x, y, -1);
            x = r.x; y = r.y;
            double dx = std::fabs(x) - b.scale_x * 0.5;
            double dy = std::fabs(y) - b.scale_y * 0.5;
            if (dx > dy) x = nsgn(x) * 0.5 * b.scale_x;
            else y = nsgn(y) * 0.5 * b.scale_y;
            r = rotate(x, y, +1);
            return {r.x + b.pos_x, r.y + b.pos_y};
        }
        /* SolidSphere.java:67 via globalToLocal / localToGlobal (A_SolidBody.java:121,138) */
        x -= b.pos_x; y -= b.pos_y;
        Vec2 r = rotate(x, y, -1);
        x = r.x / b.scale_x; y = r.y / b.scale_y;
        double rr = length(x, y);
        if (rr < 1e-4) { x = 0.5; y = 0.0; }
        else { x /= 2.0 * rr; y /= 2.0 * rr; }
        x *= b.scale_x; y *= b.scale_y;
        r = rotate(x, y, +1);
        return {r.x + b.pos_x, r.y + b.pos_y};
    }
    Vec2 distanceNormal(double x, double y) const {
        if (b.type == IFL_BODY_BOX) {                  /* SolidBox.java:118 */
            x -= b.pos_x; y -= b.pos_y;
            Vec2 r = rotate(x, y, -1);
            x = r.x; y = r.y;
            double nx, ny;
            if (std::fabs(x) - b.scale_x * 0.5 > std::fabs(y) - b.scale_y * 0.5) { nx = nsgn(x); ny = 0.0; }
            else { nx = 0.0; ny = nsgn(y); }
            return rotate(nx, ny, +1);
        }
        x -= b.pos_x; y -= b.pos_y;                    /* SolidSphere.java:85 */
        double rr = length(x, y);
        if (rr < 1e-4) return {1.0, 0.0};
        return {x / rr, y / rr};
    }
    double velocityX(double /*x*/, double y) const { return (b.pos_y - y) * b.vel_theta + b.vel_x; }   /* A_SolidBody.java:194 */
    double velocityY(double x, double /*y*/) const { return (x - b.pos_x) * b.vel_theta + b.vel_y; }   /* A_SolidBody.java:205 */
};

/* F5:135 / F5:144 / F5:165 (identical in F6, F7, F8) */
inline double triangleOccupancy(double out1, double in, double out2) { return 0.5 * in * in / ((out1 - in) * (out2 - in)); }
inline double trapezoidOccupancy(double out1, double out2, double in1, double in2) {
    return 0.5 * (-in1 / (out1 - in1) - in2 / (out2 - in2));
}
inline double occupancy(double d11, double d12, double d21, double d22) {
    double ds[4] = {d11, d12, d22, d21};
    int b = 0;
    for (int i = 3; i >= 0; i--) b = (b << 1) | (ds[i] < 0.0 ? 1 : 0);
    switch (b) {
        case 0x0: return 0.0;
        case 0x1: return triangleOccupancy(d21, d11, d12);
        case 0x2: return triangleOccupancy(d11, d12, d22);
        case 0x4: return triangleOccupancy(d12, d22, d21);
        case 0x8: return triangleOccupancy(d22, d21, d11);
        case 0xE: return 1.0 - triangleOccupancy(-d21, -d11, -d12);
        case 0xD: return 1.0 - triangleOccupancy(-d11, -d12, -d22);
        case 0xB: return 1.0 - triangleOccupancy(-d12, -d22, -d21);
        case 0x7: return 1.0 - triangleOccupancy(-d22, -d21, -d11);
        case 0x3: return trapezoidOccupancy(d21, d22, d11, d12);
        case 0x6: return trapezoidOccupancy(d11, d21, d12, d22);
        case 0x9: return trapezoidOccupancy(d12, d22, d11, d21);
        case 0xC: return trapezoidOccupancy(d11, d12, d21, d22);
        case 0x5: return triangleOccupancy(d11, d12, d22) + triangleOccupancy(d22, d21, d11);
        case 0xA: return triangleOccupancy(d21, d11, d12) + triangleOccupancy(d12, d22, d21);
        case 0xF: return 1.0;
    }
    return 0.0;
}


/* ---- FluidQuantity (F1:105, F3:114, F4:126, F5:234, F7:225) ----------------------------- */
struct Quantity {
    std::vector<double> src, dst;
    std::vector<double> phi, volume, normalX, normalY;     /* F4:138-153, F5:241-266 */
    std::vector<int> cell, body, mask;
    int w = 0, h = 0;
    double ox = 0, oy = 0, hx = 0;

    void init(int w_, int h_, double ox_, double oy_, double hx_, int variant = 0) {
        w = w_; h = h_; ox = ox_; oy = oy_; hx = hx_;
        src.assign((size_t)w * h, 0.0);
        dst.assign((size_t)w * h, 0.0);
        if (variant >= IFL_F4_SOLID_BOUNDARIES) {
            size_t n = (size_t)w * h;
            phi.assign((size_t)(w + 1) * (h + 1), 0.0);
            /* F5:313-316 initialises _cell = FLUID, _volume = 1; F4:183 leaves _cell null */
            volume.assign(n, 1.0);
            normalX.assign(n, 0.0); normalY.assign(n, 0.0);
            cell.assign(n, variant == IFL_F4_SOLID_BOUNDARIES ? CELL_NULL : CELL_FLUID);
            body.assign(n, 0); mask.assign(n, 0);
        }
    }
    double vol(int x, int y) const { return volume[(size_t)x + (size_t)y * w]; }
    void swap() { src.swap(dst); }                      /* F1:143 */
    double getSrc(int x, int y) const { return src[(size_t)x + (size_t)y * w]; }
    void setSrc(int x, int y, double v) { src[(size_t)x + (size_t)y * w] = v; }

    /* F1:161 (=F2:165, F3:167, F4:207, F5:341, F7:363, F8:365) */
    double lerpGrid(double x, double y) const {
        x = jmin(jmax(x - ox, 0.0), w - 1.001);
        y = jmin(jmax(y - oy, 0.0), h - 1.001);
        int ix = jint(x);
        int iy = jint(y);
        x -= ix;
        y -= iy;
        double x00 = getSrc(ix + 0, iy + 0), x10 = getSrc(ix + 1, iy + 0);
        double x01 = getSrc(ix + 0, iy + 1), x11 = getSrc(ix + 1, iy + 1);
        return lerp(lerp(x00, x10, x), lerp(x01, x11, x), y);
    }

    /* F2:183 (=F3:185, F4:225, F5:359, F7:381) */
    double cerpGrid(double x, double y) const {
        x = jmin(jmax(x - ox, 0.0), w - 1.001);
        y = jmin(jmax(y - oy, 0.0), h - 1.001);
        int ix = jint(x);
        int iy = jint(y);
        x -= ix;
        y -= iy;
        int x0 = jmaxi(ix - 1, 0), x1 = ix, x2 = ix + 1, x3 = jmini(ix + 2, w - 1);
        int y0 = jmaxi(iy - 1, 0), y1 = iy, y2 = iy + 1, y3 = jmini(iy + 2, h - 1);
        double q0 = cerp(getSrc(x0, y0), getSrc(x1, y0), getSrc(x2, y0), getSrc(x3, y0), x);
        double q1 = cerp(getSrc(x0, y1), getSrc(x1, y1), getSrc(x2, y1), getSrc(x3, y1), x);
        double q2 = cerp(getSrc(x0, y2), getSrc(x1, y2), getSrc(x2, y2), getSrc(x3, y2), x);
        double q3 = cerp(getSrc(x0, y3), getSrc(x1, y3), getSrc(x2, y3), getSrc(x3, y3), x);
        return cerp(q0, q1, q2, q3, y);
    }

    /* F1:220 */
    Vec2 Euler(double x, double y, double timestep, const Quantity& u, const Quantity& v) const {
        double uVel = u.lerpGrid(x, y) / hx;
        double vVel = v.lerpGrid(x, y) / hx;
        x -= uVel * timestep;
        y -= vVel * timestep;
        return {x, y};
    }

    /* F2:253 (=F3:255, F4:505, F5:669, F6:313, F7:315).  NOTE lastU/lastV are NOT divided by hx. */
    Vec2 RungeKutta3(double x, double y, double timestep, const Quantity& u, const Quantity& v) const {
        double firstU = u.lerpGrid(x, y) / hx;
        double firstV = v.lerpGrid(x, y) / hx;
        double midX = x - 0.5 * timestep * firstU;
        double midY = y - 0.5 * timestep * firstV;
        double midU = u.lerpGrid(midX, midY) / hx;
        double midV = v.lerpGrid(midX, midY) / hx;
        double lastX = x - 0.75 * timestep * midU;
        double lastY = y - 0.75 * timestep * midV;
        double lastU = u.lerpGrid(lastX, lastY);
        double lastV = v.lerpGrid(lastX, lastY);
        x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
        y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
        return {x, y};
    }

    /* F1:178 (Euler + lerpGrid) / F2:205, F3:207 (RK3 + cerpGrid) */
    void advect(int variant, double timestep, const Quantity& u, const Quantity& v) {
        size_t idx = 0;
        for (int iy = 0; iy < h; iy++) {
            for (int ix = 0; ix < w; ix++, idx++) {
                double x = ix + ox;
                double y = iy + oy;
                if (variant == IFL_F1_MATRIXLESS) {
                    Vec2 r = Euler(x, y, timestep, u, v);
                    dst[idx] = lerpGrid(r.x, r.y);
                } else {
                    Vec2 r = RungeKutta3(x, y, timestep, u, v);
                    dst[idx] = cerpGrid(r.x, r.y);
                }
            }
        }
    }

    /* F4:249 (=F5:383, F7:405) */
    Vec2 backProject(double x, double y, const std::vector<Body>& bodies) const {
        int rx = jmini(jmaxi(jint(x - ox), 0), w - 1);
        int ry = jmini(jmaxi(jint(y - oy), 0), h - 1);
        size_t c = (size_t)rx + (size_t)ry * w;
        if (cell[c] != CELL_FLUID) {
            x = (x - ox) * hx;
            y = (y - oy) * hx;
            Vec2 r = bodies[body[c]].closestSurfacePoint(x, y);
            x = r.x / hx + ox;
            y = r.y / hx + oy;
        }
        return {x, y};
    }

    /* F4:268 (=F5:402, F6:417, F7:423): only CELL_FLUID samples are written, the rest keeps stale _dst */
    void advectMasked(double timestep, const Quantity& u, const Quantity& v, const std::vector<Body>& bodies) {
        size_t idx = 0;
        for (int iy = 0; iy < h; iy++) {
            for (int ix = 0; ix < w; ix++, idx++) {
                if (cell[idx] == CELL_FLUID) {
                    double x = ix + ox;
                    double y = iy + oy;
                    Vec2 r = RungeKutta3(x, y, timestep, u, v);
                    r = backProject(r.x, r.y, bodies);
                    dst[idx] = cerpGrid(r.x, r.y);
                }
            }
        }
    }

    /* F4:329 (voxelised) / F5:463 (=F6:478, F7:484): phi, closest body, volume, normal, cell type */
    void fillSolidFields(int variant, const std::vector<Body>& bodies) {
        if (bodies.empty()) return;
        if (variant >= IFL_F5_CURVED_BOUNDARIES) {
            size_t idx = 0;
            for (int iy = 0; iy <= h; iy++) {
                for (int ix = 0; ix <= w; ix++, idx++) {
                    double x = (ix + ox - 0.5) * hx;
                    double y = (iy + oy - 0.5) * hx;
                    phi[idx] = bodies[0].distance(x, y);
                    for (size_t i = 1; i < bodies.size(); i++) phi[idx] = jmin(phi[idx], bodies[i].distance(x, y));
                }
            }
        }
        size_t idx = 0;
        for (int iy = 0; iy < h; iy++) {
            for (int ix = 0; ix < w; ix++, idx++) {
                double x = (ix + ox) * hx;
                double y = (iy + oy) * hx;
                body[idx] = 0;
                double d = bodies[0].distance(x, y);
                for (size_t i = 1; i < bodies.size(); i++) {
                    double id = bodies[i].distance(x, y);
                    if (id < d) { body[idx] = (int)i; d = id; }
                }
                if (variant == IFL_F4_SOLID_BOUNDARIES) {
                    cell[idx] = d < 0.0 ? CELL_SOLID : CELL_FLUID;
                } else {
                    size_t idxp = (size_t)ix + (size_t)iy * (w + 1);
                    volume[idx] = 1.0 - occupancy(phi[idxp], phi[idxp + 1], phi[idxp + w + 1], phi[idxp + w + 2]);
                    if (volume[idx] < 0.01) volume[idx] = 0.0;
                }
                Vec2 n = bodies[body[idx]].distanceNormal(x, y);
                normalX[idx] = n.x;
                normalY[idx] = n.y;
                if (variant != IFL_F4_SOLID_BOUNDARIES) cell[idx] = volume[idx] == 0.0 ? CELL_SOLID : CELL_FLUID;
            }
        }
    }

    /* F4:396 */
    void fillSolidMask() {
        for (int y = 1; y < h - 1; y++) {
            for (int x = 1; x < w - 1; x++) {
                size_t idx = (size_t)x + (size_t)y * w;
                if (cell[idx] == CELL_FLUID) continue;
                double nx = normalX[idx], ny = normalY[idx];
                mask[idx] = 0;
                if (nx != 0.0 && cell[idx + (int)jsignum(nx)] != CELL_FLUID) mask[idx] |= 1;
                if (ny != 0.0 && cell[idx + (int)jsignum(ny) * w] != CELL_FLUID) mask[idx] |= 2;
            }
        }
    }
    /* F4:426: weights are abs((float) n) */
    double extrapolateNormal(size_t idx) const {
        double nx = normalX[idx], ny = normalY[idx];
        double srcX = src[idx + (int)jsignum(nx)];
        double srcY = src[idx + (int)jsignum(ny) * w];
        return (jabsf(nx) * srcX + jabsf(ny) * srcY) / (jabsf(nx) + jabsf(ny));
    }
    void freeNeighbour(size_t idx, std::vector<size_t>& border, int m) {   /* F4:441 */
        mask[idx] &= ~m;
        if (cell[idx] != CELL_FLUID && mask[idx] == 0) border.push_back(idx);
    }
    /* F4:452 (=F5:616, F6:632, F7:632): dependency ordered fill of solid cells, LIFO stack */
    void extrapolate() {
        fillSolidMask();
        std::vector<size_t> border;
        for (int y = 1; y < h - 1; y++)
            for (int x = 1; x < w - 1; x++) {
                size_t idx = (size_t)x + (size_t)y * w;
                if (cell[idx] != CELL_FLUID && mask[idx] == 0) border.push_back(idx);
            }
        while (!border.empty()) {
            size_t idx = border.back();
            border.pop_back();
            src[idx] = extrapolateNormal(idx);
            if (normalX[idx - 1] > 0.0) freeNeighbour(idx - 1, border, 1);
            if (normalX[idx + 1] < 0.0) freeNeighbour(idx + 1, border, 1);
            if (normalY[idx - w] > 0.0) freeNeighbour(idx - w, border, 2);
            if (normalY[idx + w] < 0.0) freeNeighbour(idx + w, border, 2);
        }
    }

    /* ================= F8 (FLIP) grid side ================= */
    /* F8:312 */
    void addSample(std::vector<double>& weight, double value, double x, double y, int ix, int iy) {
        if (ix < 0 || iy < 0 || ix >= w || iy >= h) return;
        double k = (1.0 - std::fabs(ix - x)) * (1.0 - std::fabs(iy - y));
        weight[(size_t)ix + (size_t)iy * w] += k;
        src[(size_t)ix + (size_t)iy * w] += k * value;
    }
    /* F8:722: particle-index order; FLUID cells without weight become EMPTY */
    void fromParticles(int count, const std::vector<double>& posX, const std::vector<double>& posY,
                       const std::vector<double>& property) {
        std::fill(src.begin(), src.end(), 0.0);
        std::vector<double> weight((size_t)w * h, 0.0);
        for (int i = 0; i < count; i++) {
            double x = posX[i] - ox;
            double y = posY[i] - oy;
            x = jmax(0.5, jmin(w - 1.5, x));
            y = jmax(0.5, jmin(h - 1.5, y));
            int ix = jint(x), iy = jint(y);
            addSample(weight, property[i], x, y, ix + 0, iy + 0);
            addSample(weight, property[i], x, y, ix + 1, iy + 0);
            addSample(weight, property[i], x, y, ix + 0, iy + 1);
            addSample(weight, property[i], x, y, ix + 1, iy + 1);
        }
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) {
            if (weight[i] != 0.0) src[i] /= weight[i];
            else if (cell[i] == CELL_FLUID) cell[i] = CELL_EMPTY;
        }
    }
    void copyOld() { dst = src; }                                                   /* F8:339 (_old lives in dst) */
    void diff(double alpha) { for (size_t i = 0; i < src.size(); i++) src[i] -= (1.0 - alpha) * dst[i]; }    /* F8:346 */
    void undiff(double alpha) { for (size_t i = 0; i < src.size(); i++) src[i] += (1.0 - alpha) * dst[i]; }  /* F8:355 */

    /* F8:483 */
    void fillSolidMaskFlip() {
        for (int x = 0; x < w; x++) mask[x] = mask[(size_t)x + (size_t)(h - 1) * w] = 0xFF;
        for (int y = 0; y < h; y++) mask[(size_t)y * w] = mask[(size_t)y * w + w - 1] = 0xFF;
        for (int y = 1; y < h - 1; y++) {
            for (int x = 1; x < w - 1; x++) {
                size_t idx = (size_t)x + (size_t)y * w;
                mask[idx] = 0;
                if (cell[idx] == CELL_SOLID) {
                    double nx = normalX[idx], ny = normalY[idx];
                    if (nx != 0.0 && cell[idx + (int)jsignum(nx)] != CELL_FLUID) mask[idx] |= 1;
                    if (ny != 0.0 && cell[idx + (int)jsignum(ny) * w] != CELL_FLUID) mask[idx] |= 2;
                } else if (cell[idx] == CELL_EMPTY) {
                    mask[idx] = (cell[idx - 1] != CELL_FLUID && cell[idx + 1] != CELL_FLUID &&
                                 cell[idx - w] != CELL_FLUID && cell[idx + w] != CELL_FLUID) ? 1 : 0;
                }
            }
        }
    }
    double extrapolateAverage(size_t idx) const {                                    /* F8:544 */
        double value = 0.0;
        int count = 0;
        if (cell[idx - 1] == CELL_FLUID) { value += src[idx - 1]; count++; }
        if (cell[idx + 1] == CELL_FLUID) { value += src[idx + 1]; count++; }
        if (cell[idx - w] == CELL_FLUID) { value += src[idx - w]; count++; }
        if (cell[idx + w] == CELL_FLUID) { value += src[idx + w]; count++; }
        return value / count;
    }
    void freeSolidNeighbour(size_t idx, std::vector<size_t>& border, int m) {        /* F8:567 */
        if (cell[idx] == CELL_SOLID) {
            mask[idx] &= ~m;
            if (mask[idx] == 0) border.push_back(idx);
        }
    }
    void freeEmptyNeighbour(size_t idx, std::vector<size_t>& border) {               /* F8:584 */
        if (cell[idx] == CELL_EMPTY && mask[idx] == 1) {
            mask[idx] = 0;
            border.push_back(idx);
        }
    }
    void extrapolateEmptyBorders() {                                                 /* F8:598 */
        for (int x = 1; x < w - 1; x++) {
            size_t idxT = x, idxB = (size_t)x + (size_t)(h - 1) * w;
            if (cell[idxT] == CELL_EMPTY) src[idxT] = src[idxT + w];
            if (cell[idxB] == CELL_EMPTY) src[idxB] = src[idxB - w];
        }
        for (int y = 1; y < h - 1; y++) {
            size_t idxL = (size_t)y * w, idxR = (size_t)y * w + w - 1;
            if (cell[idxL] == CELL_EMPTY) src[idxL] = src[idxL + 1];
            if (cell[idxR] == CELL_EMPTY) src[idxR] = src[idxR - 1];
        }
        size_t idxTL = 0, idxTR = w - 1, idxBL = (size_t)(h - 1) * w, idxBR = (size_t)h * w - 1;
        if (cell[idxTL] == CELL_EMPTY) src[idxTL] = 0.5 * (src[idxTL + 1] + src[idxTL + w]);
        if (cell[idxTR] == CELL_EMPTY) src[idxTR] = 0.5 * (src[idxTR - 1] + src[idxTR + w]);
        if (cell[idxBL] == CELL_EMPTY) src[idxBL] = 0.5 * (src[idxBL + 1] + src[idxBL - w]);
        if (cell[idxBR] == CELL_EMPTY) src[idxBR] = 0.5 * (src[idxBR - 1] + src[idxBR - w]);
        for (size_t i = 0; i < (size_t)w * h; i++) if (cell[i] == CELL_EMPTY) cell[i] = CELL_FLUID;
    }
    /* F8:651: LIFO stack; an extrapolated EMPTY cell is relabelled FLUID at once, so later averages see it */
    void extrapolateFlip() {
        fillSolidMaskFlip();
        std::vector<size_t> border;
        for (int y = 1; y < h - 1; y++)
            for (int x = 1; x < w - 1; x++) {
                size_t idx = (size_t)x + (size_t)y * w;
                if (cell[idx] != CELL_FLUID && mask[idx] == 0) border.push_back(idx);
            }
        while (!border.empty()) {
            size_t idx = border.back();
            border.pop_back();
            if (cell[idx] == CELL_EMPTY) {
                src[idx] = extrapolateAverage(idx);
                cell[idx] = CELL_FLUID;
            } else {
                src[idx] = extrapolateNormal(idx);
            }
            if (normalX[idx - 1] > 0.0) freeSolidNeighbour(idx - 1, border, 1);
            if (normalX[idx + 1] < 0.0) freeSolidNeighbour(idx + 1, border, 1);
            if (normalY[idx - w] > 0.0) freeSolidNeighbour(idx - w, border, 2);
            if (normalY[idx + w] < 0.0) freeSolidNeighbour(idx + w, border, 2);
            freeEmptyNeighbour(idx - 1, border);
            freeEmptyNeighbour(idx + 1, border);
            freeEmptyNeighbour(idx - w, border);
            freeEmptyNeighbour(idx + w, border);
        }
        extrapolateEmptyBorders();
    }

    /* F1:202 (hard rectangle) / F2:230 (=F3:232, F4:306, F5:440, F7:461, F8:383) smooth.
     * NOTE the inner loop bound is min(ix1, _h) -- _h, not _w -- in every copy. */
    void addInflow(int variant, double x0, double y0, double x1, double y1, double v) {
        int ix0 = jint(x0 / hx - ox);
        int iy0 = jint(y0 / hx - oy);
        int ix1 = jint(x1 / hx - ox);
        int iy1 = jint(y1 / hx - oy);
        for (int y = jmaxi(iy0, 0); y < jmini(iy1, h); y++) {
            for (int x = jmaxi(ix0, 0); x < jmini(ix1, h); x++) {
                size_t i = (size_t)x + (size_t)y * w;
                if (variant == IFL_F1_MATRIXLESS) {
                    if (std::fabs((float)src[i]) < std::fabs((float)v)) src[i] = v;
                } else {
                    double l = length((2.0 * (x + 0.5) * hx - (x0 + x1)) / (x1 - x0),
                                      (2.0 * (y + 0.5) * hx - (y0 + y1)) / (y1 - y0));
                    double vi = cubicPulse(l) * v;
                    if (std::fabs((float)src[i]) < std::fabs((float)vi)) src[i] = vi;
                }
            }
        }
    }
};

/* ---- FluidSolver (F1:235, F3:282) ---------------------------------------------------- */
struct Solver {
    ifl_config cfg;
    int variant, w, h;
    double hx, density;
    Quantity d, t, u, v;
    std::vector<double> r, p, z, s, precon, aDiag, aPlusX, aPlusY;
    std::vector<double> uDensity, vDensity;                            /* F7:757-758 */
    std::vector<Body> bodies;
    ifl_step_status st;
    /* F8 ParticleQuantities (F8:1479) */
    std::vector<double> posX, posY, prop[4];
    std::vector<int> counts;
    int particleCount = 0, maxParticles = 0;
    uint64_t rngCounter = 0;
    double inflow[8] = {0.45, 0.2, 0.2, 0.05, 1.0, 0.0, 0.0, 0.0};   /* F8:1330 (t filled with tAmb in the ctor) */
    bool particlesReady = false;

    explicit Solver(const ifl_config& c) : cfg(c) {
        variant = c.variant; w = c.w; h = c.h; density = c.density;
        hx = 1.0 / jmini(w, h);                                       /* F1:266 */
        d.init(w, h, 0.5, 0.5, hx, variant);                          /* F1:268 */
        u.init(w + 1, h, 0.0, 0.5, hx, variant);                      /* F1:269 */
        v.init(w, h + 1, 0.5, 0.0, hx, variant);                      /* F1:270 */
        if (variant == IFL_F8_FLIP) {
            maxParticles = w * h * cfg.particles_max_per_cell;         /* F8:1549 */
            posX.assign(maxParticles, 0.0); posY.assign(maxParticles, 0.0);
            for (auto& pr : prop) pr.assign(maxParticles, 0.0);
            counts.assign((size_t)w * h, 0);
            inflow[5] = cfg.t_ambient;
        }
        if (variant >= IFL_F6_HEAT) {
            t.init(w, h, 0.5, 0.5, hx, variant);                      /* F6:792 */
            std::fill(t.src.begin(), t.src.end(), cfg.t_ambient);     /* F6:797-799 */
            uDensity.assign((size_t)(w + 1) * h, 0.0);
            vDensity.assign((size_t)w * (h + 1), 0.0);
        }
        size_t n = (size_t)w * h;
        r.assign(n, 0.0); p.assign(n, 0.0);
        z.assign(n, 0.0); s.assign(n, 0.0);
        precon.assign(n, 0.0); aDiag.assign(n, 0.0); aPlusX.assign(n, 0.0); aPlusY.assign(n, 0.0);
        std::memset(&st, 0, sizeof st);
    }

    bool usesPcg() const { return variant >= IFL_F3_CONJUGATE_GRADIENTS; }

    /* F1:364 (=F2:379, F3:420) */
    void buildRhs() {
        double scale = 1.0 / hx;
        size_t idx = 0;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++, idx++)
                r[idx] = -scale * (u.getSrc(x + 1, y) - u.getSrc(x, y) + v.getSrc(x, y + 1) - v.getSrc(x, y));
    }

    /* F1:380 (=F2:395): in-place lexicographic Gauss-Seidel, _p warm-started */
    void projectGS(int limit, double timestep) {
        double scale = timestep / (density * hx * hx);
        double maxDelta = 0.0;
        st.exceeded_budget &= ~1;
        for (int iter = 0; iter < limit; iter++) {
            maxDelta = 0.0;
            size_t idx = 0;
            for (int y = 0; y < h; y++) {
                for (int x = 0; x < w; x++, idx++) {
                    double diag = 0.0, offDiag = 0.0;
                    if (x > 0)     { diag += scale; offDiag -= scale * p[idx - 1]; }
                    if (y > 0)     { diag += scale; offDiag -= scale * p[idx - w]; }
                    if (x < w - 1) { diag += scale; offDiag -= scale * p[idx + 1]; }
                    if (y < h - 1) { diag += scale; offDiag -= scale * p[idx + w]; }
                    double newP = (r[idx] - offDiag) / diag;
                    /* F1:414: the OLD p is rounded through float, the difference is double */
                    maxDelta = jmax(maxDelta, std::fabs(jfloat(p[idx]) - newP));
                    p[idx] = newP;
                }
            }
            if (maxDelta < cfg.tolerance) {
                st.iterations_pressure = iter + 1;
                st.residual_pressure = maxDelta;
                return;
            }
        }
        st.iterations_pressure = limit;
        st.residual_pressure = maxDelta;
        st.exceeded_budget |= 1;
    }

    /* F3:435 */
    void buildPressureMatrix(double timestep) {
        double scale = timestep / (density * hx * hx);
        std::fill(aDiag.begin(), aDiag.end(), 0.0);                  /* F3:438 new double[] */
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (x < w - 1) { aDiag[idx] += scale; aDiag[idx + 1] += scale; aPlusX[idx] = -scale; }
                else           { aPlusX[idx] = 0.0; }
                if (y < h - 1) { aDiag[idx] += scale; aDiag[idx + w] += scale; aPlusY[idx] = -scale; }
                else           { aPlusY[idx] = 0.0; }
            }
        }
    }

    /* F3:464 -- MIC(0), tau = 0.97, sigma = 0.25 */
    void buildPreconditioner() {
        const double tau = cfg.mic_tau, sigma = cfg.mic_sigma;
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                double e = aDiag[idx];
                if (x > 0) {
                    double px = aPlusX[idx - 1] * precon[idx - 1];
                    double py = aPlusY[idx - 1] * precon[idx - 1];
                    e -= (px * px + tau * px * py);
                }
                if (y > 0) {
                    double px = aPlusX[idx - w] * precon[idx - w];
                    double py = aPlusY[idx - w] * precon[idx - w];
                    e -= (py * py + tau * px * py);
                }
                if (e < sigma * aDiag[idx]) e = aDiag[idx];
                precon[idx] = 1.0 / std::sqrt(e);
            }
        }
    }

    /* F3:495 */
    void applyPreconditioner(std::vector<double>& dst, const std::vector<double>& a) {
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                double t = a[idx];
                if (x > 0) t -= aPlusX[idx - 1] * precon[idx - 1] * dst[idx - 1];
                if (y > 0) t -= aPlusY[idx - w] * precon[idx - w] * dst[idx - w];
                dst[idx] = t * precon[idx];
            }
        }
        idx = (size_t)w * h - 1;
        for (int y = h - 1; y >= 0; y--) {
            for (int x = w - 1; x >= 0; x--, idx--) {
                double t = dst[idx];
                if (x < w - 1) t -= aPlusX[idx] * precon[idx] * dst[idx + 1];
                if (y < h - 1) t -= aPlusY[idx] * precon[idx] * dst[idx + w];
                dst[idx] = t * precon[idx];
            }
        }
    }

    /* F3:532 */
    double dotProduct(const std::vector<double>& a, const std::vector<double>& b) {
        double result = 0.0;
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) result += a[i] * b[i];
        return result;
    }

    /* F3:544 */
    void matrixVectorProduct(std::vector<double>& dst, const std::vector<double>& b) {
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                double t = aDiag[idx] * b[idx];
                if (x > 0)     t += aPlusX[idx - 1] * b[idx - 1];
                if (y > 0)     t += aPlusY[idx - w] * b[idx - w];
                if (x < w - 1) t += aPlusX[idx] * b[idx + 1];
                if (y < h - 1) t += aPlusY[idx] * b[idx + w];
                dst[idx] = t;
            }
        }
    }

    /* F3:570 */
    void scaledAdd(std::vector<double>& dst, const std::vector<double>& a, const std::vector<double>& b, double sc) {
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) dst[i] = a[i] + b[i] * sc;
    }

    /* F3:579 -- max |(float) a[i]| */
    double infinityNorm(const std::vector<double>& a) {
        double maxA = 0.0;
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) maxA = jmax(maxA, jabsf(a[i]));
        return maxA;
    }

    /* F3:590 -- MIC(0)-preconditioned conjugate gradients.  Writes iteration count /
     * residual into (its, res); returns false on NaN (reference: System.exit(-1)). */
    bool projectPCG(int limit, int32_t& its, double& res, int budgetBit) {
        std::fill(p.begin(), p.end(), 0.0);                          /* F3:594 new double[] */
        applyPreconditioner(z, r);
        s = z;                                                        /* F3:596 arraycopy */
        st.exceeded_budget &= ~budgetBit;
        double maxError = infinityNorm(r);
        its = 0; res = maxError;
        if (maxError < cfg.tolerance) return true;
        double sigma = dotProduct(z, r);
        for (int iter = 0; iter < limit; iter++) {
            matrixVectorProduct(z, s);
            double alpha = sigma / dotProduct(z, s);
            scaledAdd(p, p, s, alpha);
            scaledAdd(r, r, z, -alpha);
            maxError = infinityNorm(r);
            its = iter + 1; res = maxError;
            if (maxError < cfg.tolerance) return true;
            if (maxError != maxError) { st.nan_residual = 1; return false; }
            applyPreconditioner(z, r);
            double sigmaNew = dotProduct(z, r);
            scaledAdd(s, z, s, sigmaNew / sigma);
            sigma = sigmaNew;
        }
        st.exceeded_budget |= budgetBit;
        return true;
    }

    /* F1:432 (=F2:447, F3:634) */
    void applyPressure(double timestep) {
        double scale = timestep / (density * hx);
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                u.setSrc(x, y, u.getSrc(x, y) - scale * p[idx]);
                u.setSrc(x + 1, y, u.getSrc(x + 1, y) + scale * p[idx]);
                v.setSrc(x, y, v.getSrc(x, y) - scale * p[idx]);
                v.setSrc(x, y + 1, v.getSrc(x, y + 1) + scale * p[idx]);
            }
        }
        for (int y = 0; y < h; y++) { u.setSrc(0, y, 0.0); u.setSrc(w, y, 0.0); }
        for (int x = 0; x < w; x++) { v.setSrc(x, 0, 0.0); v.setSrc(x, h, 0.0); }
    }

    /* ================= F4-F7: solid boundaries, curved boundaries, heat, variable density ============ */
    bool masksVectorOps() const { return variant >= IFL_F6_HEAT; }     /* F6:999-1064 vs F4:825-878 */
    bool fluid(size_t i) const { return d.cell[i] == CELL_FLUID; }

    /* F4:693 (masked) / F5:785 (=F6:804, F7:813) volume weighted with body velocity terms */
    void buildRhsSolid() {
        double scale = 1.0 / hx;
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) { r[idx] = 0.0; continue; }
                if (variant == IFL_F4_SOLID_BOUNDARIES) {
                    r[idx] = -scale * (u.getSrc(x + 1, y) - u.getSrc(x, y) + v.getSrc(x, y + 1) - v.getSrc(x, y));
                    continue;
                }
                r[idx] = -scale * (u.vol(x + 1, y) * u.getSrc(x + 1, y) - u.vol(x, y) * u.getSrc(x, y)
                                   + v.vol(x, y + 1) * v.getSrc(x, y + 1) - v.vol(x, y) * v.getSrc(x, y));
                double vol = d.vol(x, y);
                if (bodies.empty()) continue;
                if (x > 0)     r[idx] -= (u.vol(x, y) - vol) * bodies[d.body[idx - 1]].velocityX(x * hx, (y + 0.5) * hx);
                if (y > 0)     r[idx] -= (v.vol(x, y) - vol) * bodies[d.body[idx - w]].velocityY((x + 0.5) * hx, y * hx);
                if (x < w - 1) r[idx] += (u.vol(x + 1, y) - vol) * bodies[d.body[idx + 1]].velocityX((x + 1.0) * hx, (y + 0.5) * hx);
                if (y < h - 1) r[idx] += (v.vol(x, y + 1) - vol) * bodies[d.body[idx + w]].velocityY((x + 0.5) * hx, (y + 1.0) * hx);
            }
        }
    }

    /* F7:854 */
    void computeDensities() {
        double alpha = (cfg.density_soot - density) / density;
        std::fill(uDensity.begin(), uDensity.end(), 0.0);
        std::fill(vDensity.begin(), vDensity.end(), 0.0);
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                double dens = density * cfg.t_ambient / t.getSrc(x, y) * (1.0 + alpha * d.getSrc(x, y));
                dens = jmax(dens, 0.05 * density);
                uDensity[(size_t)x + (size_t)y * (w + 1)] += 0.5 * dens;
                vDensity[(size_t)x + (size_t)y * w] += 0.5 * dens;
                uDensity[(size_t)(x + 1) + (size_t)y * (w + 1)] += 0.5 * dens;
                vDensity[(size_t)x + (size_t)(y + 1) * w] += 0.5 * dens;
            }
        }
    }

    /* F4:713 (scale) / F5:825 (=F6:844; scale*volume) / F7:877 (scale*volume/density; the y factor indexes
     * _vDensity with _u.I(x, y+1), i.e. stride w+1 -- kept) */
    void buildPressureMatrixSolid(double timestep) {
        double scale = variant >= IFL_F7_VARIABLE_DENSITY ? timestep / (hx * hx) : timestep / (density * hx * hx);
        std::fill(aDiag.begin(), aDiag.end(), 0.0);
        std::fill(aPlusX.begin(), aPlusX.end(), 0.0);
        std::fill(aPlusY.begin(), aPlusY.end(), 0.0);
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) continue;
                if (x < w - 1 && fluid(idx + 1)) {
                    double factor = scale;
                    if (variant >= IFL_F7_VARIABLE_DENSITY) factor = scale * u.vol(x + 1, y) / uDensity[(size_t)(x + 1) + (size_t)y * (w + 1)];
                    else if (variant >= IFL_F5_CURVED_BOUNDARIES) factor = scale * u.vol(x + 1, y);
                    aDiag[idx] += factor; aDiag[idx + 1] += factor; aPlusX[idx] = -factor;
                }
                if (y < h - 1 && fluid(idx + w)) {
                    double factor = scale;
                    if (variant >= IFL_F7_VARIABLE_DENSITY) factor = scale * v.vol(x, y + 1) / vDensity[(size_t)x + (size_t)(y + 1) * (w + 1)];
                    else if (variant >= IFL_F5_CURVED_BOUNDARIES) factor = scale * v.vol(x, y + 1);
                    aDiag[idx] += factor; aDiag[idx + w] += factor; aPlusY[idx] = -factor;
                }
            }
        }
    }

    /* F6:880 (=F7:913) */
    void buildHeatDiffusionMatrix(double timestep) {
        std::fill(aDiag.begin(), aDiag.end(), 1.0);
        std::fill(aPlusX.begin(), aPlusX.end(), 0.0);
        std::fill(aPlusY.begin(), aPlusY.end(), 0.0);
        double scale = cfg.diffusion * timestep * 1.0 / (hx * hx);
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) continue;
                if (x < w - 1 && fluid(idx + 1)) { aDiag[idx] += scale; aDiag[idx + 1] += scale; aPlusX[idx] = -scale; }
                if (y < h - 1 && fluid(idx + w)) { aDiag[idx] += scale; aDiag[idx + w] += scale; aPlusY[idx] = -scale; }
            }
        }
    }

    /* F4:744 (=F5:858, F6:918, F7:951) */
    void buildPreconditionerSolid() {
        const double tau = cfg.mic_tau, sigma = cfg.mic_sigma;
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) continue;
                double e = aDiag[idx];
                if (x > 0 && fluid(idx - 1)) {
                    double px = aPlusX[idx - 1] * precon[idx - 1];
                    double py = aPlusY[idx - 1] * precon[idx - 1];
                    e -= (px * px + tau * px * py);
                }
                if (y > 0 && fluid(idx - w)) {
                    double px = aPlusX[idx - w] * precon[idx - w];
                    double py = aPlusY[idx - w] * precon[idx - w];
                    e -= (py * py + tau * px * py);
                }
                if (e < sigma * aDiag[idx]) e = aDiag[idx];
                precon[idx] = 1.0 / std::sqrt(e);
            }
        }
    }

    /* F4:780 (=F5:894, F6:954, F7:987) */
    void applyPreconditionerSolid(std::vector<double>& dst, const std::vector<double>& a) {
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) continue;
                double tt = a[idx];
                if (x > 0 && fluid(idx - 1)) tt -= aPlusX[idx - 1] * precon[idx - 1] * dst[idx - 1];
                if (y > 0 && fluid(idx - w)) tt -= aPlusY[idx - w] * precon[idx - w] * dst[idx - w];
                dst[idx] = tt * precon[idx];
            }
        }
        idx = (size_t)w * h - 1;
        for (int y = h - 1; y >= 0; y--) {
            for (int x = w - 1; x >= 0; x--, idx--) {
                if (!fluid(idx)) continue;
                double tt = dst[idx];
                if (x < w - 1 && fluid(idx + 1)) tt -= aPlusX[idx] * precon[idx] * dst[idx + 1];
                if (y < h - 1 && fluid(idx + w)) tt -= aPlusY[idx] * precon[idx] * dst[idx + w];
                dst[idx] = tt * precon[idx];
            }
        }
    }

    double dotProductM(const std::vector<double>& a, const std::vector<double>& b) {     /* F6:999 */
        double result = 0.0;
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) if (!masksVectorOps() || fluid(i)) result += a[i] * b[i];
        return result;
    }
    void scaledAddM(std::vector<double>& dst, const std::vector<double>& a, const std::vector<double>& b, double sc) {   /* F6:1041 */
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) if (!masksVectorOps() || fluid(i)) dst[i] = a[i] + b[i] * sc;
    }
    double infinityNormM(const std::vector<double>& a) {                                  /* F6:1054 */
        double maxA = 0.0;
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) if (!masksVectorOps() || fluid(i)) maxA = jmax(maxA, jabsf(a[i]));
        return maxA;
    }

    /* F4:883 (=F5:997, F6:1069, F7:1102) */
    bool projectPCGSolid(int limit, int32_t& its, double& res, int budgetBit) {
        std::fill(p.begin(), p.end(), 0.0);
        applyPreconditionerSolid(z, r);
        s = z;
        st.exceeded_budget &= ~budgetBit;
        double maxError = infinityNormM(r);
        its = 0; res = maxError;
        if (maxError < cfg.tolerance) return true;
        double sigma = dotProductM(z, r);
        for (int iter = 0; iter < limit; iter++) {
            matrixVectorProduct(z, s);
            double alpha = sigma / dotProductM(z, s);
            scaledAddM(p, p, s, alpha);
            scaledAddM(r, r, z, -alpha);
            maxError = infinityNormM(r);
            its = iter + 1; res = maxError;
            if (maxError < cfg.tolerance) return true;
            if (maxError != maxError) { st.nan_residual = 1; return false; }
            applyPreconditionerSolid(z, r);
            double sigmaNew = dotProductM(z, r);
            scaledAddM(s, z, s, sigmaNew / sigma);
            sigma = sigmaNew;
        }
        st.exceeded_budget |= budgetBit;
        return true;
    }

    /* F4:927 (=F5:1041, F6:1114) / F7:1148 (divided by the face density) */
    void applyPressureSolid(double timestep) {
        const bool vd = variant >= IFL_F7_VARIABLE_DENSITY;
        double scale = vd ? timestep / hx : timestep / (density * hx);
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (!fluid(idx)) continue;
                if (vd) {
                    u.setSrc(x, y, u.getSrc(x, y) - scale * p[idx] / uDensity[(size_t)x + (size_t)y * (w + 1)]);
                    v.setSrc(x, y, v.getSrc(x, y) - scale * p[idx] / vDensity[(size_t)x + (size_t)y * w]);
                    u.setSrc(x + 1, y, u.getSrc(x + 1, y) + scale * p[idx] / uDensity[(size_t)(x + 1) + (size_t)y * (w + 1)]);
                    v.setSrc(x, y + 1, v.getSrc(x, y + 1) + scale * p[idx] / vDensity[(size_t)x + (size_t)(y + 1) * w]);
                } else {
                    u.setSrc(x, y, u.getSrc(x, y) - scale * p[idx]);
                    u.setSrc(x + 1, y, u.getSrc(x + 1, y) + scale * p[idx]);
                    v.setSrc(x, y, v.getSrc(x, y) - scale * p[idx]);
                    v.setSrc(x, y + 1, v.getSrc(x, y + 1) + scale * p[idx]);
                }
            }
        }
    }

    /* F6:1135 (=F7:1169) */
    void addBuoyancy(double timestep) {
        double alpha = (cfg.density_soot - density) / density;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                double buoyancy = timestep * cfg.gravity * (alpha * d.getSrc(x, y) - (t.getSrc(x, y) - cfg.t_ambient) / cfg.t_ambient);
                v.setSrc(x, y, v.getSrc(x, y) + buoyancy * 0.5);
                v.setSrc(x, y + 1, v.getSrc(x, y + 1) + buoyancy * 0.5);
            }
        }
    }

    /* F4:948 (=F5:1062, F6:1154, F7:1185): all four faces use velocityX (kept) */
    void setBoundaryCondition() {
        size_t idx = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++, idx++) {
                if (d.cell[idx] == CELL_SOLID) {
                    const Body& b = bodies[d.body[idx]];
                    u.setSrc(x, y, b.velocityX(x * hx, (y + 0.5) * hx));
                    v.setSrc(x, y, b.velocityX((x + 0.5) * hx, y * hx));
                    u.setSrc(x + 1, y, b.velocityX((x + 1.0) * hx, (y + 0.5) * hx));
                    v.setSrc(x, y + 1, b.velocityX((x + 0.5) * hx, (y + 1.0) * hx));
                }
            }
        }
        for (int y = 0; y < h; y++) { u.setSrc(0, y, 0.0); u.setSrc(w, y, 0.0); }
        for (int x = 0; x < w; x++) { v.setSrc(x, 0, 0.0); v.setSrc(x, h, 0.0); }
    }

    void fillSolidFieldsAll() {
        d.fillSolidFields(variant, bodies);
        if (variant >= IFL_F6_HEAT) t.fillSolidFields(variant, bodies);
        u.fillSolidFields(variant, bodies);
        v.fillSolidFields(variant, bodies);
    }

    /* F6:1190-1201 (=F7:1221-1232): r <- T; heat matrix; MIC; project; T <- p */
    void heatSolve(double timestep) {
        r = t.src;
        buildHeatDiffusionMatrix(timestep);
        buildPreconditionerSolid();
        projectPCGSolid(cfg.solver_limit, st.iterations_heat, st.residual_heat, 2);
        t.src = p;
    }

    /* F4:617 / F5:1089 / F6:1181 / F7:1212 */
    void updateSolid(double timestep) {
        fillSolidFieldsAll();
        if (variant >= IFL_F6_HEAT) {
            heatSolve(timestep);
            t.extrapolate();
            addBuoyancy(timestep);
        }
        setBoundaryCondition();
        buildRhsSolid();
        if (variant >= IFL_F7_VARIABLE_DENSITY) computeDensities();
        buildPressureMatrixSolid(timestep);
        buildPreconditionerSolid();
        projectPCGSolid(cfg.solver_limit, st.iterations_pressure, st.residual_pressure, 1);
        applyPressureSolid(timestep);
        d.extrapolate(); u.extrapolate(); v.extrapolate();
        setBoundaryCondition();
        d.advectMasked(timestep, u, v, bodies);
        if (variant >= IFL_F6_HEAT) t.advectMasked(timestep, u, v, bodies);
        u.advectMasked(timestep, u, v, bodies);
        v.advectMasked(timestep, u, v, bodies);
        d.swap(); u.swap(); v.swap();
        if (variant >= IFL_F6_HEAT) t.swap();
    }

    /* ================= F8: FLIP / PIC (ParticleQuantities F8:1479-1790) ================= */
    Quantity* qOf(int q) { return q == 0 ? &d : q == 1 ? &t : q == 2 ? &u : &v; }      /* addQuantity order F8:883-886 */
    float nextRandom() { return rng_uniform(cfg.rng_seed, rngCounter++); }
    bool pointInBody(double x, double y) const {                                       /* F8:1564 */
        for (const Body& b : bodies) if (b.distance(x * hx, y * hx) < 0.0) return true;
        return false;
    }
    /* F8:1571: _AvgPerCell tries per cell, rejected ones are dropped; x + (float) random() is a FLOAT addition */
    void initParticles() {
        int idx = 0;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                for (int i = 0; i < cfg.particles_avg_per_cell; i++, idx++) {
                    posX[idx] = (double)((float)x + nextRandom());
                    posY[idx] = (double)((float)y + nextRandom());
                    if (pointInBody(posX[idx], posY[idx])) idx--;
                }
        particleCount = idx;
    }
    void countParticles() {                                                            /* F8:1595 */
        std::fill(counts.begin(), counts.end(), 0);
        for (int i = 0; i < particleCount; i++) {
            int ix = jint(posX[i]), iy = jint(posY[i]);
            if (ix >= 0 && iy >= 0 && ix < w && iy < h) counts[(size_t)ix + (size_t)iy * w]++;
        }
    }
    void pruneParticles() {                                                            /* F8:1610 (the range guard uses && and never fires) */
        for (int i = 0; i < particleCount; i++) {
            int ix = jint(posX[i]), iy = jint(posY[i]);
            size_t idx = (size_t)ix + (size_t)iy * w;
            if (counts[idx] > cfg.particles_max_per_cell) {
                int j = --particleCount;
                posX[i] = posX[j]; posY[i] = posY[j];
                for (auto& pr : prop) pr[i] = pr[j];
                counts[idx]--;
                i--;
            }
        }
    }
    void seedParticles() {                                                             /* F8:1637 */
        size_t idx = 0;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++, idx++)
                for (int i = 0; i < cfg.particles_min_per_cell - counts[idx]; i++) {
                    if (particleCount == maxParticles) return;
                    int j = particleCount;
                    posX[j] = (double)((float)x + nextRandom());
                    posY[j] = (double)((float)y + nextRandom());
                    if (pointInBody(posX[idx], posY[idx])) continue;                   /* F8:1653: tests particle `idx`, not j (kept) */
                    for (int q = 0; q < 4; q++) prop[q][j] = qOf(q)->lerpGrid(posX[j], posY[j]);
                    particleCount++;
                }
    }
    Vec2 particleBackProject(double x, double y) const {                               /* F8:1673: threshold d < -1.0 (kept) */
        double dd = 1e30;
        int closest = -1;
        for (size_t i = 0; i < bodies.size(); i++) {
            double id = bodies[i].distance(x * hx, y * hx);
            if (id < dd) { dd = id; closest = (int)i; }
        }
        if (dd < -1.0) {
            x *= hx; y *= hx;
            Vec2 r = bodies[closest].closestSurfacePoint(x, y);
            x = r.x; y = r.y;
            Vec2 n = bodies[closest].distanceNormal(x, y);
            x -= n.x * hx; y -= n.y * hx;
            x /= hx; y /= hx;
        }
        return {x, y};
    }
    Vec2 particleRK3(double x, double y, double timestep) const {                      /* F8:1708: forward in time */
        double firstU = u.lerpGrid(x, y) / hx;
        double firstV = v.lerpGrid(x, y) / hx;
        double midX = x + 0.5 * timestep * firstU;
        double midY = y + 0.5 * timestep * firstV;
        double midU = u.lerpGrid(midX, midY) / hx;
        double midV = v.lerpGrid(midX, midY) / hx;
        double lastX = x + 0.75 * timestep * midU;
        double lastY = y + 0.75 * timestep * midV;
        double lastU = u.lerpGrid(lastX, lastY);
        double lastV = v.lerpGrid(lastX, lastY);
        x += timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
        y += timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
        return {x, y};
    }
    void gridToParticles(double alpha) {                                               /* F8:1748 */
        for (int q = 0; q < 4; q++)
            for (int i = 0; i < particleCount; i++) {
                prop[q][i] *= 1.0 - alpha;
                prop[q][i] += qOf(q)->lerpGrid(posX[i], posY[i]);
            }
    }
    void particlesToGrid() {                                                           /* F8:1761 */
        for (int q = 0; q < 4; q++) {
            qOf(q)->fromParticles(particleCount, posX, posY, prop[q]);
            qOf(q)->extrapolateFlip();
        }
        countParticles();
        pruneParticles();
        seedParticles();
        st.particle_count = particleCount;
    }
    void advectParticles(double timestep) {                                            /* F8:1778 */
        for (int i = 0; i < particleCount; i++) {
            Vec2 r = particleRK3(posX[i], posY[i], timestep);
            r = particleBackProject(r.x, r.y);
            posX[i] = jmax(jmin(r.x, w - 0.001), 0.0);
            posY[i] = jmax(jmin(r.y, h - 0.001), 0.0);
        }
    }
    /* the part of the F8 constructor that needs the body list: ParticleQuantities ctor + gridToParticles(1.0) (F8:887-891) */
    void ensureParticles() {
        if (particlesReady) return;
        initParticles();
        gridToParticles(1.0);
        particlesReady = true;
        st.particle_count = particleCount;
    }
    void heatSolveFlip(double timestep) {
        r = t.src;
        buildHeatDiffusionMatrix(timestep);
        buildPreconditionerSolid();
        projectPCGSolid(cfg.solver_limit, st.iterations_heat, st.residual_heat, 2);
        t.src = p;
    }
    /* F8:1306 */
    void updateFlip(double timestep) {
        ensureParticles();
        fillSolidFieldsAll();
        particlesToGrid();
        d.copyOld(); t.copyOld(); u.copyOld(); v.copyOld();
        addInflow(inflow[0], inflow[1], inflow[2], inflow[3], inflow[4], inflow[5], inflow[6], inflow[7]);   /* F8:1330 */
        heatSolveFlip(timestep);
        t.extrapolateFlip();
        addBuoyancy(timestep);
        setBoundaryCondition();
        buildRhsSolid();
        computeDensities();
        buildPressureMatrixSolid(timestep);
        buildPreconditionerSolid();
        projectPCGSolid(cfg.solver_limit, st.iterations_pressure, st.residual_pressure, 1);
        applyPressureSolid(timestep);
        d.extrapolateFlip(); u.extrapolateFlip(); v.extrapolateFlip();
        setBoundaryCondition();
        const double a = cfg.flip_alpha;
        d.diff(a); t.diff(a); u.diff(a); v.diff(a);
        gridToParticles(a);
        d.undiff(a); t.undiff(a); u.undiff(a); v.undiff(a);
        advectParticles(timestep);
    }

    void project(double timestep) {
        if (variant >= IFL_F4_SOLID_BOUNDARIES) { projectPCGSolid(cfg.solver_limit, st.iterations_pressure, st.residual_pressure, 1); return; }
        if (usesPcg()) projectPCG(cfg.solver_limit, st.iterations_pressure, st.residual_pressure, 1);
        else           projectGS(cfg.solver_limit, timestep);
    }

    /* F1:298 / F3:385 */
    void addInflow(double x, double y, double ww, double hh, double dv, double tv, double uv, double vv) {
        d.addInflow(variant, x, y, x + ww, y + hh, dv);
        if (variant >= IFL_F6_HEAT) t.addInflow(variant, x, y, x + ww, y + hh, tv);       /* F6:1233 */
        u.addInflow(variant, x, y, x + ww, y + hh, uv);
        v.addInflow(variant, x, y, x + ww, y + hh, vv);
    }

    /* F1:276 / F2:322 / F3:361 */
    void update(double timestep) {
        if (variant == IFL_F8_FLIP) { updateFlip(timestep); return; }
        if (variant >= IFL_F4_SOLID_BOUNDARIES) { updateSolid(timestep); return; }
        buildRhs();
        if (usesPcg()) { buildPressureMatrix(timestep); buildPreconditioner(); }
        project(timestep);
        applyPressure(timestep);
        d.advect(variant, timestep, u, v);
        u.advect(variant, timestep, u, v);
        v.advect(variant, timestep, u, v);
        d.swap(); u.swap(); v.swap();
    }

    /* F1:309 */
    double maxTimestep() {
        double maxVelocity = 0.0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                double uu = u.lerpGrid(x + 0.5, y + 0.5);
                double vv = v.lerpGrid(x + 0.5, y + 0.5);
                double velocity = std::sqrt(uu * uu + vv * vv);
                maxVelocity = jmax(maxVelocity, velocity);
            }
        }
        double maxTimestep = 2.0 * hx / maxVelocity;
        return jmin(maxTimestep, 1.0);
    }

    /* toImage: the integers of one frame, rgb[3 * pixels], pixels = (renderHeat ? 2w : w) * h.
     * F1:341 (=F2:356, F3:397), F4:664, F5:1133; toImage(destination, renderHeat) F6:1250 (=F7:1282), F8:1417. */
    void toImage(bool renderHeat, int32_t* rgb) const {
        if (variant < IFL_F6_HEAT) {
            for (int i = 0; i < w * h; i++) {
                double val = 1.0 - d.src[i];
                if (variant == IFL_F5_CURVED_BOUNDARIES) val = val * d.vol(i % w, i / w);        /* F5:1142 */
                int shade = jint(val * 255.0);                                                  /* F1:351 */
                shade = jmaxi(jmini(shade, 255), 0);                                            /* F1:352 */
                if (variant == IFL_F4_SOLID_BOUNDARIES && d.cell[i] == CELL_SOLID) shade = 0;   /* F4:678 */
                rgb[3 * i] = shade; rgb[3 * i + 1] = shade; rgb[3 * i + 2] = shade;
            }
            return;
        }
        const size_t fw = renderHeat ? 2 * (size_t)w : (size_t)w;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                const size_t idxl = 3 * ((size_t)x + (size_t)y * fw);                            /* F6:1263 */
                const size_t idxr = renderHeat ? 3 * ((size_t)x + (size_t)y * fw + w) : idxl;    /* F6:1264,1266 */
                double volume = d.vol(x, y);
                double shade = (1.0 - d.getSrc(x, y)) * volume;                                  /* F6:1271 */
                shade = jmin(jmax(shade, 0.0), 1.0);
                rgb[idxr + 0] = jint(shade * 255.0);
                rgb[idxr + 1] = jint(shade * 255.0);
                rgb[idxr + 2] = jint(shade * 255.0);
                if (variant == IFL_F8_FLIP && d.cell[(size_t)x + (size_t)y * w] == CELL_EMPTY) { /* F8:1444 */
                    rgb[idxr] = 0xFF;
                    rgb[idxr + 1] = rgb[idxr + 2] = 0;
                }
                if (renderHeat) {
                    double tt;
                    if (variant == IFL_F8_FLIP) tt = std::fabs(jfloat(t.getSrc(x, y)) - cfg.t_ambient) / 70.0;   /* F8:1450 */
                    else tt = (t.getSrc(x, y) - cfg.t_ambient) / 700.0;                          /* F6:1279 */
                    tt = jmin(jmax(tt, 0.0), 1.0);
                    double r = 1.0 + volume * (jmin(tt * 4.0, 1.0) - 1.0);
                    double g = 1.0 + volume * (jmin(tt * 2.0, 1.0) - 1.0);
                    double b = 1.0 + volume * (jmax(jmin(tt * 4.0 - 3.0, 1.0), 0.0) - 1.0);
                    rgb[idxl + 0] = jint(r * 255.0);
                    rgb[idxl + 1] = jint(g * 255.0);
                    rgb[idxl + 2] = jint(b * 255.0);
                }
            }
        }
    }
    /* the text the reference writes (F1:344-357 / F6:1255-1299): note `i % _w` with the SOLVER width */
    std::string ppmText(bool renderHeat) const {
        const size_t pixels = (size_t)(renderHeat ? 2 * w : w) * h;
        std::vector<int32_t> rgb(3 * pixels);
        toImage(renderHeat, rgb.data());
        std::string out = "P3\n" + std::to_string(renderHeat ? w * 2 : w) + " " + std::to_string(h) + "\n" + std::to_string(255) + "\n";
        for (size_t i = 0; i < pixels; i++) {
            out += std::to_string(rgb[3 * i]) + " " + std::to_string(rgb[3 * i + 1]) + " " + std::to_string(rgb[3 * i + 2]) + " ";
            if (i % (size_t)w == 0) out += "\n";
        }
        return out;
    }

    Quantity* quantity(int q) { return q == 0 ? &d : q == 1 ? &t : q == 2 ? &u : &v; }

    std::vector<double>* field(int f) {
        const bool solids = variant >= IFL_F4_SOLID_BOUNDARIES, heat = variant >= IFL_F6_HEAT;
        switch (f) {
            case IFL_FIELD_D: return &d.src;   case IFL_FIELD_U: return &u.src;   case IFL_FIELD_V: return &v.src;
            case IFL_FIELD_D_DST: return &d.dst; case IFL_FIELD_U_DST: return &u.dst; case IFL_FIELD_V_DST: return &v.dst;
            case IFL_FIELD_T: return heat ? &t.src : nullptr;
            case IFL_FIELD_T_DST: return heat ? &t.dst : nullptr;
            case IFL_FIELD_P: return &p; case IFL_FIELD_R: return &r; case IFL_FIELD_Z: return &z; case IFL_FIELD_S: return &s;
            case IFL_FIELD_PRECON: return &precon; case IFL_FIELD_ADIAG: return &aDiag;
            case IFL_FIELD_APLUSX: return &aPlusX; case IFL_FIELD_APLUSY: return &aPlusY;
            case IFL_FIELD_UDENSITY: return variant >= IFL_F7_VARIABLE_DENSITY ? &uDensity : nullptr;
            case IFL_FIELD_VDENSITY: return variant >= IFL_F7_VARIABLE_DENSITY ? &vDensity : nullptr;
            default: break;
        }
        if (!solids || f < IFL_FIELD_VOLUME_D || f > IFL_FIELD_PHI_V) return nullptr;
        int q = (f - IFL_FIELD_VOLUME_D) % 4, grp = (f - IFL_FIELD_VOLUME_D) / 4;
        if (q == 1 && !heat) return nullptr;
        Quantity* Q = quantity(q);
        switch (grp) {
            case 0: return variant >= IFL_F5_CURVED_BOUNDARIES ? &Q->volume : nullptr;
            case 1: return &Q->normalX;
            case 2: return &Q->normalY;
            case 5: return variant >= IFL_F5_CURVED_BOUNDARIES ? &Q->phi : nullptr;
            default: return nullptr;   /* 3, 4: integer arrays, see int_field */
        }
    }
    std::vector<int>* int_field(int f) {
        if (variant < IFL_F4_SOLID_BOUNDARIES || f < IFL_FIELD_CELL_D || f > IFL_FIELD_BODY_V) return nullptr;
        int q = (f - IFL_FIELD_CELL_D) % 4;
        if (q == 1 && variant < IFL_F6_HEAT) return nullptr;
        Quantity* Q = quantity(q);
        return f < IFL_FIELD_BODY_D ? &Q->cell : &Q->body;
    }
};

} // namespace

extern "C" {

int ifo_create(const ifl_config* cfg, void** out) {
    if (!cfg || !out || cfg->w < 2 || cfg->h < 2) return IFL_ERR_INVALID;
    if (cfg->variant < IFL_F1_MATRIXLESS || cfg->variant > IFL_F8_FLIP) return IFL_ERR_UNSUPPORTED;
    *out = new Solver(*cfg);
    return IFL_OK;
}
void ifo_destroy(void* h) { delete (Solver*)h; }

int ifo_set_bodies(void* h, const ifl_body* bodies, int32_t count) {
    Solver* s = (Solver*)h;
    if (s->variant < IFL_F4_SOLID_BOUNDARIES) return count == 0 ? IFL_OK : IFL_ERR_UNSUPPORTED;
    s->bodies.clear();
    for (int i = 0; i < count; i++) s->bodies.push_back(Body{bodies[i]});
    return IFL_OK;
}
int ifo_add_inflow(void* h, double x, double y, double w, double hh, double d, double t, double u, double v) {
    Solver* s = (Solver*)h;
    if (s->variant == IFL_F8_FLIP) {           /* F8 calls addInflow inside update() (F8:1330): this sets that rectangle */
        double r[8] = {x, y, w, hh, d, t, u, v};
        std::memcpy(s->inflow, r, sizeof r);
        return IFL_OK;
    }
    s->addInflow(x, y, w, hh, d, t, u, v);
    return IFL_OK;
}
int ifo_particle_count(void* h, int32_t* count) {
    Solver* s = (Solver*)h;
    if (s->variant != IFL_F8_FLIP) return IFL_ERR_UNSUPPORTED;
    s->ensureParticles();
    *count = s->particleCount;
    return IFL_OK;
}
int ifo_read_particles(void* h, double* px, double* py, double* pd, double* pt, double* pu, double* pv, size_t capacity) {
    Solver* s = (Solver*)h;
    if (s->variant != IFL_F8_FLIP) return IFL_ERR_UNSUPPORTED;
    s->ensureParticles();
    if (capacity < (size_t)s->particleCount) return IFL_ERR_INVALID;
    size_t n = (size_t)s->particleCount * sizeof(double);
    std::memcpy(px, s->posX.data(), n); std::memcpy(py, s->posY.data(), n);
    std::memcpy(pd, s->prop[0].data(), n); std::memcpy(pt, s->prop[1].data(), n);
    std::memcpy(pu, s->prop[2].data(), n); std::memcpy(pv, s->prop[3].data(), n);
    return IFL_OK;
}
int ifo_write_particles(void* h, const double* px, const double* py, const double* pd, const double* pt, const double* pu,
                        const double* pv, size_t count) {
    Solver* s = (Solver*)h;
    if (s->variant != IFL_F8_FLIP || count > (size_t)s->maxParticles) return IFL_ERR_INVALID;
    size_t n = count * sizeof(double);
    std::memcpy(s->posX.data(), px, n); std::memcpy(s->posY.data(), py, n);
    std::memcpy(s->prop[0].data(), pd, n); std::memcpy(s->prop[1].data(), pt, n);
    std::memcpy(s->prop[2].data(), pu, n); std::memcpy(s->prop[3].data(), pv, n);
    s->particleCount = (int)count;
    s->particlesReady = true;
    s->st.particle_count = s->particleCount;
    return IFL_OK;
}
int ifo_update(void* h, double dt, ifl_step_status* st) {
    Solver* s = (Solver*)h;
    s->update(dt);
    if (st) *st = s->st;
    return s->st.nan_residual ? IFL_ERR_NAN : IFL_OK;
}
int ifo_run_steps(void* h, int n, double dt, double x, double y, double w, double hh,
                  double d, double t, double u, double v, ifl_step_status* st) {
    Solver* s = (Solver*)h;
    for (int i = 0; i < n; i++) {
        s->addInflow(x, y, w, hh, d, t, u, v);
        s->update(dt);
        if (s->st.nan_residual) break;
    }
    if (st) *st = s->st;
    return s->st.nan_residual ? IFL_ERR_NAN : IFL_OK;
}
int ifo_run_stage(void* h, int stage, double dt, ifl_step_status* st) {
    Solver* s = (Solver*)h;
    const bool solids = s->variant >= IFL_F4_SOLID_BOUNDARIES, heat = s->variant >= IFL_F6_HEAT, flip = s->variant == IFL_F8_FLIP;
    auto adv = [&](Quantity& q) { if (solids) q.advectMasked(dt, s->u, s->v, s->bodies); else q.advect(s->variant, dt, s->u, s->v); };
    switch (stage) {
        case IFL_STAGE_BUILD_RHS: if (solids) s->buildRhsSolid(); else s->buildRhs(); break;
        case IFL_STAGE_BUILD_MATRIX:
            if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED;
            if (solids) s->buildPressureMatrixSolid(dt); else s->buildPressureMatrix(dt); break;
        case IFL_STAGE_BUILD_PRECON:
            if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED;
            if (solids) s->buildPreconditionerSolid(); else s->buildPreconditioner(); break;
        case IFL_STAGE_PROJECT: s->project(dt); break;
        case IFL_STAGE_APPLY_PRESSURE: if (solids) s->applyPressureSolid(dt); else s->applyPressure(dt); break;
        case IFL_STAGE_ADVECT_D: if (flip) return IFL_ERR_UNSUPPORTED; adv(s->d); break;
        case IFL_STAGE_ADVECT_T: if (!heat || flip) return IFL_ERR_UNSUPPORTED; adv(s->t); break;
        case IFL_STAGE_ADVECT_U: if (flip) return IFL_ERR_UNSUPPORTED; adv(s->u); break;
        case IFL_STAGE_ADVECT_V: if (flip) return IFL_ERR_UNSUPPORTED; adv(s->v); break;
        case IFL_STAGE_SWAP: if (flip) return IFL_ERR_UNSUPPORTED; s->d.swap(); s->u.swap(); s->v.swap(); if (heat) s->t.swap(); break;
        case IFL_STAGE_APPLY_PRECON:
            if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED;
            if (solids) s->applyPreconditionerSolid(s->z, s->r); else s->applyPreconditioner(s->z, s->r); break;
        case IFL_STAGE_MATVEC: if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED; s->matrixVectorProduct(s->z, s->s); break;
        case IFL_STAGE_FILL_SOLID_FIELDS: if (!solids) return IFL_ERR_UNSUPPORTED; s->fillSolidFieldsAll(); break;
        case IFL_STAGE_SET_BOUNDARY: if (!solids) return IFL_ERR_UNSUPPORTED; s->setBoundaryCondition(); break;
        case IFL_STAGE_EXTRAPOLATE_D: if (!solids) return IFL_ERR_UNSUPPORTED; if (flip) s->d.extrapolateFlip(); else s->d.extrapolate(); break;
        case IFL_STAGE_EXTRAPOLATE_T: if (!heat) return IFL_ERR_UNSUPPORTED; if (flip) s->t.extrapolateFlip(); else s->t.extrapolate(); break;
        case IFL_STAGE_EXTRAPOLATE_U: if (!solids) return IFL_ERR_UNSUPPORTED; if (flip) s->u.extrapolateFlip(); else s->u.extrapolate(); break;
        case IFL_STAGE_EXTRAPOLATE_V: if (!solids) return IFL_ERR_UNSUPPORTED; if (flip) s->v.extrapolateFlip(); else s->v.extrapolate(); break;
        case IFL_STAGE_PARTICLES_TO_GRID: if (!flip) return IFL_ERR_UNSUPPORTED; s->ensureParticles(); s->particlesToGrid(); break;
        case IFL_STAGE_GRID_TO_PARTICLES: {
            if (!flip) return IFL_ERR_UNSUPPORTED;
            s->ensureParticles();
            const double a = s->cfg.flip_alpha;
            s->d.diff(a); s->t.diff(a); s->u.diff(a); s->v.diff(a);
            s->gridToParticles(a);
            s->d.undiff(a); s->t.undiff(a); s->u.undiff(a); s->v.undiff(a);
            break;
        }
        case IFL_STAGE_ADVECT_PARTICLES: if (!flip) return IFL_ERR_UNSUPPORTED; s->ensureParticles(); s->advectParticles(dt); break;
        case IFL_STAGE_HEAT_SOLVE: if (!heat) return IFL_ERR_UNSUPPORTED; s->heatSolve(dt); break;
        case IFL_STAGE_ADD_BUOYANCY: if (!heat) return IFL_ERR_UNSUPPORTED; s->addBuoyancy(dt); break;
        case IFL_STAGE_COMPUTE_DENSITIES: if (s->variant < IFL_F7_VARIABLE_DENSITY) return IFL_ERR_UNSUPPORTED; s->computeDensities(); break;
        default: return IFL_ERR_UNSUPPORTED;
    }
    if (st) *st = s->st;
    return IFL_OK;
}
int ifo_max_timestep(void* h, double* out) { *out = ((Solver*)h)->maxTimestep(); return IFL_OK; }
int ifo_image_size(void* h, int render_heat, size_t* pixels) {
    Solver* s = (Solver*)h;
    if (render_heat && s->variant < IFL_F6_HEAT) return IFL_ERR_INVALID;
    *pixels = (size_t)(render_heat ? 2 * s->w : s->w) * s->h;
    return IFL_OK;
}
int ifo_to_image(void* h, int render_heat, int32_t* rgb, size_t pixels) {
    size_t want = 0;
    if (ifo_image_size(h, render_heat, &want) != IFL_OK || want != pixels) return IFL_ERR_INVALID;
    ((Solver*)h)->toImage(render_heat != 0, rgb);
    return IFL_OK;
}
int ifo_write_ppm(void* h, const char* path, int render_heat) {
    size_t want = 0;
    if (ifo_image_size(h, render_heat, &want) != IFL_OK) return IFL_ERR_INVALID;
    std::string text = ((Solver*)h)->ppmText(render_heat != 0);
    FILE* f = fopen(path, "wb");
    if (!f) return IFL_ERR_INVALID;
    bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
    fclose(f);
    return ok ? IFL_OK : IFL_ERR_INVALID;
}

int ifo_field_size(void* h, int f, size_t* count) {
    Solver* s = (Solver*)h;
    if (auto* v = s->field(f)) { *count = v->size(); return IFL_OK; }
    if (auto* v = s->int_field(f)) { *count = v->size(); return IFL_OK; }
    return IFL_ERR_INVALID;
}
int ifo_read_field(void* h, int f, double* host, size_t count) {
    Solver* s = (Solver*)h;
    if (auto* v = s->field(f)) {
        if (count != v->size()) return IFL_ERR_INVALID;
        std::memcpy(host, v->data(), count * sizeof(double));
        return IFL_OK;
    }
    if (auto* v = s->int_field(f)) {
        if (count != v->size()) return IFL_ERR_INVALID;
        for (size_t i = 0; i < count; i++) host[i] = (double)(*v)[i];
        return IFL_OK;
    }
    return IFL_ERR_INVALID;
}
int ifo_write_field(void* h, int f, const double* host, size_t count) {
    Solver* s = (Solver*)h;
    if (auto* v = s->field(f)) {
        if (count != v->size()) return IFL_ERR_INVALID;
        std::memcpy(v->data(), host, count * sizeof(double));
        return IFL_OK;
    }
    if (auto* v = s->int_field(f)) {
        if (count != v->size()) return IFL_ERR_INVALID;
        for (size_t i = 0; i < count; i++) (*v)[i] = (int)host[i];
        return IFL_OK;
    }
    return IFL_ERR_INVALID;
}

/* scalar helpers exported so the tests can pin the interpolation invariants */
double ifo_lerp(double a, double b, double x) { return lerp(a, b, x); }
double ifo_cerp(double a, double b, double c, double d, double x) { return cerp(a, b, c, d, x); }
double ifo_cubic_pulse(double x) { return cubicPulse(x); }
/* the standalone reductions, for kernel-level comparisons */
double ifo_dot(void* h, int fa, int fb) { Solver* s = (Solver*)h; return s->dotProduct(*s->field(fa), *s->field(fb)); }
double ifo_inf_norm(void* h, int fa) { Solver* s = (Solver*)h; return s->infinityNorm(*s->field(fa)); }

} // extern "C"
