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
 *         which this restatement must reproduce.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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

enum { CELL_FLUID = 0, CELL_SOLID = 1, CELL_EMPTY = 2 };

struct Vec2 { double x, y; };

/* ---- FluidQuantity (F1:105, F3:114, F4:126, F5:234, F7:225) ----------------------------- */
struct Quantity {
    std::vector<double> src, dst;
    int w = 0, h = 0;
    double ox = 0, oy = 0, hx = 0;

    void init(int w_, int h_, double ox_, double oy_, double hx_) {
        w = w_; h = h_; ox = ox_; oy = oy_; hx = hx_;
        src.assign((size_t)w * h, 0.0);
        dst.assign((size_t)w * h, 0.0);
    }
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
    ifl_step_status st;

    explicit Solver(const ifl_config& c) : cfg(c) {
        variant = c.variant; w = c.w; h = c.h; density = c.density;
        hx = 1.0 / jmini(w, h);                                       /* F1:266 */
        d.init(w, h, 0.5, 0.5, hx);                                   /* F1:268 */
        u.init(w + 1, h, 0.0, 0.5, hx);                               /* F1:269 */
        v.init(w, h + 1, 0.5, 0.0, hx);                               /* F1:270 */
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

    void project(double timestep) {
        if (usesPcg()) projectPCG(cfg.solver_limit, st.iterations_pressure, st.residual_pressure, 1);
        else           projectGS(cfg.solver_limit, timestep);
    }

    /* F1:298 / F3:385 */
    void addInflow(double x, double y, double ww, double hh, double dv, double /*tv*/, double uv, double vv) {
        d.addInflow(variant, x, y, x + ww, y + hh, dv);
        u.addInflow(variant, x, y, x + ww, y + hh, uv);
        v.addInflow(variant, x, y, x + ww, y + hh, vv);
    }

    /* F1:276 / F2:322 / F3:361 */
    void update(double timestep) {
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

    std::vector<double>* field(int f) {
        switch (f) {
            case IFL_FIELD_D: return &d.src;   case IFL_FIELD_U: return &u.src;   case IFL_FIELD_V: return &v.src;
            case IFL_FIELD_D_DST: return &d.dst; case IFL_FIELD_U_DST: return &u.dst; case IFL_FIELD_V_DST: return &v.dst;
            case IFL_FIELD_P: return &p; case IFL_FIELD_R: return &r; case IFL_FIELD_Z: return &z; case IFL_FIELD_S: return &s;
            case IFL_FIELD_PRECON: return &precon; case IFL_FIELD_ADIAG: return &aDiag;
            case IFL_FIELD_APLUSX: return &aPlusX; case IFL_FIELD_APLUSY: return &aPlusY;
            default: return nullptr;
        }
    }
};

} // namespace

extern "C" {

int ifo_create(const ifl_config* cfg, void** out) {
    if (!cfg || !out || cfg->w < 2 || cfg->h < 2) return IFL_ERR_INVALID;
    if (cfg->variant < IFL_F1_MATRIXLESS || cfg->variant > IFL_F3_CONJUGATE_GRADIENTS) return IFL_ERR_UNSUPPORTED;
    *out = new Solver(*cfg);
    return IFL_OK;
}
void ifo_destroy(void* h) { delete (Solver*)h; }

int ifo_add_inflow(void* h, double x, double y, double w, double hh, double d, double t, double u, double v) {
    ((Solver*)h)->addInflow(x, y, w, hh, d, t, u, v);
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
    switch (stage) {
        case IFL_STAGE_BUILD_RHS: s->buildRhs(); break;
        case IFL_STAGE_BUILD_MATRIX: if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED; s->buildPressureMatrix(dt); break;
        case IFL_STAGE_BUILD_PRECON: if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED; s->buildPreconditioner(); break;
        case IFL_STAGE_PROJECT: s->project(dt); break;
        case IFL_STAGE_APPLY_PRESSURE: s->applyPressure(dt); break;
        case IFL_STAGE_ADVECT_D: s->d.advect(s->variant, dt, s->u, s->v); break;
        case IFL_STAGE_ADVECT_U: s->u.advect(s->variant, dt, s->u, s->v); break;
        case IFL_STAGE_ADVECT_V: s->v.advect(s->variant, dt, s->u, s->v); break;
        case IFL_STAGE_SWAP: s->d.swap(); s->u.swap(); s->v.swap(); break;
        case IFL_STAGE_APPLY_PRECON: if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED; s->applyPreconditioner(s->z, s->r); break;
        case IFL_STAGE_MATVEC: if (!s->usesPcg()) return IFL_ERR_UNSUPPORTED; s->matrixVectorProduct(s->z, s->s); break;
        default: return IFL_ERR_UNSUPPORTED;
    }
    if (st) *st = s->st;
    return IFL_OK;
}
int ifo_max_timestep(void* h, double* out) { *out = ((Solver*)h)->maxTimestep(); return IFL_OK; }

int ifo_field_size(void* h, int f, size_t* count) {
    auto* v = ((Solver*)h)->field(f);
    if (!v) return IFL_ERR_INVALID;
    *count = v->size();
    return IFL_OK;
}
int ifo_read_field(void* h, int f, double* host, size_t count) {
    auto* v = ((Solver*)h)->field(f);
    if (!v || count != v->size()) return IFL_ERR_INVALID;
    std::memcpy(host, v->data(), count * sizeof(double));
    return IFL_OK;
}
int ifo_write_field(void* h, int f, const double* host, size_t count) {
    auto* v = ((Solver*)h)->field(f);
    if (!v || count != v->size()) return IFL_ERR_INVALID;
    std::memcpy(v->data(), host, count * sizeof(double));
    return IFL_OK;
}

/* scalar helpers exported so the tests can pin the interpolation invariants */
double ifo_lerp(double a, double b, double x) { return lerp(a, b, x); }
double ifo_cerp(double a, double b, double c, double d, double x) { return cerp(a, b, c, d, x); }
double ifo_cubic_pulse(double x) { return cubicPulse(x); }
/* the standalone reductions, for kernel-level comparisons */
double ifo_dot(void* h, int fa, int fb) { Solver* s = (Solver*)h; return s->dotProduct(*s->field(fa), *s->field(fb)); }
double ifo_inf_norm(void* h, int fa) { Solver* s = (Solver*)h; return s->infinityNorm(*s->field(fa)); }

} // extern "C"
