"""Literal Python transcription of ExplicitFluid3conjugateGradients.java (RK3 + Catmull-Rom advection, MIC(0)-PCG
projection), made from the Java text and not from the oracle -- see java_f1_transcript.py for the why and the rules.
Port quirks are kept as written: `min(ix1, _h)` in addInflow (F3:238), `lastU`/`lastV` NOT divided by _hx (F3:271-272)."""
import math

from java_f1_transcript import f32, jmax, jmin, lerp


def length(x, y):                                   # F3:97
    return math.sqrt(x * x + y * y)


def cubicPulse(x):                                  # F3:109
    x = jmin(abs(f32(x)), 1.0)
    return 1.0 - x * x * (3.0 - 2.0 * x)


def cerp(a, b, c, d, x):                            # math/Interpolation.java:135-152
    xsq = x * x
    xcu = xsq * x
    minV = jmin(a, jmin(b, jmin(c, d)))
    maxV = jmax(a, jmax(b, jmax(c, d)))
    t = (a * (0.0 - 0.5 * x + 1.0 * xsq - 0.5 * xcu)
         + b * (1.0 + 0.0 * x - 2.5 * xsq + 1.5 * xcu)
         + c * (0.0 + 0.5 * x + 2.0 * xsq - 1.5 * xcu)
         + d * (0.0 + 0.0 * x - 0.5 * xsq + 0.5 * xcu))
    return jmin(jmax(t, minV), maxV)


class FluidQuantity:                                # F3:114
    def __init__(self, w, h, ox, oy, hx):
        self._w, self._h, self._ox, self._oy, self._hx = w, h, ox, oy, hx
        self._src = [0.0] * (w * h)
        self._dst = [0.0] * (w * h)

    def swap(self):
        self._src, self._dst = self._dst, self._src

    def getSrc(self, x, y):
        return self._src[x + y * self._w]

    def setSrc(self, x, y, value):
        self._src[x + y * self._w] = value

    def lerpGrid(self, x, y):                       # F3:167
        x = jmin(jmax(x - self._ox, 0.0), self._w - 1.001)
        y = jmin(jmax(y - self._oy, 0.0), self._h - 1.001)
        ix = int(x)
        iy = int(y)
        x -= ix
        y -= iy
        x00, x10 = self.getSrc(ix + 0, iy + 0), self.getSrc(ix + 1, iy + 0)
        x01, x11 = self.getSrc(ix + 0, iy + 1), self.getSrc(ix + 1, iy + 1)
        return lerp(lerp(x00, x10, x), lerp(x01, x11, x), y)

    def cerpGrid(self, x, y):                       # F3:185
        x = jmin(jmax(x - self._ox, 0.0), self._w - 1.001)
        y = jmin(jmax(y - self._oy, 0.0), self._h - 1.001)
        ix = int(x)
        iy = int(y)
        x -= ix
        y -= iy
        x0, x1, x2, x3 = max(ix - 1, 0), ix, ix + 1, min(ix + 2, self._w - 1)
        y0, y1, y2, y3 = max(iy - 1, 0), iy, iy + 1, min(iy + 2, self._h - 1)
        g = self.getSrc
        q0 = cerp(g(x0, y0), g(x1, y0), g(x2, y0), g(x3, y0), x)
        q1 = cerp(g(x0, y1), g(x1, y1), g(x2, y1), g(x3, y1), x)
        q2 = cerp(g(x0, y2), g(x1, y2), g(x2, y2), g(x3, y2), x)
        q3 = cerp(g(x0, y3), g(x1, y3), g(x2, y3), g(x3, y3), x)
        return cerp(q0, q1, q2, q3, y)

    def advect(self, timestep, u, v):               # F3:207
        idx = 0
        for iy in range(self._h):
            for ix in range(self._w):
                x = ix + self._ox
                y = iy + self._oy
                x, y = self.RungeKutta3(x, y, timestep, u, v)
                self._dst[idx] = self.cerpGrid(x, y)
                idx += 1

    def addInflow(self, x0, y0, x1, y1, v):         # F3:230
        ix0 = int(x0 / self._hx - self._ox)
        iy0 = int(y0 / self._hx - self._oy)
        ix1 = int(x1 / self._hx - self._ox)
        iy1 = int(y1 / self._hx - self._oy)
        for y in range(max(iy0, 0), min(iy1, self._h)):
            for x in range(max(ix0, 0), min(ix1, self._h)):          # sic: _h
                l = length((2.0 * (x + 0.5) * self._hx - (x0 + x1)) / (x1 - x0),
                           (2.0 * (y + 0.5) * self._hx - (y0 + y1)) / (y1 - y0))
                vi = cubicPulse(l) * v
                if abs(f32(self._src[x + y * self._w])) < abs(f32(vi)):
                    self._src[x + y * self._w] = vi

    def RungeKutta3(self, x, y, timestep, u, v):    # F3:255
        firstU = u.lerpGrid(x, y) / self._hx
        firstV = v.lerpGrid(x, y) / self._hx
        midX = x - 0.5 * timestep * firstU
        midY = y - 0.5 * timestep * firstV
        midU = u.lerpGrid(midX, midY) / self._hx
        midV = v.lerpGrid(midX, midY) / self._hx
        lastX = x - 0.75 * timestep * midU
        lastY = y - 0.75 * timestep * midV
        lastU = u.lerpGrid(lastX, lastY)             # sic: no division by _hx
        lastV = v.lerpGrid(lastX, lastY)
        x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU)
        y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV)
        return x, y


class FluidSolver:                                  # F3:280
    def __init__(self, w, h, density):              # F3:339
        self._w, self._h, self._density = w, h, density
        self._hx = 1.0 / min(w, h)
        self._d = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)
        n = w * h
        self._r, self._p, self._z, self._s = [0.0] * n, [0.0] * n, [0.0] * n, [0.0] * n
        self._aDiag, self._aPlusX, self._aPlusY, self._precon = [0.0] * n, [0.0] * n, [0.0] * n, [0.0] * n
        self.lastIterations = None                  # the `iter` the reference logs (F3:613 / F3:628), None = early return
        self.lastMaxError = None

    def update(self, timestep):                     # F3:361
        self.buildRhs()
        self.buildPressureMatrix(timestep)
        self.buildPreconditioner()
        self.project(600)
        self.applyPressure(timestep)
        self._d.advect(timestep, self._u, self._v)
        self._u.advect(timestep, self._u, self._v)
        self._v.advect(timestep, self._u, self._v)
        self._d.swap()
        self._u.swap()
        self._v.swap()

    def addInflow(self, x, y, w, h, d, u, v):       # F3:385
        self._d.addInflow(x, y, x + w, y + h, d)
        self._u.addInflow(x, y, x + w, y + h, u)
        self._v.addInflow(x, y, x + w, y + h, v)

    def buildRhs(self):                             # F3:420
        scale = 1.0 / self._hx
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                self._r[idx] = -scale * (self._u.getSrc(x + 1, y) - self._u.getSrc(x, y)
                                         + self._v.getSrc(x, y + 1) - self._v.getSrc(x, y))
                idx += 1

    def buildPressureMatrix(self, timestep):        # F3:435
        scale = timestep / (self._density * self._hx * self._hx)
        _w, _h = self._w, self._h
        self._aDiag = [0.0] * (_w * _h)
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if x < _w - 1:
                    self._aDiag[idx] += scale
                    self._aDiag[idx + 1] += scale
                    self._aPlusX[idx] = -scale
                else:
                    self._aPlusX[idx] = 0.0
                if y < _h - 1:
                    self._aDiag[idx] += scale
                    self._aDiag[idx + _w] += scale
                    self._aPlusY[idx] = -scale
                else:
                    self._aPlusY[idx] = 0.0
                idx += 1

    def buildPreconditioner(self):                  # F3:464
        tau = 0.97
        sigma = 0.25
        _w, _h = self._w, self._h
        aD, aX, aY, pre = self._aDiag, self._aPlusX, self._aPlusY, self._precon
        idx = 0
        for y in range(_h):
            for x in range(_w):
                e = aD[idx]
                if x > 0:
                    px = aX[idx - 1] * pre[idx - 1]
                    py = aY[idx - 1] * pre[idx - 1]
                    e -= (px * px + tau * px * py)
                if y > 0:
                    px = aX[idx - _w] * pre[idx - _w]
                    py = aY[idx - _w] * pre[idx - _w]
                    e -= (py * py + tau * px * py)
                if e < sigma * aD[idx]:
                    e = aD[idx]
                pre[idx] = 1.0 / math.sqrt(e)
                idx += 1

    def applyPreconditioner(self, dst, a):          # F3:495
        _w, _h = self._w, self._h
        aX, aY, pre = self._aPlusX, self._aPlusY, self._precon
        idx = 0
        for y in range(_h):
            for x in range(_w):
                t = a[idx]
                if x > 0:
                    t -= aX[idx - 1] * pre[idx - 1] * dst[idx - 1]
                if y > 0:
                    t -= aY[idx - _w] * pre[idx - _w] * dst[idx - _w]
                dst[idx] = t * pre[idx]
                idx += 1
        idx = _w * _h - 1
        for y in range(_h - 1, -1, -1):
            for x in range(_w - 1, -1, -1):
                t = dst[idx]
                if x < _w - 1:
                    t -= aX[idx] * pre[idx] * dst[idx + 1]
                if y < _h - 1:
                    t -= aY[idx] * pre[idx] * dst[idx + _w]
                dst[idx] = t * pre[idx]
                idx -= 1

    def dotProduct(self, a, b):                     # F3:532
        result = 0.0
        for i in range(self._w * self._h):
            result += a[i] * b[i]
        return result

    def matrixVectorProduct(self, dst, b):          # F3:544
        _w, _h = self._w, self._h
        aD, aX, aY = self._aDiag, self._aPlusX, self._aPlusY
        idx = 0
        for y in range(_h):
            for x in range(_w):
                t = aD[idx] * b[idx]
                if x > 0:
                    t += aX[idx - 1] * b[idx - 1]
                if y > 0:
                    t += aY[idx - _w] * b[idx - _w]
                if x < _w - 1:
                    t += aX[idx] * b[idx + 1]
                if y < _h - 1:
                    t += aY[idx] * b[idx + _w]
                dst[idx] = t
                idx += 1

    def scaledAdd(self, dst, a, b, s):              # F3:570
        for i in range(self._w * self._h):
            dst[i] = a[i] + b[i] * s

    def infinityNorm(self, a):                      # F3:579
        maxA = 0.0
        for i in range(self._w * self._h):
            maxA = jmax(maxA, abs(f32(a[i])))
        return maxA

    def project(self, limit):                       # F3:590
        self._p = [0.0] * (self._w * self._h)
        self.applyPreconditioner(self._z, self._r)
        self._s[:] = self._z
        maxError = self.infinityNorm(self._r)
        self.lastIterations, self.lastMaxError = None, maxError
        if maxError < 1e-5:
            return
        sigma = self.dotProduct(self._z, self._r)
        for it in range(limit):
            self.matrixVectorProduct(self._z, self._s)
            alpha = sigma / self.dotProduct(self._z, self._s)
            self.scaledAdd(self._p, self._p, self._s, alpha)
            self.scaledAdd(self._r, self._r, self._z, -alpha)
            maxError = self.infinityNorm(self._r)
            self.lastIterations, self.lastMaxError = it, maxError
            if maxError < 1e-5:
                return
            if maxError != maxError:
                raise FloatingPointError("System.exit(-1), F3:618")
            self.applyPreconditioner(self._z, self._r)
            sigmaNew = self.dotProduct(self._z, self._r)
            self.scaledAdd(self._s, self._z, self._s, sigmaNew / sigma)
            sigma = sigmaNew
        self.lastIterations = limit

    def applyPressure(self, timestep):              # F3:634
        scale = timestep / (self._density * self._hx)
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                self._u.setSrc(x, y, self._u.getSrc(x, y) - scale * self._p[idx])
                self._u.setSrc(x + 1, y, self._u.getSrc(x + 1, y) + scale * self._p[idx])
                self._v.setSrc(x, y, self._v.getSrc(x, y) - scale * self._p[idx])
                self._v.setSrc(x, y + 1, self._v.getSrc(x, y + 1) + scale * self._p[idx])
                idx += 1
        for y in range(self._h):
            self._u.setSrc(0, y, 0.0)
            self._u.setSrc(self._w, y, 0.0)
        for x in range(self._w):
            self._v.setSrc(x, 0, 0.0)
            self._v.setSrc(x, self._h, 0.0)
