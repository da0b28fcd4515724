"""A second, independent restatement of ExplicitFluid1matrixless.java -- literal, line by line, in plain Python.

Test infrastructure (like oracle/): it exists to pin the C++ oracle.  The reference cannot run here (no JVM), so the
oracle is cross-checked against a transcription made directly from the Java text in a different language and style
(Python floats are IEEE doubles, `*`/`+`/`-`/`/` round once, nothing is fused).  Only tiny grids: pure-Python loops.

Java semantics spelled out: `(int) x` truncates toward zero (int()), `(float) x` rounds to binary32 and widens back
(struct 'f'), `Math.max/min/abs` on doubles, `new double[n]` is zero-filled.
"""
import struct


def f32(x):
    """(float) x widened back to double"""
    return struct.unpack("f", struct.pack("f", x))[0]


def jmax(a, b):
    """Math.max(double, double): NaN propagates, max(-0.0, +0.0) = +0.0"""
    if a != a or b != b:
        return float("nan")
    if a == 0.0 and b == 0.0:
        return b if struct.pack(">d", a)[0] & 0x80 else a
    return a if a >= b else b


def jmin(a, b):
    """Math.min(double, double): NaN propagates, min(-0.0, +0.0) = -0.0"""
    if a != a or b != b:
        return float("nan")
    if a == 0.0 and b == 0.0:
        return a if struct.pack(">d", a)[0] & 0x80 else b
    return a if a <= b else b


def lerp(a, b, x):                                  # math/Interpolation.java:77-86
    return (1 - x) * a + x * b


class FluidQuantity:                                # ExplicitFluid1matrixless.java:105
    def __init__(self, w, h, ox, oy, hx):           # F1:131-141
        self._w, self._h, self._ox, self._oy, self._hx = w, h, ox, oy, hx
        self._src = [0.0] * (w * h)
        self._dst = [0.0] * (w * h)

    def swap(self):                                 # F1:143
        self._src, self._dst = self._dst, self._src

    def getSrc(self, x, y):                         # F1:149
        return self._src[x + y * self._w]

    def setSrc(self, x, y, value):                  # F1:153
        self._src[x + y * self._w] = value

    def lerpGrid(self, x, y):                       # F1:161
        x = jmin(jmax(x - self._ox, 0.0), self._w - 1.001)
        y = jmin(jmax(y - self._oy, 0.0), self._h - 1.001)
        ix = int(x)
        iy = int(y)
        x -= ix
        y -= iy
        x00, x10 = self.getSrc(ix + 0, iy + 0), self.getSrc(ix + 1, iy + 0)
        x01, x11 = self.getSrc(ix + 0, iy + 1), self.getSrc(ix + 1, iy + 1)
        return lerp(lerp(x00, x10, x), lerp(x01, x11, x), y)

    def advect(self, timestep, u, v):               # F1:178
        idx = 0
        for iy in range(self._h):
            for ix in range(self._w):
                x = ix + self._ox
                y = iy + self._oy
                x, y = self.Euler(x, y, timestep, u, v)
                self._dst[idx] = self.lerpGrid(x, y)
                idx += 1

    def addInflow(self, x0, y0, x1, y1, v):         # F1:202
        ix0 = int(x0 / self._hx - self._ox)
        iy0 = int(y0 / self._hx - self._oy)
        ix1 = int(x1 / self._hx - self._ox)
        iy1 = int(y1 / self._hx - self._oy)
        for y in range(max(iy0, 0), min(iy1, self._h)):
            for x in range(max(ix0, 0), min(ix1, self._h)):          # sic: _h, F1:209
                if abs(f32(self._src[x + y * self._w])) < abs(f32(v)):
                    self._src[x + y * self._w] = v

    def Euler(self, x, y, timestep, u, v):          # F1:220
        uVel = u.lerpGrid(x, y) / self._hx
        vVel = v.lerpGrid(x, y) / self._hx
        x -= uVel * timestep
        y -= vVel * timestep
        return x, y


class FluidSolver:                                  # F1:240
    def __init__(self, w, h, density):              # F1:262
        self._w, self._h, self._density = w, h, density
        self._hx = 1.0 / min(w, h)
        self._d = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)
        self._r = [0.0] * (w * h)
        self._p = [0.0] * (w * h)
        self.lastIterations = None                  # what the reference only logs (F1:420 / F1:425)
        self.lastMaxDelta = None

    def update(self, timestep):                     # F1:276
        self.buildRhs()
        self.project(600, timestep)
        self.applyPressure(timestep)
        self._d.advect(timestep, self._u, self._v)
        self._u.advect(timestep, self._u, self._v)
        self._v.advect(timestep, self._u, self._v)
        self._d.swap()
        self._u.swap()
        self._v.swap()

    def addInflow(self, x, y, w, h, d, u, v):       # F1:298
        self._d.addInflow(x, y, x + w, y + h, d)
        self._u.addInflow(x, y, x + w, y + h, u)
        self._v.addInflow(x, y, x + w, y + h, v)

    def buildRhs(self):                             # F1:364
        scale = 1.0 / self._hx
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                self._r[idx] = -scale * (self._u.getSrc(x + 1, y) - self._u.getSrc(x, y)
                                         + self._v.getSrc(x, y + 1) - self._v.getSrc(x, y))
                idx += 1

    def project(self, limit, timestep):             # F1:380
        scale = timestep / (self._density * self._hx * self._hx)
        _w, _h, _p, _r = self._w, self._h, self._p, self._r
        maxDelta = 0.0
        for it in range(limit):
            maxDelta = 0.0
            idx = 0
            for y in range(_h):
                for x in range(_w):
                    diag, offDiag = 0.0, 0.0
                    if x > 0:
                        diag += scale
                        offDiag -= scale * _p[idx - 1]
                    if y > 0:
                        diag += scale
                        offDiag -= scale * _p[idx - _w]
                    if x < _w - 1:
                        diag += scale
                        offDiag -= scale * _p[idx + 1]
                    if y < _h - 1:
                        diag += scale
                        offDiag -= scale * _p[idx + _w]
                    newP = (_r[idx] - offDiag) / diag
                    maxDelta = jmax(maxDelta, abs(f32(_p[idx]) - newP))      # (float) binds to _p[idx] only, F1:414
                    _p[idx] = newP
                    idx += 1
            if maxDelta < 1e-5:
                self.lastIterations, self.lastMaxDelta = it, maxDelta
                return
        self.lastIterations, self.lastMaxDelta = limit, maxDelta

    def applyPressure(self, timestep):              # F1:432
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
