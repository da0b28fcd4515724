"""Literal Python transcription of ExplicitFluid4solidBoundaries.java (solid bodies, extrapolation into solids,
back-projected advection, masked MIC(0)-PCG), made from the Java text -- see java_f1_transcript.py.  lerpGrid, cerpGrid,
addInflow and RungeKutta3 of its FluidQuantity are textually those of ExplicitFluid3 (checked mechanically against the
reference) and are inherited from that transcription.  Quirks kept as written: setBoundaryCondition calls velocityX for
the v faces too (F4:956-958); advect leaves stale _dst in non-fluid cells (F4:272)."""
import math

import java_f3_transcript as f3
import java_solid_transcript as js
from java_f1_transcript import f32, jmax

CELL_FLUID, CELL_SOLID = js.CELL_FLUID, js.CELL_SOLID


def isignum(v):
    """(int) signum(v)"""
    return int(js.signum(v))


def closestSurfacePoint(body, x, y):
    if isinstance(body, js.SolidBox):               # BOX:95
        x -= body._posX
        y -= body._posY
        x, y = body.rotate(x, y, body._c, -body._s)
        dx = abs(x) - body._scaleX * 0.5
        dy = abs(y) - body._scaleY * 0.5
        if dx > dy:
            x = body.nsgn(x) * 0.5 * body._scaleX
        else:
            y = body.nsgn(y) * 0.5 * body._scaleY
        x, y = body.rotate(x, y, body._c, body._s)
        return x + body._posX, y + body._posY
    # SPH:67 with globalToLocal SB:121 / localToGlobal SB:138
    x -= body._posX
    y -= body._posY
    x, y = body.rotate(x, y, body._c, -body._s)
    x, y = x / body._scaleX, y / body._scaleY
    r = js.length(x, y)
    if r < 1e-4:
        x, y = 0.5, 0.0
    else:
        x /= 2.0 * r
        y /= 2.0 * r
    x *= body._scaleX
    y *= body._scaleY
    x, y = body.rotate(x, y, body._c, body._s)
    return x + body._posX, y + body._posY


class FluidQuantity(f3.FluidQuantity):              # F4:128
    def __init__(self, w, h, ox, oy, hx):           # F4:176
        super().__init__(w, h, ox, oy, hx)
        n = w * h
        self._normalX, self._normalY = [0.0] * n, [0.0] * n
        self._cell, self._body, self._mask = [None] * n, [0] * n, [0] * n

    def backProject(self, x, y, bodies):            # F4:249
        rx = min(max(int(x - self._ox), 0), self._w - 1)
        ry = min(max(int(y - self._oy), 0), self._h - 1)
        if self._cell[rx + ry * self._w] != CELL_FLUID:
            x = (x - self._ox) * self._hx
            y = (y - self._oy) * self._hx
            x, y = closestSurfacePoint(bodies[self._body[rx + ry * self._w]], x, y)
            x = x / self._hx + self._ox
            y = y / self._hx + self._oy
        return x, y

    def advect(self, timestep, u, v, bodies):       # F4:268
        idx = 0
        for iy in range(self._h):
            for ix in range(self._w):
                if self._cell[idx] == CELL_FLUID:
                    x = ix + self._ox
                    y = iy + self._oy
                    x, y = self.RungeKutta3(x, y, timestep, u, v)
                    x, y = self.backProject(x, y, bodies)
                    self._dst[idx] = self.cerpGrid(x, y)
                idx += 1

    def fillSolidFields(self, bodies):              # F4:329
        cell, body = js.fillSolidFields(self._w, self._h, self._ox, self._oy, self._hx, bodies, self._normalX, self._normalY)
        if bodies:
            self._cell, self._body = cell, body

    def fillSolidMask(self):                        # F4:396
        _w, _h = self._w, self._h
        for y in range(1, _h - 1):
            for x in range(1, _w - 1):
                idx = x + y * _w
                if self._cell[idx] == CELL_FLUID:
                    continue
                nx = self._normalX[idx]
                ny = self._normalY[idx]
                self._mask[idx] = 0
                if nx != 0.0 and self._cell[idx + isignum(nx)] != CELL_FLUID:
                    self._mask[idx] |= 1
                if ny != 0.0 and self._cell[idx + isignum(ny) * _w] != CELL_FLUID:
                    self._mask[idx] |= 2

    def extrapolateNormal(self, idx):               # F4:426
        nx = self._normalX[idx]
        ny = self._normalY[idx]
        srcX = self._src[idx + isignum(nx)]
        srcY = self._src[idx + isignum(ny) * self._w]
        return (abs(f32(nx)) * srcX + abs(f32(ny)) * srcY) / (abs(f32(nx)) + abs(f32(ny)))

    def freeNeighbour(self, idx, border, mask):     # F4:441
        self._mask[idx] &= ~mask
        if self._cell[idx] != CELL_FLUID and self._mask[idx] == 0:
            border.append(idx)
        return border

    def extrapolate(self):                          # F4:452
        self.fillSolidMask()
        _w, _h = self._w, self._h
        border = []
        for y in range(1, _h - 1):
            for x in range(1, _w - 1):
                idx = x + y * _w
                if self._cell[idx] != CELL_FLUID and self._mask[idx] == 0:
                    border.append(idx)
        while border:
            idx = border.pop()
            self._src[idx] = self.extrapolateNormal(idx)
            if self._normalX[idx - 1] > 0.0:
                border = self.freeNeighbour(idx - 1, border, 1)
            if self._normalX[idx + 1] < 0.0:
                border = self.freeNeighbour(idx + 1, border, 1)
            if self._normalY[idx - _w] > 0.0:
                border = self.freeNeighbour(idx - _w, border, 2)
            if self._normalY[idx + _w] < 0.0:
                border = self.freeNeighbour(idx + _w, border, 2)


class FluidSolver(f3.FluidSolver):                  # F4:560
    def __init__(self, w, h, density, bodies):      # F4:594
        super().__init__(w, h, density)
        self._bodies = bodies
        self._d = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)

    def update(self, timestep):                     # F4:617
        self._d.fillSolidFields(self._bodies)
        self._u.fillSolidFields(self._bodies)
        self._v.fillSolidFields(self._bodies)
        self.setBoundaryCondition()
        self.buildRhs()
        self.buildPressureMatrix(timestep)
        self.buildPreconditioner()
        self.project(2000)
        self.applyPressure(timestep)
        self._d.extrapolate()
        self._u.extrapolate()
        self._v.extrapolate()
        self.setBoundaryCondition()
        self._d.advect(timestep, self._u, self._v, self._bodies)
        self._u.advect(timestep, self._u, self._v, self._bodies)
        self._v.advect(timestep, self._u, self._v, self._bodies)
        self._d.swap()
        self._u.swap()
        self._v.swap()

    def buildRhs(self):                             # F4:693
        scale = 1.0 / self._hx
        cell = self._d._cell
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                if cell[idx] == CELL_FLUID:
                    self._r[idx] = -scale * (self._u.getSrc(x + 1, y) - self._u.getSrc(x, y)
                                             + self._v.getSrc(x, y + 1) - self._v.getSrc(x, y))
                else:
                    self._r[idx] = 0.0
                idx += 1

    def buildPressureMatrix(self, timestep):        # F4:713
        scale = timestep / (self._density * self._hx * self._hx)
        cell = self._d._cell
        _w, _h = self._w, self._h
        n = _w * _h
        self._aDiag, self._aPlusX, self._aPlusY = [0.0] * n, [0.0] * n, [0.0] * n
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if cell[idx] == CELL_FLUID:
                    if x < _w - 1 and cell[idx + 1] == CELL_FLUID:
                        self._aDiag[idx] += scale
                        self._aDiag[idx + 1] += scale
                        self._aPlusX[idx] = -scale
                    if y < _h - 1 and cell[idx + _w] == CELL_FLUID:
                        self._aDiag[idx] += scale
                        self._aDiag[idx + _w] += scale
                        self._aPlusY[idx] = -scale
                idx += 1

    def buildPreconditioner(self):                  # F4:744
        tau, sigma = 0.97, 0.25
        cell = self._d._cell
        _w, _h = self._w, self._h
        aD, aX, aY, pre = self._aDiag, self._aPlusX, self._aPlusY, self._precon
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if cell[idx] == CELL_FLUID:
                    e = aD[idx]
                    if x > 0 and cell[idx - 1] == CELL_FLUID:
                        px = aX[idx - 1] * pre[idx - 1]
                        py = aY[idx - 1] * pre[idx - 1]
                        e -= (px * px + tau * px * py)
                    if y > 0 and cell[idx - _w] == CELL_FLUID:
                        px = aX[idx - _w] * pre[idx - _w]
                        py = aY[idx - _w] * pre[idx - _w]
                        e -= (py * py + tau * px * py)
                    if e < sigma * aD[idx]:
                        e = aD[idx]
                    pre[idx] = 1.0 / math.sqrt(e)
                idx += 1

    def applyPreconditioner(self, dst, a):          # F4:780
        cell = self._d._cell
        _w, _h = self._w, self._h
        aX, aY, pre = self._aPlusX, self._aPlusY, self._precon
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if cell[idx] == CELL_FLUID:
                    t = a[idx]
                    if x > 0 and cell[idx - 1] == CELL_FLUID:
                        t -= aX[idx - 1] * pre[idx - 1] * dst[idx - 1]
                    if y > 0 and cell[idx - _w] == CELL_FLUID:
                        t -= aY[idx - _w] * pre[idx - _w] * dst[idx - _w]
                    dst[idx] = t * pre[idx]
                idx += 1
        idx = _w * _h - 1
        for y in range(_h - 1, -1, -1):
            for x in range(_w - 1, -1, -1):
                if cell[idx] == CELL_FLUID:
                    t = dst[idx]
                    if x < _w - 1 and cell[idx + 1] == CELL_FLUID:
                        t -= aX[idx] * pre[idx] * dst[idx + 1]
                    if y < _h - 1 and cell[idx + _w] == CELL_FLUID:
                        t -= aY[idx] * pre[idx] * dst[idx + _w]
                    dst[idx] = t * pre[idx]
                idx -= 1

    # dotProduct F4:825, matrixVectorProduct F4:837, scaledAdd F4:863, infinityNorm F4:872, project F4:883: as in F3

    def applyPressure(self, timestep):              # F4:927 (no zeroing of the domain edges here)
        scale = timestep / (self._density * self._hx)
        cell = self._d._cell
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                if cell[idx] == CELL_FLUID:
                    self._u.setSrc(x, y, self._u.getSrc(x, y) - scale * self._p[idx])
                    self._u.setSrc(x + 1, y, self._u.getSrc(x + 1, y) + scale * self._p[idx])
                    self._v.setSrc(x, y, self._v.getSrc(x, y) - scale * self._p[idx])
                    self._v.setSrc(x, y + 1, self._v.getSrc(x, y + 1) + scale * self._p[idx])
                idx += 1

    def setBoundaryCondition(self):                 # F4:948
        cell, body = self._d._cell, self._d._body
        _hx = self._hx
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                if cell[idx] == CELL_SOLID:
                    b = self._bodies[body[idx]]
                    self._u.setSrc(x, y, b.velocityX(x * _hx, (y + 0.5) * _hx))
                    self._v.setSrc(x, y, b.velocityX((x + 0.5) * _hx, y * _hx))                 # sic: velocityX
                    self._u.setSrc(x + 1, y, b.velocityX((x + 1.0) * _hx, (y + 0.5) * _hx))
                    self._v.setSrc(x, y + 1, b.velocityX((x + 0.5) * _hx, (y + 1.0) * _hx))     # sic
                idx += 1
        for y in range(self._h):
            self._u.setSrc(0, y, 0.0)
            self._u.setSrc(self._w, y, 0.0)
        for x in range(self._w):
            self._v.setSrc(x, 0, 0.0)
            self._v.setSrc(x, self._h, 0.0)
