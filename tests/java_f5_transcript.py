"""Literal Python transcription of what ExplicitFluid5curvedBoundaries.java changes with respect to ExplicitFluid4
(the rest of the two files is textually identical, checked mechanically): the level set and marching-squares volumes in
fillSolidFields F5:463, the volume-weighted right-hand side with the body velocities F5:785 and the volume-weighted
matrix F5:825.  See java_f1_transcript.py."""
import java_f4_transcript as f4
import java_solid_transcript as js

CELL_FLUID = f4.CELL_FLUID


class FluidQuantity(f4.FluidQuantity):              # F5:235
    def __init__(self, w, h, ox, oy, hx):           # F5:298
        super().__init__(w, h, ox, oy, hx)
        self._phi = [0.0] * ((w + 1) * (h + 1))
        self._volume = [1.0] * (w * h)
        self._cell = [CELL_FLUID] * (w * h)

    def volume(self, x, y):                         # F5:349
        return self._volume[x + y * self._w]

    def fillSolidFields(self, bodies):              # F5:463
        if not bodies:
            return
        self._phi, self._volume, self._cell, self._body = js.fillSolidFields5(
            self._w, self._h, self._ox, self._oy, self._hx, bodies, self._normalX, self._normalY)


class FluidSolver(f4.FluidSolver):
    def __init__(self, w, h, density, bodies):
        super().__init__(w, h, density, bodies)
        self._d = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)

    def buildRhs(self):                             # F5:785
        scale = 1.0 / self._hx
        cell, body = self._d._cell, self._d._body
        _u, _v, _d, _hx, _w, _h, _bodies = self._u, self._v, self._d, self._hx, self._w, self._h, self._bodies
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if cell[idx] == CELL_FLUID:
                    self._r[idx] = -scale \
                        * (_u.volume(x + 1, y) * _u.getSrc(x + 1, y) - _u.volume(x, y) * _u.getSrc(x, y)
                           + _v.volume(x, y + 1) * _v.getSrc(x, y + 1) - _v.volume(x, y) * _v.getSrc(x, y))
                    vol = _d.volume(x, y)
                    if _bodies:
                        if x > 0:
                            self._r[idx] -= (_u.volume(x, y) - vol) * _bodies[body[idx - 1]].velocityX(x * _hx, (y + 0.5) * _hx)
                        if y > 0:
                            self._r[idx] -= (_v.volume(x, y) - vol) * _bodies[body[idx - _w]].velocityY((x + 0.5) * _hx, y * _hx)
                        if x < _w - 1:
                            self._r[idx] += (_u.volume(x + 1, y) - vol) * _bodies[body[idx + 1]].velocityX((x + 1.0) * _hx, (y + 0.5) * _hx)
                        if y < _h - 1:
                            self._r[idx] += (_v.volume(x, y + 1) - vol) * _bodies[body[idx + _w]].velocityY((x + 0.5) * _hx, (y + 1.0) * _hx)
                else:
                    self._r[idx] = 0.0
                idx += 1

    def buildPressureMatrix(self, timestep):        # F5:825
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
                        factor = scale * self._u.volume(x + 1, y)
                        self._aDiag[idx] += factor
                        self._aDiag[idx + 1] += factor
                        self._aPlusX[idx] = -factor
                    if y < _h - 1 and cell[idx + _w] == CELL_FLUID:
                        factor = scale * self._v.volume(x, y + 1)
                        self._aDiag[idx] += factor
                        self._aDiag[idx + _w] += factor
                        self._aPlusY[idx] = -factor
                idx += 1
