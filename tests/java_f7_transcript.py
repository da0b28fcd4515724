"""Literal Python transcription of what ExplicitFluid7variableDensity.java changes with respect to ExplicitFluid6 (the
rest is textually identical, checked mechanically): computeDensities F7:854, the density-weighted pressure matrix F7:877
(with its `_vDensity[_u.I(x, y + 1)]` index quirk kept as written), applyPressure F7:1148 and the order of update()
F7:1212.  See java_f1_transcript.py."""
import java_f6_transcript as f6
from java_f1_transcript import jmax

CELL_FLUID = f6.CELL_FLUID


class FluidSolver(f6.FluidSolver):
    def __init__(self, w, h, rhoAir, rhoSoot, diffusion, bodies):
        super().__init__(w, h, rhoAir, rhoSoot, diffusion, bodies)
        self._uDensity = [0.0] * ((w + 1) * h)
        self._vDensity = [0.0] * (w * (h + 1))

    def computeDensities(self):                     # F7:854
        _w, _h = self._w, self._h
        alpha = (self._densitySoot - self._densityAir) / self._densityAir
        self._uDensity = [0.0] * ((_w + 1) * _h)
        self._vDensity = [0.0] * (_w * (_h + 1))
        uI = lambda x, y: x + y * self._u._w
        vI = lambda x, y: x + y * self._v._w
        for y in range(_h):
            for x in range(_w):
                density = self._densityAir * self._tAmb / self._t.getSrc(x, y) * (1.0 + alpha * self._d.getSrc(x, y))
                density = jmax(density, 0.05 * self._densityAir)
                self._uDensity[uI(x, y)] += 0.5 * density
                self._vDensity[vI(x, y)] += 0.5 * density
                self._uDensity[uI(x + 1, y)] += 0.5 * density
                self._vDensity[vI(x, y + 1)] += 0.5 * density

    def buildPressureMatrix(self, timestep):        # F7:877
        scale = timestep / (self._hx * self._hx)
        cell = self._d._cell
        _w, _h = self._w, self._h
        n = _w * _h
        self._aDiag, self._aPlusX, self._aPlusY = [0.0] * n, [0.0] * n, [0.0] * n
        uI = lambda x, y: x + y * self._u._w
        idx = 0
        for y in range(_h):
            for x in range(_w):
                if cell[idx] == CELL_FLUID:
                    if x < _w - 1 and cell[idx + 1] == CELL_FLUID:
                        factor = scale * self._u.volume(x + 1, y) / self._uDensity[uI(x + 1, y)]
                        self._aDiag[idx] += factor
                        self._aDiag[idx + 1] += factor
                        self._aPlusX[idx] = -factor
                    if y < _h - 1 and cell[idx + _w] == CELL_FLUID:
                        factor = scale * self._v.volume(x, y + 1) / self._vDensity[uI(x, y + 1)]     # sic: _u.I
                        self._aDiag[idx] += factor
                        self._aDiag[idx + _w] += factor
                        self._aPlusY[idx] = -factor
                idx += 1

    def applyPressure(self, timestep):              # F7:1148
        scale = timestep / self._hx
        cell = self._d._cell
        uI = lambda x, y: x + y * self._u._w
        vI = lambda x, y: x + y * self._v._w
        _u, _v, _p = self._u, self._v, self._p
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                if cell[idx] == CELL_FLUID:
                    _u.setSrc(x, y, _u.getSrc(x, y) - scale * _p[idx] / self._uDensity[uI(x, y)])
                    _v.setSrc(x, y, _v.getSrc(x, y) - scale * _p[idx] / self._vDensity[vI(x, y)])
                    _u.setSrc(x + 1, y, _u.getSrc(x + 1, y) + scale * _p[idx] / self._uDensity[uI(x + 1, y)])
                    _v.setSrc(x, y + 1, _v.getSrc(x, y + 1) + scale * _p[idx] / self._vDensity[vI(x, y + 1)])
                idx += 1

    def update(self, timestep):                     # F7:1212
        n = self._w * self._h
        self._d.fillSolidFields(self._bodies)
        self._t.fillSolidFields(self._bodies)
        self._u.fillSolidFields(self._bodies)
        self._v.fillSolidFields(self._bodies)
        self._r[:n] = self._t._src[:n]
        self.buildHeatDiffusionMatrix(timestep)
        self.buildPreconditioner()
        self.project(2000)
        self.heatIterations, self.heatMaxError = self.lastIterations, self.lastMaxError
        self._t._src[:n] = self._p[:n]
        self._t.extrapolate()
        self.addBuoyancy(timestep)
        self.setBoundaryCondition()
        self.buildRhs()
        self.computeDensities()
        self.buildPressureMatrix(timestep)
        self.buildPreconditioner()
        self.project(2000)
        self.applyPressure(timestep)
        self._d.extrapolate()
        self._u.extrapolate()
        self._v.extrapolate()
        self.setBoundaryCondition()
        self._d.advect(timestep, self._u, self._v, self._bodies)
        self._t.advect(timestep, self._u, self._v, self._bodies)
        self._u.advect(timestep, self._u, self._v, self._bodies)
        self._v.advect(timestep, self._u, self._v, self._bodies)
        self._d.swap()
        self._t.swap()
        self._u.swap()
        self._v.swap()
