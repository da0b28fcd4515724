"""Literal Python transcription of what ExplicitFluid6heat.java changes with respect to ExplicitFluid5 (the rest of the
two files is textually identical, checked mechanically): the temperature quantity, the implicit heat diffusion solve
(buildHeatDiffusionMatrix F6:880 + the same MIC(0)-PCG), buoyancy F6:1135, the vector operations restricted to fluid
cells (F6:999, 1041, 1054) and the order of update() F6:1181.  See java_f1_transcript.py."""
import java_f5_transcript as f5
from java_f1_transcript import f32, jmax

CELL_FLUID = f5.CELL_FLUID


class FluidSolver(f5.FluidSolver):
    def __init__(self, w, h, rhoAir, rhoSoot, diffusion, bodies):     # F6:763
        super().__init__(w, h, rhoAir, bodies)
        self._densityAir, self._densitySoot, self._diffusion = rhoAir, rhoSoot, diffusion
        self._tAmb = 294.0
        self._g = 9.81
        self._t = f5.FluidQuantity(w, h, 0.5, 0.5, self._hx)
        for i in range(w * h):
            self._t._src[i] = self._tAmb
        self.heatIterations = self.heatMaxError = None

    def buildHeatDiffusionMatrix(self, timestep):   # F6:880
        _w, _h = self._w, self._h
        for i in range(_w * _h):
            self._aDiag[i] = 1.0
        self._aPlusX, self._aPlusY = [0.0] * (_w * _h), [0.0] * (_w * _h)
        cell = self._d._cell
        scale = self._diffusion * timestep * 1.0 / (self._hx * self._hx)
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

    def dotProduct(self, a, b):                     # F6:999
        cell = self._d._cell
        result = 0.0
        for i in range(self._w * self._h):
            if cell[i] == CELL_FLUID:
                result += a[i] * b[i]
        return result

    def scaledAdd(self, dst, a, b, s):              # F6:1041
        cell = self._d._cell
        for i in range(self._w * self._h):
            if cell[i] == CELL_FLUID:
                dst[i] = a[i] + b[i] * s

    def infinityNorm(self, a):                      # F6:1054
        cell = self._d._cell
        maxA = 0.0
        for i in range(self._w * self._h):
            if cell[i] == CELL_FLUID:
                maxA = jmax(maxA, abs(f32(a[i])))
        return maxA

    def addBuoyancy(self, timestep):                # F6:1135
        alpha = (self._densitySoot - self._densityAir) / self._densityAir
        for y in range(self._h):
            for x in range(self._w):
                buoyancy = timestep * self._g * (alpha * self._d.getSrc(x, y) - (self._t.getSrc(x, y) - self._tAmb) / self._tAmb)
                self._v.setSrc(x, y, self._v.getSrc(x, y) + buoyancy * 0.5)
                self._v.setSrc(x, y + 1, self._v.getSrc(x, y + 1) + buoyancy * 0.5)

    def update(self, timestep):                     # F6:1181
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

    def addInflow(self, x, y, w, h, d, t, u, v):    # F6:1233
        self._d.addInflow(x, y, x + w, y + h, d)
        self._t.addInflow(x, y, x + w, y + h, t)
        self._u.addInflow(x, y, x + w, y + h, u)
        self._v.addInflow(x, y, x + w, y + h, v)
