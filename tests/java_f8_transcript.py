"""Literal Python transcription of ExplicitFluid8flip.java, made from the Java text -- see java_f1_transcript.py.  First the
pieces that the C ABI's stages let a test isolate (gridToParticles F8:1748, the particle Runge-Kutta F8:1708 -- forward
in time: `+` --, backProject F8:1673, advect F8:1778, fromParticles F8:722 / addSample F8:312, the EMPTY-cell
extrapolation F8:651, count / prune / seed F8:1595-1660 on the injected random stream), then the whole program assembled
from them and from the ExplicitFluid7 transcription (its projection, heat solve, buoyancy, densities and boundary
conditions are textually F8's, checked mechanically)."""
import java_f4_transcript as f4
from java_f1_transcript import jmax, jmin


def gridToParticles(quantities, properties, posX, posY, count, alpha):      # F8:1748
    for t in range(len(quantities)):
        for i in range(count):
            properties[t][i] *= 1.0 - alpha
            properties[t][i] += quantities[t].lerpGrid(posX[i], posY[i])


def rungeKutta3(x, y, timestep, u, v, _hx):         # F8:1708
    firstU = u.lerpGrid(x, y) / _hx
    firstV = v.lerpGrid(x, y) / _hx
    midX = x + 0.5 * timestep * firstU
    midY = y + 0.5 * timestep * firstV
    midU = u.lerpGrid(midX, midY) / _hx
    midV = v.lerpGrid(midX, midY) / _hx
    lastX = x + 0.75 * timestep * midU
    lastY = y + 0.75 * timestep * midV
    lastU = u.lerpGrid(lastX, lastY)                 # sic: no division by _hx
    lastV = v.lerpGrid(lastX, lastY)
    x += timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU)
    y += timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV)
    return x, y


def backProject(x, y, bodies, _hx):                 # F8:1673
    d = 1e30
    closestBody = -1
    for i in range(len(bodies)):
        di = bodies[i].distance(x * _hx, y * _hx)
        if di < d:
            d = di
            closestBody = i
    if d < -1.0:
        x *= _hx
        y *= _hx
        x, y = f4.closestSurfacePoint(bodies[closestBody], x, y)
        nx, ny = bodies[closestBody].distanceNormal(0.0, 0.0, x, y)
        x -= nx * _hx
        y -= ny * _hx
        x /= _hx
        y /= _hx
    return x, y


def advect(posX, posY, count, timestep, u, v, bodies, _w, _h, _hx):         # F8:1778
    for i in range(count):
        posX[i], posY[i] = rungeKutta3(posX[i], posY[i], timestep, u, v, _hx)
        posX[i], posY[i] = backProject(posX[i], posY[i], bodies, _hx)
        posX[i] = jmax(jmin(posX[i], _w - 0.001), 0.0)
        posY[i] = jmax(jmin(posY[i], _h - 0.001), 0.0)


# ---- particle -> grid: FluidQuantity.addSample F8:312 and fromParticles F8:722 (without the extrapolation that follows)
CELL_FLUID, CELL_SOLID, CELL_EMPTY = 0, 1, 2


def addSample(q_src, weight, _w, _h, value, x, y, ix, iy):                  # F8:312
    if ix < 0 or iy < 0 or ix >= _w or iy >= _h:
        return
    k = (1.0 - abs(ix - x)) * (1.0 - abs(iy - y))
    weight[ix + iy * _w] += k
    q_src[ix + iy * _w] += k * value


def fromParticles(_w, _h, _ox, _oy, cell, count, posX, posY, prop):         # F8:722
    _src = [0.0] * (_w * _h)
    weight = [0.0] * (_w * _h)
    for i in range(count):
        x = posX[i] - _ox
        y = posY[i] - _oy
        x = jmax(0.5, jmin(_w - 1.5, x))
        y = jmax(0.5, jmin(_h - 1.5, y))
        ix = int(x)
        iy = int(y)
        addSample(_src, weight, _w, _h, prop[i], x, y, ix + 0, iy + 0)
        addSample(_src, weight, _w, _h, prop[i], x, y, ix + 1, iy + 0)
        addSample(_src, weight, _w, _h, prop[i], x, y, ix + 0, iy + 1)
        addSample(_src, weight, _w, _h, prop[i], x, y, ix + 1, iy + 1)
    for i in range(_w * _h):
        if weight[i] != 0.0:
            _src[i] /= weight[i]
        elif cell[i] == CELL_FLUID:
            cell[i] = CELL_EMPTY
    return _src


# ---- FluidQuantity.extrapolate of ExplicitFluid8flip.java (F8:651): SOLID cells along the normal as in F4, EMPTY cells
# (no particle weight) by the average of their FLUID neighbours, EMPTY cells on the domain border from their inner
# neighbour; every EMPTY cell ends up FLUID.
def isignum(v):
    return int(v and (1 if v > 0 else -1))


def extrapolate8(_src, cell, normalX, normalY, _w, _h):
    from java_f1_transcript import f32
    mask = [0] * (_w * _h)
    # fillSolidMask F8:~520
    for x in range(_w):
        mask[x] = mask[x + (_h - 1) * _w] = 0xFF
    for y in range(_h):
        mask[y * _w] = mask[y * _w + _w - 1] = 0xFF
    for y in range(1, _h - 1):
        for x in range(1, _w - 1):
            idx = x + y * _w
            mask[idx] = 0
            if cell[idx] == CELL_SOLID:
                nx, ny = normalX[idx], normalY[idx]
                if nx != 0.0 and cell[idx + isignum(nx)] != CELL_FLUID:
                    mask[idx] |= 1
                if ny != 0.0 and cell[idx + isignum(ny) * _w] != CELL_FLUID:
                    mask[idx] |= 2
            elif cell[idx] == CELL_EMPTY:
                mask[idx] = 1 if (cell[idx - 1] != CELL_FLUID and cell[idx + 1] != CELL_FLUID
                                  and cell[idx - _w] != CELL_FLUID and cell[idx + _w] != CELL_FLUID) else 0

    def extrapolateNormal(idx):                     # as F4:426
        nx, ny = normalX[idx], normalY[idx]
        srcX = _src[idx + isignum(nx)]
        srcY = _src[idx + isignum(ny) * _w]
        return (abs(f32(nx)) * srcX + abs(f32(ny)) * srcY) / (abs(f32(nx)) + abs(f32(ny)))

    def extrapolateAverage(idx):                    # F8:~545
        value, count = 0.0, 0
        if cell[idx - 1] == CELL_FLUID:
            value += _src[idx - 1]; count += 1
        if cell[idx + 1] == CELL_FLUID:
            value += _src[idx + 1]; count += 1
        if cell[idx - _w] == CELL_FLUID:
            value += _src[idx - _w]; count += 1
        if cell[idx + _w] == CELL_FLUID:
            value += _src[idx + _w]; count += 1
        return value / count

    def freeSolidNeighbour(idx, border, m):         # F8:567
        if cell[idx] == CELL_SOLID:
            mask[idx] &= ~m
            if mask[idx] == 0:
                border.append(idx)

    def freeEmptyNeighbour(idx, border):            # F8:~585
        if cell[idx] == CELL_EMPTY and mask[idx] == 1:
            mask[idx] = 0
            border.append(idx)

    border = []
    for y in range(1, _h - 1):
        for x in range(1, _w - 1):
            idx = x + y * _w
            if cell[idx] != CELL_FLUID and mask[idx] == 0:
                border.append(idx)
    while border:
        idx = border.pop()
        if cell[idx] == CELL_EMPTY:
            _src[idx] = extrapolateAverage(idx)
            cell[idx] = CELL_FLUID
        else:
            _src[idx] = extrapolateNormal(idx)
        if normalX[idx - 1] > 0.0:
            freeSolidNeighbour(idx - 1, border, 1)
        if normalX[idx + 1] < 0.0:
            freeSolidNeighbour(idx + 1, border, 1)
        if normalY[idx - _w] > 0.0:
            freeSolidNeighbour(idx - _w, border, 2)
        if normalY[idx + _w] < 0.0:
            freeSolidNeighbour(idx + _w, border, 2)
        freeEmptyNeighbour(idx - 1, border)
        freeEmptyNeighbour(idx + 1, border)
        freeEmptyNeighbour(idx - _w, border)
        freeEmptyNeighbour(idx + _w, border)
    # extrapolateEmptyBorders F8:~600
    for x in range(1, _w - 1):
        idxT, idxB = x, x + (_h - 1) * _w
        if cell[idxT] == CELL_EMPTY:
            _src[idxT] = _src[idxT + _w]
        if cell[idxB] == CELL_EMPTY:
            _src[idxB] = _src[idxB - _w]
    for y in range(1, _h - 1):
        idxL, idxR = y * _w, y * _w + _w - 1
        if cell[idxL] == CELL_EMPTY:
            _src[idxL] = _src[idxL + 1]
        if cell[idxR] == CELL_EMPTY:
            _src[idxR] = _src[idxR - 1]
    idxTL, idxTR, idxBL, idxBR = 0, _w - 1, (_h - 1) * _w, _h * _w - 1
    if cell[idxTL] == CELL_EMPTY:
        _src[idxTL] = 0.5 * (_src[idxTL + 1] + _src[idxTL + _w])
    if cell[idxTR] == CELL_EMPTY:
        _src[idxTR] = 0.5 * (_src[idxTR - 1] + _src[idxTR + _w])
    if cell[idxBL] == CELL_EMPTY:
        _src[idxBL] = 0.5 * (_src[idxBL + 1] + _src[idxBL - _w])
    if cell[idxBR] == CELL_EMPTY:
        _src[idxBR] = 0.5 * (_src[idxBR - 1] + _src[idxBR - _w])
    for i in range(_w * _h):
        if cell[i] == CELL_EMPTY:
            cell[i] = CELL_FLUID


# ---- ParticleQuantities bookkeeping after the transfer (particlesToGrid F8:1761): countParticles F8:1595,
# pruneParticles F8:1610, seedParticles F8:1637.  `Math.random()` is the injected stream of include/incfluid_b200.h
# (splitmix64 finaliser of seed + golden-ratio * (k + 1), top 24 bits as a float in [0, 1)); `x + (float) random()` is
# a FLOAT addition (int promoted to float), widened on assignment.
class InjectedRandom:
    def __init__(self, seed):
        self.seed, self.k = seed, 0

    def next_float(self):
        M = (1 << 64) - 1
        z = (self.seed + 0x9E3779B97F4A7C15 * (self.k + 1)) & M
        self.k += 1
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z = z ^ (z >> 31)
        return (z >> 40) / 16777216.0               # exact: a 24-bit integer times 2^-24


def countParticles(_w, _h, count, posX, posY):      # F8:1595
    counts = [0] * (_w * _h)
    for i in range(count):
        ix = int(posX[i])
        iy = int(posY[i])
        if ix >= 0 and iy >= 0 and ix < _w and iy < _h:
            counts[ix + iy * _w] += 1
    return counts


def pruneParticles(_w, _h, count, posX, posY, properties, counts, _MaxPerCell):      # F8:1610
    i = 0
    while i < count:
        ix = int(posX[i])
        iy = int(posY[i])
        idx = ix + iy * _w
        if ix < 0 and iy < 0 and ix >= _w and iy >= _h:      # sic: && -- never true
            i += 1
            continue
        if counts[idx] > _MaxPerCell:
            count -= 1
            j = count
            posX[i] = posX[j]
            posY[i] = posY[j]
            for t in range(len(properties)):
                properties[t][i] = properties[t][j]
            counts[idx] -= 1
            i -= 1
        i += 1
    return count


def seedParticles(_w, _h, count, maxParticles, posX, posY, properties, quantities, counts, _MinPerCell, rnd, pointInBody):   # F8:1637
    from java_f1_transcript import f32
    idx = 0
    for y in range(_h):
        for x in range(_w):
            for i in range(_MinPerCell - counts[idx]):
                if count == maxParticles:
                    return count
                j = count
                posX[j] = f32(x + rnd.next_float())
                posY[j] = f32(y + rnd.next_float())
                if pointInBody(posX[idx], posY[idx]):        # sic: [idx], the cell index
                    continue
                for t in range(len(quantities)):
                    properties[t][j] = quantities[t].lerpGrid(posX[j], posY[j])
                count += 1
            idx += 1
    return count


# ---- the whole program: FluidQuantity / ParticleQuantities / FluidSolver of ExplicitFluid8flip.java, assembled from
# the pieces above and from ExplicitFluid7 (whose projection, heat solve, buoyancy, densities and boundary conditions
# are textually F8's, checked mechanically).
import java_f5_transcript as _f5
import java_f7_transcript as _f7


class FluidQuantity(_f5.FluidQuantity):             # F8:196
    def __init__(self, w, h, ox, oy, hx):
        super().__init__(w, h, ox, oy, hx)
        self._old = [0.0] * (w * h)
        del self._dst

    def copy(self):                                 # F8:340
        self._old[:] = self._src

    def diff(self, alpha):                          # F8:346
        for i in range(self._w * self._h):
            self._src[i] -= (1.0 - alpha) * self._old[i]

    def undiff(self, alpha):                        # F8:355
        for i in range(self._w * self._h):
            self._src[i] += (1.0 - alpha) * self._old[i]

    def extrapolate(self):                          # F8:651
        extrapolate8(self._src, self._cell, self._normalX, self._normalY, self._w, self._h)

    def fromParticles(self, count, posX, posY, prop):   # F8:722
        self._src = fromParticles(self._w, self._h, self._ox, self._oy, self._cell, count, posX, posY, prop)


class ParticleQuantities:                           # F8:1479
    def __init__(self, w, h, hx, bodies, rnd, avg=4, maxpc=12, minpc=3):
        self._w, self._h, self._hx, self._bodies, self._rnd = w, h, hx, bodies, rnd
        self._AvgPerCell, self._MaxPerCell, self._MinPerCell = avg, maxpc, minpc
        self._maxParticles = w * h * maxpc
        self._posX = [0.0] * self._maxParticles
        self._posY = [0.0] * self._maxParticles
        self._properties, self._quantities = [], []
        self._counts = [0] * (w * h)
        self.initParticles()

    def pointInBody(self, x, y):                    # F8:1566
        return any(b.distance(x * self._hx, y * self._hx) < 0.0 for b in self._bodies)

    def initParticles(self):                        # F8:1571
        from java_f1_transcript import f32
        idx = 0
        for y in range(self._h):
            for x in range(self._w):
                i = 0
                while i < self._AvgPerCell:
                    self._posX[idx] = f32(x + self._rnd.next_float())
                    self._posY[idx] = f32(y + self._rnd.next_float())
                    if self.pointInBody(self._posX[idx], self._posY[idx]):
                        idx -= 1
                    i += 1
                    idx += 1
        self._particleCount = idx

    def addQuantity(self, q):                       # F8:1740
        self._quantities.append(q)
        self._properties.append([0.0] * self._maxParticles)

    def gridToParticles(self, alpha):               # F8:1748
        gridToParticles(self._quantities, self._properties, self._posX, self._posY, self._particleCount, alpha)

    def particlesToGrid(self):                      # F8:1761
        for t in range(len(self._quantities)):
            self._quantities[t].fromParticles(self._particleCount, self._posX, self._posY, self._properties[t])
            self._quantities[t].extrapolate()
        self._counts = countParticles(self._w, self._h, self._particleCount, self._posX, self._posY)
        self._particleCount = pruneParticles(self._w, self._h, self._particleCount, self._posX, self._posY,
                                             self._properties, self._counts, self._MaxPerCell)
        self._particleCount = seedParticles(self._w, self._h, self._particleCount, self._maxParticles, self._posX,
                                            self._posY, self._properties, self._quantities, self._counts,
                                            self._MinPerCell, self._rnd, self.pointInBody)

    def advect(self, timestep, u, v):               # F8:1778
        advect(self._posX, self._posY, self._particleCount, timestep, u, v, self._bodies, self._w, self._h, self._hx)


class FluidSolver(_f7.FluidSolver):                 # F8:760
    def __init__(self, w, h, rhoAir, rhoSoot, diffusion, bodies, rnd, avg=4, maxpc=12, minpc=3):     # F8:858
        super().__init__(w, h, rhoAir, rhoSoot, diffusion, bodies)
        self._flipAlpha = 0.001
        self._d = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._t = FluidQuantity(w, h, 0.5, 0.5, self._hx)
        self._u = FluidQuantity(w + 1, h, 0.0, 0.5, self._hx)
        self._v = FluidQuantity(w, h + 1, 0.5, 0.0, self._hx)
        for i in range(w * h):
            self._t._src[i] = self._tAmb
        self._qs = ParticleQuantities(w, h, self._hx, bodies, rnd, avg, maxpc, minpc)
        for q in (self._d, self._t, self._u, self._v):
            self._qs.addQuantity(q)
        self._qs.gridToParticles(1.0)

    def update(self, timestep):                     # F8:1306
        n = self._w * self._h
        self._d.fillSolidFields(self._bodies)
        self._t.fillSolidFields(self._bodies)
        self._u.fillSolidFields(self._bodies)
        self._v.fillSolidFields(self._bodies)
        self._qs.particlesToGrid()
        self._d.copy()
        self._t.copy()
        self._u.copy()
        self._v.copy()
        self.addInflow(0.45, 0.2, 0.2, 0.05, 1.0, self._tAmb, 0.0, 0.0)
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
        a = self._flipAlpha
        self._d.diff(a)
        self._t.diff(a)
        self._u.diff(a)
        self._v.diff(a)
        self._qs.gridToParticles(a)
        self._d.undiff(a)
        self._t.undiff(a)
        self._u.undiff(a)
        self._v.undiff(a)
        self._qs.advect(timestep, self._u, self._v)
