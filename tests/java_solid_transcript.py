"""Literal Python transcription of physics/solidBodies (A_SolidBody.java, SolidBox.java, SolidSphere.java) and of
FluidQuantity.fillSolidFields of ExplicitFluid4solidBoundaries.java (F4:329), made from the Java text -- see
java_f1_transcript.py.  The one documented substitution of this repository applies here too: `rotate` takes cos(phi) and
sin(phi) of the body instead of calling Math.cos / Math.sin (which Java does not specify bit for bit); for phi = -theta
they are cos(theta) and -sin(theta)."""
import math


def length(x, y):
    return math.sqrt(x * x + y * y)


def signum(v):                                      # Math.signum
    if v != v or v == 0.0:
        return v
    return 1.0 if v > 0.0 else -1.0


class SolidBody:                                    # A_SolidBody.java
    def __init__(self, posX, posY, scaleX, scaleY, theta, velX, velY, velTheta, cosTheta, sinTheta):
        self._posX, self._posY, self._scaleX, self._scaleY = posX, posY, scaleX, scaleY
        self._theta, self._velX, self._velY, self._velTheta = theta, velX, velY, velTheta
        self._c, self._s = cosTheta, sinTheta

    @staticmethod
    def rotate(x, y, c, s):                         # SB:65 with c = cos(phi), s = sin(phi)
        tmpX, tmpY = x, y
        x = c * tmpX + s * tmpY
        y = -s * tmpX + c * tmpY
        return x, y

    def velocityX(self, x, y):                      # SB:194
        return (self._posY - y) * self._velTheta + self._velX

    def velocityY(self, x, y):                      # SB:205
        return (x - self._posX) * self._velTheta + self._velY


class SolidBox(SolidBody):                          # SolidBox.java
    @staticmethod
    def nsgn(val):                                  # BOX:54
        result = signum(val)
        if abs(result) == 0:
            return 0.0
        return result

    def distance(self, x, y):                       # BOX:81
        x -= self._posX
        y -= self._posY
        rx, ry = self.rotate(x, y, self._c, -self._s)
        dx = abs(rx) - self._scaleX * 0.5
        dy = abs(ry) - self._scaleY * 0.5
        if dx >= 0.0 or dy >= 0.0:
            return length(max(dx, 0.0), max(dy, 0.0))
        return max(dx, dy)

    def distanceNormal(self, nx, ny, x, y):         # BOX:117
        x -= self._posX
        y -= self._posY
        x, y = self.rotate(x, y, self._c, -self._s)
        if abs(x) - self._scaleX * 0.5 > abs(y) - self._scaleY * 0.5:
            nx = self.nsgn(x)
            ny = 0.0
        else:
            nx = 0.0
            ny = self.nsgn(y)
        return self.rotate(nx, ny, self._c, self._s)


class SolidSphere(SolidBody):                       # SolidSphere.java
    def distance(self, x, y):                       # SPH:62
        return length(x - self._posX, y - self._posY) - self._scaleX * 0.5

    def distanceNormal(self, nx, ny, x, y):         # SPH:84
        x -= self._posX
        y -= self._posY
        r = length(x, y)
        if r < 1e-4:
            nx, ny = 1.0, 0.0
        else:
            nx, ny = x / r, y / r
        return nx, ny


CELL_FLUID, CELL_SOLID = 0, 1


def fillSolidFields(w, h, ox, oy, hx, bodies, normalX, normalY):      # F4:329
    """returns (_cell, _body); _normalX / _normalY are updated in place like the Java arrays"""
    cell, body = [None] * (w * h), [0] * (w * h)
    if not bodies:
        return cell, body
    idx = 0
    for iy in range(h):
        for ix in range(w):
            x = (ix + ox) * hx
            y = (iy + oy) * hx
            body[idx] = 0
            d = bodies[0].distance(x, y)
            for i in range(1, len(bodies)):
                di = bodies[i].distance(x, y)
                if di < d:
                    body[idx] = i
                    d = di
            cell[idx] = CELL_SOLID if d < 0.0 else CELL_FLUID
            normalX[idx], normalY[idx] = bodies[body[idx]].distanceNormal(normalX[idx], normalY[idx], x, y)
            idx += 1
    return cell, body


# ---- ExplicitFluid5curvedBoundaries.java: marching-squares cell volumes ------------------------------------------
def triangleOccupancy(out1, in_, out2):             # F5:135
    return 0.5 * in_ * in_ / ((out1 - in_) * (out2 - in_))


def trapezoidOccupancy(out1, out2, in1, in2):       # F5:146
    return 0.5 * (-in1 / (out1 - in1) - in2 / (out2 - in2))


def occupancy(d11, d12, d21, d22):                  # F5:165
    ds = [d11, d12, d22, d21]
    b = 0
    for i in range(3, -1, -1):
        b = (b << 1) | (1 if ds[i] < 0.0 else 0)
    if b == 0x0: return 0.0
    if b == 0x1: return triangleOccupancy(d21, d11, d12)
    if b == 0x2: return triangleOccupancy(d11, d12, d22)
    if b == 0x4: return triangleOccupancy(d12, d22, d21)
    if b == 0x8: return triangleOccupancy(d22, d21, d11)
    if b == 0xE: return 1.0 - triangleOccupancy(-d21, -d11, -d12)
    if b == 0xD: return 1.0 - triangleOccupancy(-d11, -d12, -d22)
    if b == 0xB: return 1.0 - triangleOccupancy(-d12, -d22, -d21)
    if b == 0x7: return 1.0 - triangleOccupancy(-d22, -d21, -d11)
    if b == 0x3: return trapezoidOccupancy(d21, d22, d11, d12)
    if b == 0x6: return trapezoidOccupancy(d11, d21, d12, d22)
    if b == 0x9: return trapezoidOccupancy(d12, d22, d11, d21)
    if b == 0xC: return trapezoidOccupancy(d11, d12, d21, d22)
    if b == 0x5: return triangleOccupancy(d11, d12, d22) + triangleOccupancy(d22, d21, d11)
    if b == 0xA: return triangleOccupancy(d21, d11, d12) + triangleOccupancy(d12, d22, d21)
    if b == 0xF: return 1.0
    return 0.0


def fillSolidFields5(w, h, ox, oy, hx, bodies, normalX, normalY):     # F5:463
    """returns (_phi, _volume, _cell, _body)"""
    from java_f1_transcript import jmin
    phi = [0.0] * ((w + 1) * (h + 1))
    volume, cell, body = [1.0] * (w * h), [CELL_FLUID] * (w * h), [0] * (w * h)
    if not bodies:
        return phi, volume, cell, body
    idx = 0
    for iy in range(h + 1):
        for ix in range(w + 1):
            x = (ix + ox - 0.5) * hx
            y = (iy + oy - 0.5) * hx
            phi[idx] = bodies[0].distance(x, y)
            for i in range(1, len(bodies)):
                phi[idx] = jmin(phi[idx], bodies[i].distance(x, y))
            idx += 1
    idx = 0
    for iy in range(h):
        for ix in range(w):
            x = (ix + ox) * hx
            y = (iy + oy) * hx
            body[idx] = 0
            d = bodies[0].distance(x, y)
            for i in range(1, len(bodies)):
                di = bodies[i].distance(x, y)
                if di < d:
                    body[idx] = i
                    d = di
            idxp = ix + iy * (w + 1)
            volume[idx] = 1.0 - occupancy(phi[idxp], phi[idxp + 1], phi[idxp + w + 1], phi[idxp + w + 2])
            if volume[idx] < 0.01:
                volume[idx] = 0.0
            normalX[idx], normalY[idx] = bodies[body[idx]].distanceNormal(normalX[idx], normalY[idx], x, y)
            cell[idx] = CELL_SOLID if volume[idx] == 0.0 else CELL_FLUID
            idx += 1
    return phi, volume, cell, body
