/* Added next to the reference's physics/solidBodies/{A_SolidBody,SolidBox,SolidSphere}.java (which stay as they are): the
 * fields of A_SolidBody are protected, so the flattening of the `bodies` list for the C ABI lives in their package. */
package physics.solidBodies;

import java.util.List;

/** Flattens the `bodies` list main() shares with FluidSolver (F4:594) into the table IncFluidB200.setBodies consumes:
 *  per body { type (0 = SolidBox, 1 = SolidSphere), _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta }
 *  -- the fields of A_SolidBody.java:73-87, read as they are at the time of the call. */
public final class BodyTable {
    private BodyTable() { }

    public static double[] of(List<A_SolidBody> bodies) {
        final double[] t = new double[bodies.size() * 9];
        int o = 0;
        for (A_SolidBody b : bodies) {
            t[o++] = (b instanceof SolidSphere) ? 1.0 : 0.0;
            t[o++] = b._posX;   t[o++] = b._posY;
            t[o++] = b._scaleX; t[o++] = b._scaleY;
            t[o++] = b._theta;
            t[o++] = b._velX;   t[o++] = b._velY;   t[o++] = b._velTheta;
        }
        return t;
    }
}
