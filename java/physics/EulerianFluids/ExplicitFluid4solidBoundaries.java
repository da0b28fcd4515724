/* ExplicitFluid4solidBoundaries of the reference (physics/EulerianFluids/ExplicitFluid4solidBoundaries.java) with FluidSolver delegating to the CUDA library:
 * same package, class, nested-class and method names and signatures (constructor F4:594, update F4:617, addInflow F4:653,
 * toImage F4:665), the same driver loop (F4:67-100).  All arithmetic of the reference's FluidSolver /
 * FluidQuantity runs on the device behind include/incfluid_b200.h (variant IFL_F4_SOLID_BOUNDARIES).
 * Source only: this repository's image has no JDK (INTEGRATION.md). */
package physics.EulerianFluids;

import java.io.File;
import java.io.IOException;
import java.util.ArrayList;
import physics.b200.IncFluidB200;
import physics.solidBodies.A_SolidBody;
import physics.solidBodies.BodyTable;
import physics.solidBodies.SolidBox;
import static java.lang.Math.PI;

class ExplicitFluid4solidBoundaries {

    public static void main(String args[]) {
        final int sizeX = 128, sizeY = 128;
        final double density = 0.1;
        final double timestep = 0.005;
        final ArrayList<A_SolidBody> bodies = new ArrayList<>();
        bodies.add(new SolidBox(0.5, 0.6, 0.7, 0.1, PI * 0.25, 0.0, 0.0, 0.0));
        final FluidSolver solver = new FluidSolver(sizeX, sizeY, density, bodies);
        double time = 0.0;
        for (int frame = 1; time < 8.0; frame++) {
            for (int i = 0; i < 4; i++, time += timestep) {
                solver.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0);
                solver.update(timestep);
            }
            try {
                solver.toImage(new File("Frame" + frame + ".ppm"));
            } catch (IOException ex) {
                ex.printStackTrace();
            }
            for (A_SolidBody body : bodies) body.update(timestep);      /* once per frame, not per substep */
        }
    }

    static class FluidSolver {
        private final IncFluidB200 gpu;
        private final int _w, _h;
        private final ArrayList<A_SolidBody> _bodies;      /* shared with main(), which moves the bodies between frames */

        FluidSolver(int w, int h, double density, ArrayList<A_SolidBody> bodies) {
            _w = w; _h = h; _bodies = bodies;
            gpu = new IncFluidB200(IncFluidB200.F4_SOLID_BOUNDARIES, w, h, density);
            gpu.setBodies(BodyTable.of(_bodies));
        }

        private void update(double timestep) {
            gpu.setBodies(BodyTable.of(_bodies));              /* fillSolidFields reads the bodies' current pose every step */
            gpu.update(timestep);
        }

        private void addInflow(double x, double y, double w, double h, double d, double u, double v) {
            gpu.addInflow(x, y, w, h, d, 0.0, u, v);
        }

        private void toImage(File destination) throws IOException {
            gpu.toImage(destination, false);
        }
    }
}
