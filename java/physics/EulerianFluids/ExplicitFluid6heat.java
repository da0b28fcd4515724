/* ExplicitFluid6heat of the reference (physics/EulerianFluids/ExplicitFluid6heat.java) with FluidSolver delegating to the CUDA library:
 * same package, class, nested-class and method names and signatures (constructor F6:763, update F6:1181, addInflow F6:1233,
 * toImage F6:1250, ambientT F6:1240), the same driver loop (F6:67-107).  All arithmetic of the reference's FluidSolver /
 * FluidQuantity runs on the device behind include/incfluid_b200.h (variant IFL_F6_HEAT).
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

class ExplicitFluid6heat {

    public static void main(String args[]) {
        final int sizeX = 128, sizeY = 128;
        final double densityAir = 0.1, densitySoot = 0.1, diffusion = 0.01;
        final boolean renderHeat = true;
        final double timestep = 0.005;
        final ArrayList<A_SolidBody> bodies = new ArrayList<>();
        bodies.add(new SolidBox(0.3, 0.6, 0.1, 0.5, -PI * 0.05, 0.0, 0.0, 0.0));
        final FluidSolver solver = new FluidSolver(sizeX, sizeY, densityAir, densitySoot, diffusion, bodies);
        double time = 0.0;
        for (int frame = 1; time < 8.0; frame++) {
            for (int i = 0; i < 10; i++, time += timestep) {
                solver.addInflow(0.35, 0.9, 0.1, 0.05, 1.0, solver.ambientT() + 300.0, 0.0, 0.0);
                solver.update(timestep);
            }
            try {
                solver.toImage(new File("Frame" + frame + ".ppm"), renderHeat);
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

        FluidSolver(int w, int h, double rhoAir, double rhoSoot, double diffusion,
                ArrayList<A_SolidBody> bodies) {
            _w = w; _h = h; _bodies = bodies;
            gpu = new IncFluidB200(IncFluidB200.F6_HEAT, w, h, rhoAir, rhoSoot, diffusion);
            gpu.setBodies(BodyTable.of(_bodies));
        }

        private void update(double timestep) {
            gpu.setBodies(BodyTable.of(_bodies));              /* fillSolidFields reads the bodies' current pose every step */
            gpu.update(timestep);
        }

        private void addInflow(double x, double y, double w, double h, double d, double t, double u, double v) {
            gpu.addInflow(x, y, w, h, d, t, u, v);
        }

        private double ambientT() {
            return gpu.ambientT();
        }

        private void toImage(File destination, boolean renderHeat) throws IOException {
            gpu.toImage(destination, renderHeat);
        }
    }
}
