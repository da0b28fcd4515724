/* ExplicitFluid8flip of the reference (physics/hybridFluids/ExplicitFluid8flip.java) with FluidSolver delegating to the CUDA library:
 * same package, class, nested-class and method names and signatures (constructor F8:858, update F8:1306, addInflow F8:1396,
 * toImage F8:1417, ambientT F8:1407), the same driver loop (F8:68-107).  All arithmetic of the reference's FluidSolver /
 * FluidQuantity / ParticleQuantities runs on the device behind include/incfluid_b200.h (variant IFL_F8_FLIP).
 * Source only: this repository's image has no JDK (INTEGRATION.md). */
package physics.hybridFluids;

import java.io.File;
import java.io.IOException;
import java.util.ArrayList;
import physics.b200.IncFluidB200;
import physics.solidBodies.A_SolidBody;
import physics.solidBodies.BodyTable;
import physics.solidBodies.SolidBox;
import static java.lang.Math.PI;

class ExplicitFluid8flip {

    public static void main(String args[]) {
        final int sizeX = 128, sizeY = 128;
        final double densityAir = 0.1, densitySoot = 0.25, diffusion = 0.01;
        final boolean renderHeat = false;
        final double timestep = 0.0025;
        final ArrayList<A_SolidBody> bodies = new ArrayList<>();
        bodies.add(new SolidBox(0.5, 0.6, 0.7, 0.1, PI * 0.25, 0.0, 0.0, 0.0));
        final FluidSolver solver = new FluidSolver(sizeX, sizeY, densityAir, densitySoot, diffusion, bodies);
        double time = 0.0;
        for (int frame = 1; time < 8.0; frame++) {
            for (int i = 0; i < 4; i++, time += timestep) {
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
            gpu = new IncFluidB200(IncFluidB200.F8_FLIP, w, h, rhoAir, rhoSoot, diffusion);
            gpu.setBodies(BodyTable.of(_bodies));
        }

        private void update(double timestep) {
            gpu.setBodies(BodyTable.of(_bodies));              /* fillSolidFields reads the bodies' current pose every step */
            gpu.update(timestep);
        }

        private void addInflow(
                double x, double y, double w, double h,
                double d, double t, double u, double v) {
            gpu.addInflow(x, y, w, h, d, t, u, v);             /* F8 applies this rectangle inside update() (F8:1330) */
        }

        private double ambientT() {
            return gpu.ambientT();
        }

        private void toImage(File destination, boolean renderHeat) throws IOException {
            gpu.toImage(destination, renderHeat);
        }
    }
}
