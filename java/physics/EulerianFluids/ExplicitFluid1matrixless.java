/* ExplicitFluid1matrixless of the reference (physics/EulerianFluids/ExplicitFluid1matrixless.java) with FluidSolver delegating to the CUDA library:
 * same package, class, nested-class and method names and signatures (constructor F1:262, update F1:276, addInflow F1:298,
 * toImage F1:341, maxTimestep F1:309), the same driver loop (F1:55-88).  All arithmetic of the reference's FluidSolver /
 * FluidQuantity runs on the device behind include/incfluid_b200.h (variant IFL_F1_MATRIXLESS).
 * Source only: this repository's image has no JDK (INTEGRATION.md). */
package physics.EulerianFluids;

import java.io.File;
import java.io.IOException;
import physics.b200.IncFluidB200;

class ExplicitFluid1matrixless {

    public static void main(String args[]) {
        final int sizeX = 128, sizeY = 128;
        final double density = 0.1;
        final double timestep = 0.005;
        final FluidSolver solver = new FluidSolver(sizeX, sizeY, density);
        double time = 0.0;
        for (int frame = 1; time < 8.0; frame++) {
            for (int i = 0; i < 4; i++, time += timestep) {
                solver.addInflow(0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 3.0);
                solver.update(timestep);
            }
            try {
                solver.toImage(new File("Frame" + frame + ".ppm"));
            } catch (IOException ex) {
                ex.printStackTrace();
            }
        }
    }

    static class FluidSolver {
        private final IncFluidB200 gpu;
        private final int _w, _h;

        FluidSolver(int w, int h, double density) {
            _w = w; _h = h;
            gpu = new IncFluidB200(IncFluidB200.F1_MATRIXLESS, w, h, density);
        }

        private void update(double timestep) {
            gpu.update(timestep);
        }

        private void addInflow(double x, double y, double w, double h, double d, double u, double v) {
            gpu.addInflow(x, y, w, h, d, 0.0, u, v);
        }

        private double maxTimestep() {
            return gpu.maxTimestep();
        }

        private void toImage(File destination) throws IOException {
            gpu.toImage(destination, false);
        }
    }
}
