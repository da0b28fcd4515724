/* Panama FFM binding of include/incfluid_b200.h for the reference's physics.* classes (INTEGRATION.md, section 2).
 * NOT compiled in this repository's image (no JDK here): build with JDK 22+,
 *   javac -d out $(find java -name '*.java') -cp lib/commons-math3-3.4.1.jar:<reference>/src
 *   java --enable-native-access=ALL-UNNAMED ...
 * tests/test_java_api.py checks that every C symbol bound here is declared in the header, that the struct layouts agree
 * with the ctypes mirror and that the program classes keep the reference's class and method signatures. */
package physics.b200;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

/** Thin binding of include/incfluid_b200.h.  One instance = one ifl_solver handle (not thread safe,
 *  like the reference).  Field layouts mirror `struct ifl_config` / `ifl_step_status` / `ifl_body` byte for byte. */
public final class IncFluidB200 implements AutoCloseable {
    /** enum ifl_variant: the reference program reproduced */
    public static final int F1_MATRIXLESS = 1, F2_BETTER_ADVECTION = 2, F3_CONJUGATE_GRADIENTS = 3, F4_SOLID_BOUNDARIES = 4,
            F5_CURVED_BOUNDARIES = 5, F6_HEAT = 6, F7_VARIABLE_DENSITY = 7, F8_FLIP = 8;
    /** enum ifl_field (the ones a Java caller touches: FluidQuantity._src of d, t, u, v and the pressure) */
    public static final int FIELD_D = 0, FIELD_T = 1, FIELD_U = 2, FIELD_V = 3, FIELD_P = 8;
    /** doubles per body in the table handed to setBodies: type, posX, posY, scaleX, scaleY, theta, velX, velY, velTheta */
    public static final int BODY_DOUBLES = 9;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libincfluid_b200.so", Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }
    // int ifl_default_config(ifl_config*, int32 variant, int32 w, int32 h)
    private static final MethodHandle DEFAULT_CONFIG = fn("ifl_default_config", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT));
    // int ifl_create(const ifl_config*, ifl_solver**)          -- new FluidSolver(...)  F1:262 F3:339 F4:594 F6:763
    private static final MethodHandle CREATE = fn("ifl_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle DESTROY = fn("ifl_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    // int ifl_add_inflow(s, x, y, w, h, d, t, u, v)            -- FluidSolver.addInflow  F1:298 F6:1233
    private static final MethodHandle ADD_INFLOW = fn("ifl_add_inflow", FunctionDescriptor.of(JAVA_INT, ADDRESS,
            JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE));
    // int ifl_update(s, timestep, ifl_step_status*)            -- FluidSolver.update     F1:276 ... F8:1306
    private static final MethodHandle UPDATE = fn("ifl_update", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE, ADDRESS));
    // int ifl_max_timestep(s, double*)                         -- FluidSolver.maxTimestep F1:309
    private static final MethodHandle MAX_TIMESTEP = fn("ifl_max_timestep", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    // int ifl_set_bodies(s, const ifl_body*, int32 count)      -- the shared `bodies` list F4:594
    private static final MethodHandle SET_BODIES = fn("ifl_set_bodies", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    // int ifl_read_field / ifl_write_field(s, int32 field, double*, size_t count)   -- solver._d._src[i]  F1:351
    private static final MethodHandle READ_FIELD = fn("ifl_read_field", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle WRITE_FIELD = fn("ifl_write_field", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
    // int ifl_write_ppm(s, const char* path, int32 render_heat) -- FluidSolver.toImage F1:341 ... F6:1250, F8:1417
    private static final MethodHandle WRITE_PPM = fn("ifl_write_ppm", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    // int ifl_particle_count(s, int32*)                        -- ParticleQuantities._particleCount F8:1771
    private static final MethodHandle PARTICLE_COUNT = fn("ifl_particle_count", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("ifl_last_error", FunctionDescriptor.of(ADDRESS));
    // several GPUs, one JVM per GPU (INTEGRATION.md section 3; the reference has no distributed path):
    // int ifl_comm_unique_id(uint8_t out[128]); int ifl_slab_rows(s, int32 field, int32* row0, int32* nrows);
    // int ifl_read_field_rows / ifl_write_field_rows(s, int32 field, int32 row0, int32 nrows, double*, size_t count)
    private static final MethodHandle COMM_UNIQUE_ID = fn("ifl_comm_unique_id", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle SLAB_ROWS = fn("ifl_slab_rows", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle READ_FIELD_ROWS = fn("ifl_read_field_rows",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle WRITE_FIELD_ROWS = fn("ifl_write_field_rows",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG));

    /** struct ifl_config (include/incfluid_b200.h) */
    static final StructLayout CONFIG = MemoryLayout.structLayout(
        JAVA_INT.withName("abi_version"), JAVA_INT.withName("variant"), JAVA_INT.withName("w"), JAVA_INT.withName("h"),
        JAVA_DOUBLE.withName("density"), JAVA_DOUBLE.withName("density_soot"), JAVA_DOUBLE.withName("diffusion"),
        JAVA_DOUBLE.withName("t_ambient"), JAVA_DOUBLE.withName("gravity"), JAVA_DOUBLE.withName("flip_alpha"),
        JAVA_DOUBLE.withName("tolerance"), JAVA_INT.withName("solver_limit"), MemoryLayout.paddingLayout(4),
        JAVA_DOUBLE.withName("mic_tau"), JAVA_DOUBLE.withName("mic_sigma"),
        JAVA_INT.withName("particles_avg_per_cell"), JAVA_INT.withName("particles_max_per_cell"),
        JAVA_INT.withName("particles_min_per_cell"), MemoryLayout.paddingLayout(4),
        JAVA_LONG.withName("rng_seed"), JAVA_INT.withName("device"), JAVA_INT.withName("rank"),
        JAVA_INT.withName("world_size"), JAVA_INT.withName("flags"),
        MemoryLayout.sequenceLayout(128, JAVA_BYTE).withName("comm_id"));
    /** struct ifl_step_status */
    static final StructLayout STATUS = MemoryLayout.structLayout(
        JAVA_INT.withName("iterations_pressure"), JAVA_INT.withName("iterations_heat"),
        JAVA_DOUBLE.withName("residual_pressure"), JAVA_DOUBLE.withName("residual_heat"),
        JAVA_INT.withName("exceeded_budget"), JAVA_INT.withName("nan_residual"),
        JAVA_INT.withName("particle_count"), JAVA_INT.withName("reserved"));
    /** struct ifl_body: the fields of A_SolidBody plus cos/sin of _theta, evaluated once here */
    static final StructLayout BODY = MemoryLayout.structLayout(
        JAVA_INT.withName("type"), JAVA_INT.withName("reserved"),
        JAVA_DOUBLE.withName("pos_x"), JAVA_DOUBLE.withName("pos_y"),
        JAVA_DOUBLE.withName("scale_x"), JAVA_DOUBLE.withName("scale_y"),
        JAVA_DOUBLE.withName("theta"), JAVA_DOUBLE.withName("cos_theta"), JAVA_DOUBLE.withName("sin_theta"),
        JAVA_DOUBLE.withName("vel_x"), JAVA_DOUBLE.withName("vel_y"), JAVA_DOUBLE.withName("vel_theta"));

    private final Arena arena = Arena.ofConfined();
    private final MemorySegment handle;
    private final MemorySegment status = arena.allocate(STATUS);
    private final double tAmbient;

    /** FluidSolver(w, h, density) of F1-F3 and FluidSolver(w, h, density, bodies) of F4-F5 */
    public IncFluidB200(int variant, int w, int h, double density) { this(variant, w, h, density, Double.NaN, Double.NaN); }

    /** FluidSolver(w, h, rhoAir, rhoSoot, diffusion, bodies) of F6-F8 (NaN = keep the default of ifl_default_config) */
    public IncFluidB200(int variant, int w, int h, double density, double densitySoot, double diffusion) {
        this(variant, w, h, density, densitySoot, diffusion, 0, 0, 1, null);
    }

    /** One rank of a solver group on several GPUs (one JVM per GPU): F3 slab-decomposed, F4-F7 with slab-decomposed solver
     *  loops.  commId = the 128 bytes rank 0 got from commUniqueId() and handed to the other JVMs. */
    public IncFluidB200(int variant, int w, int h, double density, double densitySoot, double diffusion,
                        int device, int rank, int worldSize, byte[] commId) {
        try {
            MemorySegment cfg = arena.allocate(CONFIG);
            check((int) DEFAULT_CONFIG.invokeExact(cfg, variant, w, h));
            cfg.set(JAVA_DOUBLE, offset("density"), density);
            if (!Double.isNaN(densitySoot)) cfg.set(JAVA_DOUBLE, offset("density_soot"), densitySoot);
            if (!Double.isNaN(diffusion)) cfg.set(JAVA_DOUBLE, offset("diffusion"), diffusion);
            cfg.set(JAVA_INT, offset("device"), device);
            cfg.set(JAVA_INT, offset("rank"), rank);
            cfg.set(JAVA_INT, offset("world_size"), worldSize);
            if (commId != null) MemorySegment.copy(commId, 0, cfg, JAVA_BYTE, offset("comm_id"), 128);
            tAmbient = cfg.get(JAVA_DOUBLE, offset("t_ambient"));
            MemorySegment out = arena.allocate(ADDRESS);
            check((int) CREATE.invokeExact(cfg, out));
            handle = out.get(ADDRESS, 0);
        } catch (Throwable t) { throw wrap(t); }
    }
    private static long offset(String field) { return CONFIG.byteOffset(MemoryLayout.PathElement.groupElement(field)); }

    public void addInflow(double x, double y, double w, double h, double d, double t, double u, double v) {
        try { check((int) ADD_INFLOW.invokeExact(handle, x, y, w, h, d, t, u, v)); } catch (Throwable e) { throw wrap(e); }
    }
    /** returns the iteration count the reference only logs (F3:613,628) */
    public int update(double timestep) {
        try {
            int rc = (int) UPDATE.invokeExact(handle, timestep, status);
            if (rc == 4 /* IFL_ERR_NAN: the reference calls System.exit(-1) here, F3:616-619 */) System.exit(-1);
            check(rc);
            return status.get(JAVA_INT, 0);
        } catch (Throwable e) { throw wrap(e); }
    }
    public double maxTimestep() {
        try {
            MemorySegment out = arena.allocate(JAVA_DOUBLE);
            check((int) MAX_TIMESTEP.invokeExact(handle, out));
            return out.get(JAVA_DOUBLE, 0);
        } catch (Throwable e) { throw wrap(e); }
    }
    /** _tAmb (F6:772): FluidSolver.ambientT() */
    public double ambientT() { return tAmbient; }
    /** the `bodies` list (F4:594); table = BODY_DOUBLES doubles per body, see physics.solidBodies.BodyTable.
     *  Call again after main() has advanced the bodies (body.update(dt), F4:98). */
    public void setBodies(double[] table) {
        final int n = table.length / BODY_DOUBLES;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocate(BODY, Math.max(n, 1));
            for (int i = 0; i < n; i++) {
                final long o = i * BODY.byteSize();
                final double theta = table[i * BODY_DOUBLES + 5];
                buf.set(JAVA_INT, o, (int) table[i * BODY_DOUBLES]);
                buf.set(JAVA_INT, o + 4, 0);
                for (int k = 0; k < 5; k++) buf.set(JAVA_DOUBLE, o + 8 + 8L * k, table[i * BODY_DOUBLES + 1 + k]);
                /* A_SolidBody.rotate (A_SolidBody.java:65-71) calls Math.cos/sin per evaluation; they are not bit-specified, so
                 * both sides of the boundary consume these two doubles */
                buf.set(JAVA_DOUBLE, o + 48, Math.cos(theta));
                buf.set(JAVA_DOUBLE, o + 56, Math.sin(theta));
                for (int k = 0; k < 3; k++) buf.set(JAVA_DOUBLE, o + 64 + 8L * k, table[i * BODY_DOUBLES + 6 + k]);
            }
            check((int) SET_BODIES.invokeExact(handle, buf, n));
        } catch (Throwable e) { throw wrap(e); }
    }
    /** the reference's PPM frame, byte for byte, rendered from the device fields (toImage F1:341 / F6:1250 / F8:1417) */
    public void toImage(java.io.File destination, boolean renderHeat) {
        try (Arena a = Arena.ofConfined()) {
            check((int) WRITE_PPM.invokeExact(handle, a.allocateFrom(destination.getPath()), renderHeat ? 1 : 0));
        } catch (Throwable e) { throw wrap(e); }
    }
    /** copy a device field into a Java array (FIELD_D = _d._src, FIELD_U = _u._src, FIELD_V = _v._src ...) */
    public void readField(int field, double[] dst) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocate(JAVA_DOUBLE, dst.length);
            check((int) READ_FIELD.invokeExact(handle, field, buf, (long) dst.length));
            MemorySegment.copy(buf, JAVA_DOUBLE, 0, dst, 0, dst.length);
        } catch (Throwable e) { throw wrap(e); }
    }
    /** copy a Java array into a device field */
    public void writeField(int field, double[] src) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocateFrom(JAVA_DOUBLE, src);
            check((int) WRITE_FIELD.invokeExact(handle, field, buf, (long) src.length));
        } catch (Throwable e) { throw wrap(e); }
    }
    /** rank 0 of a solver group: the 128-byte id every rank passes to the constructor */
    public static byte[] commUniqueId() {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocate(128);
            check((int) COMM_UNIQUE_ID.invokeExact(buf));
            return buf.toArray(JAVA_BYTE);
        } catch (Throwable e) { throw wrap(e); }
    }
    /** {row0, nrows} of `field` this rank owns (everything on a single GPU) */
    public int[] slabRows(int field) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment r0 = a.allocate(JAVA_INT), nr = a.allocate(JAVA_INT);
            check((int) SLAB_ROWS.invokeExact(handle, field, r0, nr));
            return new int[] {r0.get(JAVA_INT, 0), nr.get(JAVA_INT, 0)};
        } catch (Throwable e) { throw wrap(e); }
    }
    /** rows [row0, row0 + nrows) of a device field; dst.length = nrows * row length */
    public void readFieldRows(int field, int row0, int nrows, double[] dst) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocate(JAVA_DOUBLE, dst.length);
            check((int) READ_FIELD_ROWS.invokeExact(handle, field, row0, nrows, buf, (long) dst.length));
            MemorySegment.copy(buf, JAVA_DOUBLE, 0, dst, 0, dst.length);
        } catch (Throwable e) { throw wrap(e); }
    }
    public void writeFieldRows(int field, int row0, int nrows, double[] src) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment buf = a.allocateFrom(JAVA_DOUBLE, src);
            check((int) WRITE_FIELD_ROWS.invokeExact(handle, field, row0, nrows, buf, (long) src.length));
        } catch (Throwable e) { throw wrap(e); }
    }
    /** F8: ParticleQuantities._particleCount */
    public int particleCount() {
        try {
            MemorySegment out = arena.allocate(JAVA_INT);
            check((int) PARTICLE_COUNT.invokeExact(handle, out));
            return out.get(JAVA_INT, 0);
        } catch (Throwable e) { throw wrap(e); }
    }
    @Override public void close() {
        try { DESTROY.invokeExact(handle); } catch (Throwable e) { throw wrap(e); } finally { arena.close(); }
    }
    private static void check(int rc) throws Throwable {
        if (rc != 0) {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512);
            throw new IllegalStateException("ifl status " + rc + ": " + msg.getString(0));
        }
    }
    private static RuntimeException wrap(Throwable t) { return t instanceof RuntimeException r ? r : new RuntimeException(t); }
}
