#!/usr/bin/env python
"""bench.py -- cell-updates/s of the incremental-fluids timestep hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU
  python bench.py --workload c1|c2|c3|c4|c5                # any BASELINE.json config as the main workload

Headline workload (all N): BASELINE.json configs[4] -- the one its metric "(1/2/4/8 B200)" is quoted on, and it fits one
GPU -- RK3 + Catmull-Rom advection with MIC(0)-PCG projection (ExplicitFluid3conjugateGradients.java) on a 16384x16384
grid, scene constants of F3:63-76, PCG limit 600 / tolerance 1e-5 (the solve runs to the cap at this size).
One "step" = the body of the reference's main() loop: addInflow + FluidSolver.update(dt) (F3:75-79).  At N > 1 the
grid is slab-decomposed over the ranks (strong scaling: the total work is fixed).  At N = 1 the default run also
attaches the other four configs as bounded sub-records under "other_configs" (configs[0] F1 128^2, configs[1] F3
1024^2, configs[2] F7 4096^2, configs[3] F8 2048^2 with 8 particles per cell), each with its own CPU sample.

value   : cells * K / device time (CUDA events, max over ranks), fields resident in HBM.
e2e     : same metric through the reference-facing call with HOST state: every step uploads the fields a Java caller
          owns (d, u, v [, t] [, particles]) from pinned host buffers, runs addInflow+update and reads them back, all
          inside the timed region.  At N > 1 every rank moves only the rows of its slab.
roofline: the longest kernel of a PCG iteration; achieved = algorithmic bytes / average launch time from live
          CUDA-event samples taken during the timed steps.
cpu_baseline / --impl reference: the CPU oracle (oracle/ifl_oracle.cpp, a line-by-line C++ restatement of the Java
          solver; no JVM exists in this image), single-threaded like the reference, on a bounded sample (see `sample`).
state_hash: 64-bit hash of the bits of p, d, u, v after the timed steps (row-wise, combined independently of the slab
          decomposition): equal values on the N = 1, 2, 4, 8 lines prove the result does not depend on the GPU count.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"       # NCCL prints its version banner on stdout: keep stdout to the one JSON line
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cell-updates/sec per full timestep"
UNIT = "cell-updates/s"
PI = math.pi

# (x, y, w, h, d, t, u, v) rectangles of the reference's main() programs
F1_INFLOW = (0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 0.0, 3.0)      # F1:76
F3_INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)     # F3:76
F7_INFLOW = (0.45, 0.2, 0.1, 0.05, 1.0, 294.0, 0.0, 0.0)    # F7:92
F7_BODY = ("box", (0.5, 0.6, 0.7, 0.1, PI * 0.25, 0.0, 0.0, 0.0))   # F7:83 (= F8:83)

WORKLOADS = {
    "c1": dict(variant="F1_MATRIXLESS", grid=128, dt=0.005, limit=600, inflow=F1_INFLOW, fields=("d", "u", "v"),
               text="configs[0]: F1 Euler + linear advection + Gauss-Seidel projection, 128x128 inflow scene F1:60-78"),
    "c2": dict(variant="F3_CONJUGATE_GRADIENTS", grid=1024, dt=0.005, limit=600, inflow=F3_INFLOW, fields=("d", "u", "v"),
               text="configs[1]: F3 RK3+Catmull-Rom advection + MIC(0)-PCG projection, 1024x1024, inflow scene F3:63-76"),
    "c3": dict(variant="F7_VARIABLE_DENSITY", grid=4096, dt=0.005, limit=2000, inflow=F7_INFLOW, body=F7_BODY, soot=1.0,
               fields=("d", "t", "u", "v"),
               text="configs[2]: F7 solid boundaries + heat + variable density, 4096x4096, scene F7:71-92"),
    "c4": dict(variant="F8_FLIP", grid=2048, dt=0.0025, limit=2000, inflow=None, body=F7_BODY, soot=0.25, ppc=8,
               fields=("d", "t", "u", "v"), particles=True,
               text="configs[3]: F8 FLIP/PIC, 2048x2048 grid, 8 particles per cell, scene F8:72-86 + F8:1330"),
    "c5": dict(variant="F3_CONJUGATE_GRADIENTS", grid=16384, dt=0.005, limit=600, inflow=F3_INFLOW, fields=("d", "u", "v"),
               text="configs[4]: F3 MIC(0)-PCG projection + RK3/Catmull-Rom advection, 16384x16384, inflow scene F3:63-76, "
                    "slab-decomposed over the ranks"),
}
# algorithmic bytes per cell and launch (F3: the matrix is re-derived from the grid position)
#   step_a    : R z,s,p  W s,p,z                         (scaledAdd x2 + matrixVectorProduct + dotProduct)
#   precon_fwd: R r,z,aPlusX*precon,aPlusY*precon,precon  W r,z   (scaledAdd(r) + infinityNorm + forward sweep)
#   precon_bwd: R z,3 coefficients,r  W z                 (backward sweep + dotProduct)
A_KERNEL_BYTES = {"step_a": 48, "precon_fwd": 56, "precon_bwd": 48, "gs_sweep": 24}
A_KERNEL_BYTES_MASKED = {"step_a": 48 + 24 + 1, "precon_fwd": 56, "precon_bwd": 48, "gs_sweep": 24}   # + aDiag,aPlusX,aPlusY,cell


class quiet_stdout:
    """NCCL prints its version banner on stdout the first time a communicator is made (whatever NCCL_DEBUG says in some
    builds): route fd 1 to stderr meanwhile, so that stdout carries nothing but the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/ifl_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- solver construction
def make_bodies(ifl, wl):
    if "body" not in wl:
        return None
    kind, args = wl["body"]
    return [(ifl.SolidBox if kind == "box" else ifl.SolidSphere)(*args)]


def make_gpu_solver(ifl, wl, device=0, grid=None, **kw):
    abi = ifl.abi
    n = grid or wl["grid"]
    w, h = (n, n) if isinstance(n, int) else n
    variant = getattr(abi, wl["variant"])
    bodies = make_bodies(ifl, wl)
    if "ppc" in wl:
        kw.setdefault("particles_avg_per_cell", wl["ppc"])
    kw.setdefault("solver_limit", wl["limit"])
    if variant >= abi.F6_HEAT:
        return ifl.FluidSolver(w, h, 0.1, wl["soot"], 0.01, bodies, variant=variant, device=device, **kw)
    if variant >= abi.F4_SOLID_BOUNDARIES:
        return ifl.FluidSolver(w, h, 0.1, bodies, variant=variant, device=device, **kw)
    return ifl.FluidSolver(w, h, 0.1, variant=variant, device=device, **kw)


def make_cpu_solver(ifl, wl, grid, **kw):
    """the reference algorithm on the host: the CPU oracle (test infrastructure, used here only as the timed baseline)"""
    from oracle_harness import OracleSolver
    abi = ifl.abi
    w, h = (grid, grid) if isinstance(grid, int) else grid
    variant = getattr(abi, wl["variant"])
    cfg_kw = dict(kw)
    cfg_kw.setdefault("solver_limit", wl["limit"])
    if "soot" in wl:
        cfg_kw["density_soot"] = wl["soot"]
    if "ppc" in wl:
        cfg_kw["particles_avg_per_cell"] = wl["ppc"]
    o = OracleSolver(abi, abi.default_config(variant, w, h, **cfg_kw))
    bodies = make_bodies(ifl, wl)
    if bodies is not None:
        o.set_bodies([b.as_abi() for b in bodies])
    return o


def one_step(s, wl):
    if wl["inflow"] is not None:
        s.add_inflow(*wl["inflow"])
    return s.update(wl["dt"])


# ------------------------------------------------------------------------------------------------------ CPU samples
def cpu_sample(ifl, name, its_per_step=None):
    """Bounded single-thread sample of the reference algorithm for workload `name`.  Returns (cells/s, sample text, seconds).

    c1, c2: full steps of the workload itself, measured.  c3, c4, c5: a full step takes CPU-hours, so two steps of the
    same variant with the solver capped at two different iteration counts give the cost per cell-iteration and the
    per-step remainder, extrapolated to the iteration counts the GPU run of the full workload measured."""
    wl = WORKLOADS[name]
    t_all = time.perf_counter()
    if name in ("c1", "c2"):
        n = wl["grid"]
        o = make_cpu_solver(ifl, wl, n)
        steps = 30 if name == "c1" else 1
        t0 = time.perf_counter()
        its = [one_step(o, wl)["iterations_pressure"] for _ in range(steps)]
        dt = time.perf_counter() - t0
        o.close()
        v = n * n * steps / dt
        return v, (f"oracle, 1 thread: the first {steps} step(s) of this workload measured in full ({dt:.1f} s, "
                   f"{its[0]}..{its[-1]} solver iterations per step)"), time.perf_counter() - t_all
    # capped-iteration samples: (grid, (limit_lo, limit_hi))
    grid, (lo, hi) = {"c3": (1024, (4, 12)), "c4": (1024, (4, 12)), "c5": ((16384, 1024), (2, 6))}[name]
    cells = grid * grid if isinstance(grid, int) else grid[0] * grid[1]
    times, iters = [], []
    for lim in (lo, hi):
        o = make_cpu_solver(ifl, wl, grid, solver_limit=lim)
        if wl.get("particles"):
            o.particle_count()                      # the constructor's initParticles is not part of a step
        t0 = time.perf_counter()
        st = one_step(o, wl)
        times.append(time.perf_counter() - t0)
        iters.append(st["iterations_pressure"] + st["iterations_heat"])
        o.close()
    c_iter = (times[1] - times[0]) / max(iters[1] - iters[0], 1) / cells
    c_rest = max(times[0] / cells - c_iter * iters[0], 0.0)
    n_it = its_per_step if its_per_step else wl["limit"]
    per_cell_step = c_rest + c_iter * n_it
    gtxt = f"{grid}x{grid}" if isinstance(grid, int) else f"{grid[0]}x{grid[1]} (a slab of the workload's rows: same row length, DRAM resident)"
    sample = (f"oracle, 1 thread: first step of the {wl['variant']} scene at {gtxt} with the solver(s) capped at {lo} and {hi} "
              f"iterations ({times[0]:.1f} s + {times[1]:.1f} s, {iters[0]} and {iters[1]} iterations in total) -> "
              f"{c_iter * 1e9:.1f} ns per cell-iteration + {c_rest * 1e9:.0f} ns per cell-step; EXTRAPOLATED to a step with "
              f"{n_it:.0f} solver iterations (what the GPU run of the full workload executed)")
    return 1.0 / per_cell_step, sample, time.perf_counter() - t_all


def run_reference(args, name):
    """Reference arm: the reference algorithm on the host CPU (single thread, like the Java program)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import incfluid_b200 as ifl
    wl = WORKLOADS[name]
    vals, wall, sample = [], 0.0, ""
    for i in range(max(args.warmup, 0) + args.steps):
        v, sample, dt = cpu_sample(ifl, name)
        if i >= args.warmup or wall + dt > 150.0:
            vals.append(v)
        wall += dt
        if wall > 150.0:
            break
    value = statistics.mean(vals)
    k = len(vals)
    n = wl["grid"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * n * n / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["text"], "grid": [n, n], "dt": wl["dt"], "solver_limit": wl["limit"], "tolerance": 1e-5},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample + f"; mean of {k} sample(s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is single-threaded Java; no JVM in this image, so its C++ oracle port is timed (1 core); "
                "`steps` counts bounded samples, and ms_per_step is the time one full step of the named workload would "
                "take at the sampled cost (see cpu_baseline.sample) -- a stated baseline, not a same-size measurement",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------- GPU runs
def state_hash(np, arrays_with_row0):
    """64-bit hash of field bits that does not depend on how rows are split over ranks: every row is reduced with
    per-column odd multipliers, weighted by a multiplier derived from its GLOBAL row index, and the rows are summed
    modulo 2^64.  arrays_with_row0: iterable of (2-D float64 array of some rows, global index of its first row, field id)."""
    total = np.uint64(0)
    with np.errstate(over="ignore"):
        for arr, row0, fid in arrays_with_row0:
            bits = arr.view(np.uint64)
            ncol = bits.shape[1]
            col = (np.arange(ncol, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD1B54A32D192ED03)) | np.uint64(1)
            rows = (bits * col).sum(axis=1, dtype=np.uint64)
            ridx = np.arange(row0, row0 + bits.shape[0], dtype=np.uint64) + np.uint64(1 + 1000003 * fid)
            rmul = (ridx * np.uint64(0xBF58476D1CE4E5B9)) | np.uint64(1)
            rows ^= rows >> np.uint64(29)
            total = total + (rows * rmul).sum(dtype=np.uint64)
    return int(total)


def timed_gpu_run(torch, ifl, wl, steps, warmup, local_rank, flush, barrier, max_over_ranks, comm_kw, profile=True):
    """warm-up, then `steps` timed steps with resident state; returns a dict of measurements and the live solver"""
    with quiet_stdout():
        g = make_gpu_solver(ifl, wl, device=local_rank, **comm_kw())
    if wl.get("particles"):
        g.particle_count()                        # F8: initParticles (constructor work) outside the timed region
    for _ in range(warmup):
        if wl["inflow"] is not None:
            g.add_inflow(*wl["inflow"])
        g.update(wl["dt"], want_status=False)
    g.synchronize()
    if profile:
        g.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    its, its_heat = [], []
    for _ in range(steps):
        flush.zero_()
        st = one_step(g, wl)                      # status read = the stream sync of this step
        its.append(st["iterations_pressure"]); its_heat.append(st["iterations_heat"])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out = {"ms": ms, "its": its, "its_heat": its_heat, "clocks": clocks,
           "launches": int(g.launch_count() - launches0 + steps),      # + one flush kernel per step (torch)
           "phase": g.last_timing()}
    if profile:
        names = ("gs_sweep",) if wl["variant"] in ("F1_MATRIXLESS", "F2_BETTER_ADVECTION") else ("step_a", "precon_fwd", "precon_bwd")
        out["kernel_ms"] = {k: g.kernel_timing(k) for k in names}
        g.set_profiling(False)
    return out, g


def sub_record(torch, ifl, name, steps, warmup, local_rank, flush, with_cpu=True):
    """one of the other BASELINE configs as a bounded sub-record of the N = 1 line"""
    wl = WORKLOADS[name]
    ident = lambda x: x
    r, g = timed_gpu_run(torch, ifl, wl, steps, warmup, local_rank, flush, torch.cuda.synchronize, ident, dict)
    g.close()
    n = wl["grid"]
    peak = measured_hbm_peak()[0]
    masked = "body" in wl
    ab = A_KERNEL_BYTES_MASKED if masked else A_KERNEL_BYTES
    kt = {k: v[0] for k, v in r["kernel_ms"].items()}
    rec = {"workload": wl["text"], "value": n * n * steps / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / steps,
           "steps": steps, "warmup": warmup, "solver_iterations_per_step": r["its"],
           "kernel_ms": kt,
           "kernel_frac_of_hbm_peak": {k: (ab[k] * n * n / (v * 1e-3) / 1e9 / peak if v > 0 else None) for k, v in kt.items()},
           "phase_ms_last_step": r["phase"], "gpu_launches": r["launches"]}
    if r["its_heat"] and any(r["its_heat"]):
        rec["heat_iterations_per_step"] = r["its_heat"]
    if with_cpu:
        per_step = statistics.mean(a + b for a, b in zip(r["its"], r["its_heat"]))
        v, sample, _ = cpu_sample(ifl, name, its_per_step=per_step)
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-hash", action="store_true")
    args = ap.parse_args()
    name = args.workload
    wl = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference(args, name)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import incfluid_b200 as ifl

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1 and name not in ("c2", "c3", "c5"):
        raise SystemExit("several GPUs: the MIC(0)-PCG workloads c2, c5 (F3: slab decomposition) and c3 (F7: slab-decomposed "
                         "solver loops, every other stage repeated by every rank); c1 (Gauss-Seidel) and c4 (particles) run on one GPU")
    # F4-F7 on several GPUs: every rank holds the whole state, so a host that owns the state writes whole fields on every rank
    replicated = world > 1 and name == "c3"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    def comm_kw():
        """every slab-decomposed solver group needs its own id (an ncclUniqueId), made by rank 0 and broadcast"""
        if world == 1:
            return {}
        ids = [ifl.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return dict(rank=rank, world_size=world, comm_id=ids[0])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = H = wl["grid"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    s0, s1, row0, row1 = ifl.slab_partition(H, world)[rank]

    # ------------------------------------------------------------- resident arm
    r, g = timed_gpu_run(torch, ifl, wl, args.steps, args.warmup, local_rank, flush, barrier, max_over_ranks, comm_kw)
    ms, its = r["ms"], r["its"]
    value = W * H * args.steps / (ms * 1e-3)                  # strong scaling: the ranks share ONE grid
    kt = {k: v[0] for k, v in r["kernel_ms"].items()}
    ksamples = {k: v[1] for k, v in r["kernel_ms"].items()}

    # 64-bit hash of the state (outside the timed region): own rows on every rank, combined across ranks
    shash = None
    if not args.no_hash:
        parts = []
        for fid, fname in enumerate(("p", "d", "u", "v")):
            width = W + 1 if fname == "u" else W
            nrows = (H + 1 if fname == "v" else H)
            if world == 1:
                a = g.read_field(fname).reshape(nrows, width)
                parts.append((a, 0, fid))
            else:
                hi = row1 + (1 if (fname == "v" and rank == world - 1) else 0)
                a = g.read_field_rows(fname, row0, hi - row0).reshape(hi - row0, width)
                parts.append((a, row0, fid))
        hv = state_hash(np, parts)
        del parts
        if dist is not None:
            hs = [None] * world
            dist.all_gather_object(hs, hv)
            hv = sum(hs) % (1 << 64)
        shash = f"{hv:016x}"
    mem_free, mem_total = torch.cuda.mem_get_info()
    g.close()
    del g

    # ------------------------------------------------------------- end-to-end arm (host state)
    e2e = None
    if not args.no_e2e:
        with quiet_stdout():
            g2 = make_gpu_solver(ifl, wl, device=local_rank, **comm_kw())
        shapes = {"d": (H, W), "t": (H, W), "u": (H, W + 1), "v": (H + 1, W)}
        host = {}
        for f in wl["fields"]:
            nr, nc = shapes[f]
            lo, hi = (0, nr) if (world == 1 or replicated) else (row0, row1 + (1 if (f == "v" and rank == world - 1) else 0))
            host[f] = (torch.zeros((hi - lo) * nc, dtype=torch.float64).pin_memory(), lo, hi - lo)
            if f == "t":
                host[f][0].fill_(294.0)
        hp = None
        if wl.get("particles"):
            hp = g2.read_particles()
            hp = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in hp.items()}
        nbytes = sum(t.numel() * 8 for t, _, _ in host.values()) + (sum(v.nbytes for v in hp.values()) if hp else 0)

        def e2e_step():
            nonlocal hp
            for f, (t, lo, nr) in host.items():                       # H2D from pinned host memory
                if world == 1 or replicated:
                    g2.write_field(f, t.numpy())
                else:
                    g2.write_field_rows(f, lo, nr, t.numpy())
            if hp is not None:
                g2.write_particles(hp)
            if wl["inflow"] is not None:
                g2.add_inflow(*wl["inflow"])
            g2.update(wl["dt"], want_status=False)
            for f, (t, lo, nr) in host.items():                       # D2H result
                if world == 1 or replicated:
                    g2.read_field(f, t.numpy())
                else:
                    g2.read_field_rows(f, lo, nr, t.numpy())
            if hp is not None:
                hp = g2.read_particles()

        k_e2e = max(1, min(args.steps, 2 if W > 4096 else args.steps))
        e2e_step()                                                                             # warm-up
        barrier()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k_e2e):
            flush.zero_()
            e2e_step()
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), wall_ms))
        tot = nbytes
        if dist is not None:
            tb = torch.tensor([nbytes], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb)
            tot = int(tb.item())
        e2e = {"value": W * H * k_e2e / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": tot, "d2h_bytes_per_step": tot,
               "ms_per_step": e2e_ms / k_e2e, "steps": k_e2e,
               "note": "host-owned state (" + ", ".join(wl["fields"]) + (", particles" if hp is not None else "") +
                       ") uploaded from pinned memory and read back every step; bytes are summed over the ranks "
                       "(every rank moves only the rows of its slab)"}
        g2.close()
        del g2, host

    # ------------------------------------------------------------- the other BASELINE configs (N = 1 only)
    others = None
    if world == 1 and name == "c5" and not args.no_other_configs:
        others = {}
        cpu = not args.no_cpu_baseline
        others["c1"] = sub_record(torch, ifl, "c1", 40, 5, local_rank, flush, cpu)
        others["c2"] = sub_record(torch, ifl, "c2", 3, 3, local_rank, flush, cpu)
        others["c3"] = sub_record(torch, ifl, "c3", 2, 1, local_rank, flush, cpu)
        others["c4"] = sub_record(torch, ifl, "c4", 2, 1, local_rank, flush, cpu)

    peak, peak_src = measured_hbm_peak()
    masked = "body" in wl
    ab = A_KERNEL_BYTES_MASKED if masked else A_KERNEL_BYTES
    dom = max(kt, key=lambda k: kt[k]) if kt else None
    dom_ms = kt.get(dom, 0.0) if dom else 0.0
    cells_rank = W * (row1 - row0)
    achieved = ab[dom] * cells_rank / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if world == 1 and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{dom}_dram_bytes_per_launch_{name}")
        except Exception:
            traffic = None
    dom_text = {"step_a": "k_step_a (scaledAdd p, s + matrixVectorProduct + dotProduct, F3:607,626,604,605)",
                "precon_fwd": "k_precon_sweep<+1,fused> (scaledAdd(r) F3:608 + infinityNorm F3:609 + applyPreconditioner forward sweep F3:496-507)",
                "precon_bwd": "k_precon_sweep<-1,dot> (applyPreconditioner backward sweep F3:509-520 + dotProduct(z,r) F3:625)",
                "gs_sweep": "k_gs_sweep (one in-place lexicographic Gauss-Seidel sweep, F1:383-417)"}

    if rank == 0:
        n_it = statistics.mean(its) if its else 0
        t_chain = (kt.get("precon_fwd", 0.0) + kt.get("precon_bwd", 0.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"], "grid": [W, H], "dt": wl["dt"], "solver_limit": wl["limit"], "tolerance": 1e-5,
                       "solver_iterations_per_step": its,
                       "l2": "flushed between steps (256 MiB device memset inside the timed region); the PCG working set "
                             f"of this grid is {10 * 8 * W * H / 2**30:.1f} GiB (L2 is 126 MB)",
                       "state_hash": shash,
                       "multi_gpu": "single GPU" if world == 1 else
                       (f"solver loops (heat and pressure MIC(0)-PCG) slab-decomposed over {world} ranks; every rank holds the "
                        "whole state and repeats the other stages; the strips of p are gathered after a solve") if replicated else
                       f"grid slab-decomposed over {world} ranks: every rank computes only its rows (advection with a wide "
                       "halo sized from the CFL reduction, RHS, PCG, applyPressure); strip hand-over of the triangular "
                       "sweeps by peer stores over NVLink, row halos by ncclSend/Recv, scalars by grouped ncclBroadcast"},
            "gpu_launches": r["launches"],
            "clocks": r["clocks"],
            "roofline": {"bound": "hbm", "kernel": dom_text.get(dom, dom),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ab[dom] * cells_rank if dom else None,
                         "avg_launch_ms": dom_ms, "samples": ksamples.get(dom),
                         "note": "the longest kernel of a solver iteration, live CUDA-event samples; the triangular sweeps are "
                                 "lexicographic recurrences whose critical path is w + h dependent cell steps (DESIGN.md 5)"
                                 + ("; per-rank bytes (each rank sweeps its own strips)" if world > 1 else "")},
            "kernel_ms": kt,
            "kernel_gbs": {k: (ab[k] * cells_rank / (v * 1e-3) / 1e9 if v > 0 else None) for k, v in kt.items()},
            "kernel_note": "algorithmic bytes per cell and launch: " + ", ".join(f"{k} {ab[k]}" for k in kt),
            "iteration_model": {"t_chain_ms": t_chain, "t_elementwise_ms": kt.get("step_a", 0.0),
                                "note": "one PCG iteration = step_a + precon_fwd + precon_bwd; the two sweeps are one dependency "
                                        "chain through all ranks (T_chain does not shrink with N), step_a scales as 1/N"},
            "phase_ms_last_step": r["phase"],
            "device_memory_used_gib": (mem_total - mem_free) / 2**30,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if others is not None:
            line["other_configs"] = others
        if not args.no_cpu_baseline and world == 1:
            v, sample, _ = cpu_sample(ifl, name, its_per_step=n_it)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
