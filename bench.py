#!/usr/bin/env python
"""bench.py -- cell-updates/s of the incremental-fluids timestep hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Workload (all N): BASELINE.json configs[4] -- the one its metric "(1/2/4/8 B200)" is quoted on, and it fits one GPU --
RK3 + Catmull-Rom advection with MIC(0)-PCG projection (ExplicitFluid3conjugateGradients.java) on a 16384x16384 grid,
scene constants of F3:63-76, PCG limit 600 / tolerance 1e-5 (the solve runs to the cap at this size).
One "step" = the body of the reference's main() loop: addInflow + FluidSolver.update(dt) (F3:75-79).  At N > 1 the
pressure solve is slab-decomposed over the ranks (strong scaling: the total work is fixed).  `--workload c2` selects
configs[1] (1024x1024) instead; at N = 1 its numbers are also attached under "configs1_1024".

value   : cells * K / device time (CUDA events, max over ranks), fields resident in HBM.
e2e     : same metric through the reference-facing call with HOST state: every step uploads d/u/v from pinned host
          buffers, runs addInflow+update and reads d/u/v back, all inside the timed region.
roofline: dominant kernel = applyPreconditioner forward sweep (F3:496-507); achieved = 40 B/cell * cells /
          (average launch time from live CUDA-event samples taken during the timed steps).
cpu_baseline / --impl reference: the CPU oracle (oracle/ifl_oracle.cpp, a line-by-line C++ restatement of the Java
          solver; no JVM exists in this image), single-threaded like the reference, on a bounded sample (see `sample`).
"""
import argparse
import json
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

DT = 0.005
DENSITY = 0.1
INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 0.0, 3.0)     # F3:76
LIMIT = 600                                              # F3:365
WORKLOADS = {
    "c5": (16384, "configs[4]: F3 MIC(0)-PCG projection + RK3/Catmull-Rom advection, 16384x16384, inflow scene F3:63-76, "
                  "slab-decomposed over the ranks"),
    "c2": (1024, "configs[1]: F3 RK3+Catmull-Rom advection + MIC(0)-PCG projection, 1024x1024, inflow scene F3:63-76"),
}
METRIC = "cell-updates/sec per full timestep"
UNIT = "cell-updates/s"
A_SWEEP_BYTES_PER_CELL = 40       # R a,aPlusX,aPlusY,precon -> W dst
A_BWD_BYTES_PER_CELL = 48         # backward sweep (a14, 40 B) fused with dotProduct(z, r) (a16: the 8 B of r; z is in registers)
# algorithmic bytes per cell and launch of the five kernels of a PCG iteration
# matvec_dot: F3's matrix is re-derived from the grid position (R s -> W z); update_s carries the deferred p += alpha*s
A_KERNEL_BYTES = {"matvec_dot": 16, "update_pr": 24, "precon_fwd": 40, "precon_bwd": 48, "update_s": 40}
CPU_SAMPLE_GRID = 1024
CPU_SAMPLE_LIMITS = (40, 120)


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


def cpu_sample(limit_total=LIMIT):
    """Bounded CPU sample: two first steps of the F3 scene at CPU_SAMPLE_GRID^2 with the PCG capped at 40 and at 120
    iterations give the cost per cell-iteration and the per-step remainder; cells/s of a full step with `limit_total`
    iterations follows.  (A real 16384^2 step is ~80 CPU-minutes.)  The small grid is cache resident on the host, which
    flatters the CPU relative to a DRAM-bound 16384^2 run."""
    import incfluid_b200 as ifl
    from oracle_harness import OracleSolver
    abi = ifl.abi
    n = CPU_SAMPLE_GRID
    times, its = [], []
    for lim in CPU_SAMPLE_LIMITS:
        o = OracleSolver(abi, abi.default_config(abi.F3_CONJUGATE_GRADIENTS, n, n, density=DENSITY, solver_limit=lim))
        t0 = time.perf_counter()
        o.add_inflow(*INFLOW)
        st = o.update(DT)
        times.append(time.perf_counter() - t0)
        its.append(st["iterations_pressure"])
        o.close()
    c_iter = (times[1] - times[0]) / max(its[1] - its[0], 1) / (n * n)
    c_rest = times[0] / (n * n) - c_iter * its[0]
    per_cell_step = c_rest + c_iter * limit_total
    sample = (f"oracle, 1 thread: first step of the F3 scene at {n}x{n} with the PCG capped at {its[0]} and {its[1]} iterations "
              f"({times[0]:.1f} s + {times[1]:.1f} s) -> {c_iter*1e9:.1f} ns per cell-iteration + {c_rest*1e9:.0f} ns per cell-step; "
              f"extrapolated to a full step with {limit_total} iterations (what the 16384x16384 scene runs)")
    return 1.0 / per_cell_step, sample, sum(times)


def run_reference(args, grid, workload):
    """Reference arm: the reference algorithm on the host CPU (single thread, like the Java program)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    vals, wall = [], 0.0
    sample = ""
    for i in range(max(args.warmup, 0) + args.steps):
        v, sample, dt = cpu_sample()
        if i >= args.warmup:
            vals.append(v); wall += dt
        if wall > 200.0:
            break
    value = statistics.mean(vals)
    k = len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * grid * grid / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "grid": [grid, grid], "dt": DT, "solver_limit": LIMIT, "tolerance": 1e-5},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample + f"; mean of {k} samples"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is single-threaded Java; no JVM in this image, so its C++ oracle port is timed (1 core); "
                "ms_per_step is the extrapolated time of one full step of the named workload",
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    grid, workload = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, grid, workload)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import incfluid_b200 as ifl

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

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

    abi = ifl.abi
    W = H = grid
    inflow7 = INFLOW[:5] + INFLOW[6:]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def make_solver():
        return ifl.FluidSolver(W, H, DENSITY, variant=abi.F3_CONJUGATE_GRADIENTS, device=local_rank, **comm_kw())

    # ------------------------------------------------------------- resident arm
    g = make_solver()
    for _ in range(args.warmup):
        g.addInflow(*inflow7); g.update(DT, want_status=False)
    g.synchronize()
    g.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    its = []
    for _ in range(args.steps):
        flush.zero_()
        g.addInflow(*inflow7)
        its.append(g.update(DT)["iterations_pressure"])      # status read = the stream sync of this step
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = g.launch_count() - launches0 + args.steps      # + one flush kernel per step (torch)
    value = W * H * args.steps / (ms * 1e-3)                  # strong scaling: the ranks share ONE grid
    sweep_ms, nsamp = g.kernel_timing("precon_bwd")
    kt = {k: g.kernel_timing(k)[0] for k in ("matvec_dot", "update_pr", "precon_fwd", "precon_bwd", "update_s")}
    phase = g.last_timing()
    g.set_profiling(False)
    g.close()
    del g

    # ------------------------------------------------------------- end-to-end arm (host state)
    e2e = None
    if not args.no_e2e:
        nd, nu, nv = W * H, (W + 1) * H, W * (H + 1)
        hd = torch.zeros(nd, dtype=torch.float64).pin_memory()
        hu = torch.zeros(nu, dtype=torch.float64).pin_memory()
        hv = torch.zeros(nv, dtype=torch.float64).pin_memory()
        nhd, nhu, nhv = hd.numpy(), hu.numpy(), hv.numpy()
        g2 = make_solver()

        def e2e_step():
            g2.write_field("d", nhd); g2.write_field("u", nhu); g2.write_field("v", nhv)      # H2D from pinned
            g2.addInflow(*inflow7)
            g2.update(DT, want_status=False)
            g2.read_field("d", nhd); g2.read_field("u", nhu); g2.read_field("v", nhv)         # D2H result

        k_e2e = max(1, min(args.steps, 2 if grid > 4096 else args.steps))
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
        e2e = {"value": W * H * k_e2e / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": 8 * (nd + nu + nv) * world, "d2h_bytes_per_step": 8 * (nd + nu + nv) * world,
               "ms_per_step": e2e_ms / k_e2e, "steps": k_e2e,
               "note": "every rank uploads and downloads the full d/u/v (state outside the solve is replicated per rank)"}
        g2.close()
        del g2, hd, hu, hv

    # ------------------------------------------------------------- configs[1] for reference (N = 1 only)
    extra = None
    if world == 1 and args.workload == "c5":
        n1 = 1024
        s1 = ifl.FluidSolver(n1, n1, DENSITY, variant=abi.F3_CONJUGATE_GRADIENTS, device=local_rank)
        for _ in range(3):
            s1.addInflow(*inflow7); s1.update(DT, want_status=False)
        s1.synchronize(); s1.set_profiling(True)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        its1 = []
        for _ in range(3):
            flush.zero_(); s1.addInflow(*inflow7); its1.append(s1.update(DT)["iterations_pressure"])
        a1.record(); torch.cuda.synchronize()
        t1 = a0.elapsed_time(a1) / 3
        fw = s1.kernel_timing("precon_fwd")[0]
        extra = {"workload": WORKLOADS["c2"][1], "value": n1 * n1 / (t1 * 1e-3), "unit": UNIT, "ms_per_step": t1,
                 "pcg_iterations_per_step": its1, "precon_fwd_ms": fw,
                 "roofline_frac_precon_fwd": (A_SWEEP_BYTES_PER_CELL * n1 * n1 / (fw * 1e-3) / 1e9 / measured_hbm_peak()[0]) if fw > 0 else None}
        s1.close()

    peak, peak_src = measured_hbm_peak()
    achieved = A_BWD_BYTES_PER_CELL * W * H / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"precon_bwd_dram_bytes_per_launch_{grid}")
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "grid": [W, H], "dt": DT, "solver_limit": LIMIT, "tolerance": 1e-5,
                       "pcg_iterations_per_step": its,
                       "l2": "flushed between steps (256 MiB device memset inside the timed region); the PCG working set "
                             f"of this grid is {8 * 8 * W * H / 2**30:.1f} GiB (L2 is 126 MB)",
                       "multi_gpu": "single GPU" if world == 1 else
                       f"pressure solve slab-decomposed over {world} ranks (strip hand-over by peer stores over NVLink, "
                       "row halos by ncclSend/Recv, scalars by ncclAllGather); advection and the once-per-step kernels "
                       "are computed redundantly on every rank"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_precon_sweep<-1,dot> (applyPreconditioner backward sweep F3:509-520 fused "
                                                   "with dotProduct(z,r) F3:625; the longest kernel of an iteration)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": A_BWD_BYTES_PER_CELL * W * H // world,
                         "avg_launch_ms": sweep_ms, "samples": nsamp,
                         "note": "a lexicographic recurrence: at most one anti-diagonal is in flight and the critical path is "
                                 "w + h dependent steps, so the kernel is latency bound below ~16k^2 (DESIGN.md 5)"},
            "kernel_ms": kt,
            "kernel_gbs": {k: (A_KERNEL_BYTES[k] * W * H / world / (v * 1e-3) / 1e9 if v > 0 else None) for k, v in kt.items()},
            "kernel_note": "algorithmic bytes per cell: matvec_dot 16 (F3 matrix re-derived per cell: R s, W z), update_pr 24 "
                           "(r -= alpha*z + infinityNorm), sweeps 40 / 48, update_s 40 (k_update_sp: p += alpha*s and "
                           "s = z + beta*s in one pass)",
            "phase_ms_last_step": phase,
        }
        if world > 1:
            line["roofline"]["achieved"] = achieved / world
            line["roofline"]["frac"] = achieved / world / peak
            line["roofline"]["note"] += "; per-rank bytes (each rank sweeps its own strips)"
        if e2e is not None:
            line["e2e"] = e2e
        if extra is not None:
            line["configs1_1024"] = extra
        if not args.no_cpu_baseline and world == 1:
            v, sample, _ = cpu_sample()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
