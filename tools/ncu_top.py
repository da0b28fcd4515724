"""Top stall sites of an ncu report (source page): python tools/ncu_top.py report.ncu-rep [kernel-substring] [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print({k.split(".")[0][-28:]: r[hdr.index(k)][:48] for k in keys if k in hdr})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = set()
for si, start in enumerate(secs):
    name = rows[start][1]
    if want not in name or name in seen: continue
    seen.add(name)
    hdr = rows[start + 1]; end = secs[si + 1] if si + 1 < len(secs) else len(rows)
    body = [r for r in rows[start + 2:end] if len(r) == len(hdr)]
    cs, cx, csrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[cs] or 0) for r in body)
    agg = {h[6:]: sum(int(r[i] or 0) for r in body) for i, h in stall}
    print("\n####", name[:90], "| instructions", len(body), "| samples", tot)
    print("   ", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]})
    for r in sorted(body, key=lambda r: -int(r[cs] or 0))[:N]:
        rs = sorted(((h[6:], int(r[i] or 0)) for i, h in stall if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{int(r[cs]):6d} {100 * int(r[cs]) / max(tot, 1):5.1f}%  ex {r[cx]:>9}  {r[csrc].strip()[:56]:56s} {rs}")
