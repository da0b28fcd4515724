"""Short run of one BASELINE workload for ncu captures: python tools/prof_target.py c5|c1|c2|c3|c4 [steps] [solver_limit]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import incfluid_b200 as ifl
name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[name]
kw = {"solver_limit": int(sys.argv[3])} if len(sys.argv) > 3 else {}
g = bench.make_gpu_solver(ifl, wl, **kw)
for _ in range(steps):
    st = bench.one_step(g, wl)
print(name, st, g.launch_count(), flush=True)
