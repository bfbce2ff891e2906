"""Per-phase clock shares of the planner kernel (diagnostic build: RRTFND_LIB=.../_variants/lib_phase.so, compiled with
-DRRTFND_PHASE_CLOCKS): python scripts/phase_probe.py <workload> <planners> [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rrt_star_fnd_algorithm_b200 import _abi, workloads as W
from rrt_star_fnd_algorithm_b200.batch import PlannerBatch

NAMES = ["sample+occupancy", "nearest scan", "nearest segment test", "annulus search", "steer+insert", "near-set scan+staging",
         "candidate segment tests", "parent-chain walks", "choose-parent eval + rewire decisions", "iteration tail (goal test, best-path refresh, trace)",
         "rewire commit", "forced removal"]
name, P = sys.argv[1], int(sys.argv[2])
w = W.WORKLOADS[name]()
if len(sys.argv) > 3: w["iter_num"] = int(sys.argv[3])
s, g, seeds, sids = W.planner_problems(w, 0, P)
b = PlannerBatch(w["mode"], starts=s, goals=g, obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=w["iter_num"],
                 step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"],
                 max_nodes=w["max_nodes"], seeds=seeds, stream_ids=sids)
out = (C.c_ulonglong * 16)()
fn = _abi.lib.rrtfnd_debug_phase_clocks
fn.argtypes = [C.c_void_p, C.c_int]
b.reset(); b.run(); torch.cuda.synchronize()
fn(out, 1)
b.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
fn(out, 1)
v = np.array(list(out)[:12], dtype=np.float64)
pl = b.planners()
print(f"{name} P={P} ms={e0.elapsed_time(e1):.1f} ext/s={pl['iter'].sum() / e0.elapsed_time(e1) * 1e3:.3e} (instrumented build)")
for n, x in zip(NAMES, v):
    print(f"  {100 * x / v.sum():5.1f} %  {x / pl['iter'].sum():9.0f} clk/ext  {n}")
