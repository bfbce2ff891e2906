"""Config 4 (one RRT* tree grown to N nodes through the device node grid): python scripts/c4_run.py <iters> [replicas]
Prints one JSON line: extensions/s of the whole launch and of one tree, nodes touched per extension, chain hops per extension,
parent-chain depth statistics of the final tree and, with a -DRRTFND_PHASE_CLOCKS build (RRTFND_LIB), the per-phase clock shares."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rrt_star_fnd_algorithm_b200 import _abi, workloads as W
from rrt_star_fnd_algorithm_b200.batch import PlannerBatch

iters, R = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = W.c4()
w["iter_num"] = iters
s, g, seeds, sids = W.planner_problems(w, 0, R)
cap = max(65536, iters // 2)
b = PlannerBatch(w["mode"], starts=s, goals=g, obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=iters,
                 step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"], seeds=seeds,
                 stream_ids=sids, node_grid=True, finish_cap=cap, path_cap=cap, node_grid_min_nodes=int(os.environ.get("NG_MIN", 4096)),
                 node_grid_cell=float(os.environ.get("NG_CELL", 0)))
phase = hasattr(_abi.lib, "rrtfnd_debug_phase_clocks")
out = (C.c_ulonglong * 16)()
if phase:
    _abi.lib.rrtfnd_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int]
    _abi.lib.rrtfnd_debug_phase_clocks(out, 1)
b.reset(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
pl = b.planners()
b.check_status()
t = b.tree(0)
ids, par = t["ids"], t["parent"]
slot = {int(i): k for k, i in enumerate(ids.tolist())}
ps = np.array([slot[int(q)] if q >= 0 else -1 for q in par])
depth = np.zeros(len(ids), np.int64)
order = np.argsort(ids)                       # parents of rewired nodes may be younger: iterate to a fixed point
todo = ps >= 0
d = np.where(todo, 1, 0)
anc = ps.copy()
while (anc >= 0).any():                       # pointer jumping: depth = hops to the root
    live = anc >= 0
    nxt = np.where(live, ps[np.clip(anc, 0, None)], -1)
    d = d + np.where(nxt >= 0, 1, 0) * live
    anc = nxt
line = {"workload": "c4", "iters": iters, "replicas": R, "ms": ms, "ext_per_s_all": float(pl["iter"].sum() / ms * 1e3),
        "ext_per_s_one_tree": float(pl["iter"][0] / ms * 1e3), "nodes_touched_per_ext": float(pl["n_scan_nodes"].sum() / pl["iter"].sum()),
        "chain_hops_per_ext": float(pl["n_chain_hops"].sum() / pl["iter"].sum()), "near_set_mean": float((pl["n_edge_checks"].sum() - pl["iter"].sum()) / pl["iter"].sum()),
        "circle_tests_per_ext": float(pl["n_circle_tests"].sum() / pl["iter"].sum()), "rewired_per_ext": float(pl["n_rewired"].sum() / pl["iter"].sum()),
        "depth_mean": float(d.mean()), "depth_p50": float(np.median(d)), "depth_max": int(d.max()), "best_cost": float(pl["best_cost"][0]),
        "solved": int(pl["solution_found"].sum())}
if phase:
    _abi.lib.rrtfnd_debug_phase_clocks(out, 1)
    v = np.array(list(out)[:12], dtype=np.float64)
    names = ["sample+occupancy", "nearest (grid gather + scan)", "nearest segment test", "annulus search", "steer+insert", "near-set gather + staging",
             "candidate segment tests", "parent-chain walks (get_cost)", "choose-parent + rewire decisions", "iteration tail", "rewire commit", "forced removal"]
    line["phase_share"] = {n: round(float(x / v.sum()), 4) for n, x in zip(names, v)}
print(json.dumps(line))
