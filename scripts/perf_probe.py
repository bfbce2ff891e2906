"""Ad-hoc timing probe: python scripts/perf_probe.py <workload> <planners> [iters] [eval_cp]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rrt_star_fnd_algorithm_b200 import workloads as W
from rrt_star_fnd_algorithm_b200.batch import PlannerBatch

name, P = sys.argv[1], int(sys.argv[2])
w = W.WORKLOADS[name]()
if len(sys.argv) > 3: w["iter_num"] = int(sys.argv[3])
ecp = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
launch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ncap = int(sys.argv[6]) if len(sys.argv) > 6 else 64
cell = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0
s, g, seeds, sids = W.planner_problems(w, 0, P)
b = PlannerBatch(w["mode"], starts=s, goals=g, obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=w["iter_num"],
                 step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"],
                 max_nodes=w["max_nodes"], seeds=seeds, stream_ids=sids, eval_choose_parent=ecp, launch=launch, near_cap=ncap, grid_cell=cell,
                 **(dict(node_grid=True, finish_cap=65536, path_cap=65536, node_grid_cell=float(os.environ.get("NG_CELL", 0)), node_grid_min_nodes=int(os.environ.get("NG_MIN", 4096))) if w.get("node_grid") or os.environ.get("NODE_GRID") else {}))
for rep in range(2):
    b.reset(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
pl = b.planners()
ext = pl["iter"].sum()
print(f"{name} launch={launch} cell={cell} P={P} iters={w['iter_num']} smem={b.smem_bytes} ms={ms:.1f} ext/s={ext/ms*1e3:.3e} edge/s={pl['n_edge_checks'].sum()/ms*1e3:.3e} "
      f"attempts={pl['attempts'].sum()} status={np.unique(pl['status'])} scan_nodes={pl['n_scan_nodes'].sum():.3e} hops={pl['n_chain_hops'].sum():.3e} "
      f"circle={pl['n_circle_tests'].sum():.3e} rew={pl['n_rewired'].sum()} solved={pl['solution_found'].sum()} "
      f"algGB/s={(16*pl['n_scan_nodes'].sum()+12*pl['n_chain_hops'].sum())/ms/1e6:.1f}")
