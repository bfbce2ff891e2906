"""c3 at the per-GPU planner counts of the sharded config (N = 2, 4: 8,192 and 4,096 planners) in one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_star_fnd_algorithm_b200 import workloads as W
from rrt_star_fnd_algorithm_b200.batch import PlannerBatch

w = W.c3()
for P in [int(a) for a in sys.argv[1:]] or [8192, 4096]:
    s, g, seeds, sids = W.planner_problems(w, 0, P)
    b = PlannerBatch(w["mode"], starts=s, goals=g, obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=w["iter_num"],
                     step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"],
                     max_nodes=w["max_nodes"], seeds=seeds, stream_ids=sids, eval_choose_parent=True)
    best = 1e30
    for rep in range(3):
        b.reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ext = int(b.planners()["iter"].sum())
    print(f"c3 P={P} ms={best:.1f} ext/s={ext / best * 1e3:.3e}", flush=True)
    del b
    torch.cuda.empty_cache()
