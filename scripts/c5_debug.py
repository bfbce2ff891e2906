"""Phase-by-phase timing of the config-5 loop (debug aid): python scripts/c5_debug.py [planners] [ticks]"""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", 45)), exit=True)
import numpy as np, torch
from rrt_star_fnd_algorithm_b200 import workloads as W
from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
from rrt_star_fnd_algorithm_b200.fnd_loop import FNDReplanner, moved_obstacles

P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = W.c5()
T = int(sys.argv[2]) if len(sys.argv) > 2 else w["ticks"]
t00 = time.time()
def log(*a):
    print(f"[{time.time() - t00:7.2f}s]", *a, flush=True)
s, g, seeds, sids = W.planner_problems(w, 0, P)
b = PlannerBatch(w["mode"], starts=s, goals=g, seeds=seeds, stream_ids=sids, obstacles=w["obstacles"], xy_range=w["xy_range"],
                 iter_num=w["iter_num"], step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"],
                 bias=w["bias"], max_nodes=w["max_nodes"])
log("batch built, capacity", b.prm.capacity)
b.reset(); b.run(); torch.cuda.synchronize()
pl = b.planners()
log("planned; status", np.unique(pl["status"]), "solved", int(pl["solution_found"].sum()), "best_len max", int(pl["best_len"].max()))
b.grow_capacity(b.prm.capacity + T * w["regrow_iters"])
loop = FNDReplanner(b, w["reconnect_radius"], w["regrow_iters"])
for k in range(1, T + 1):
    t0 = time.time()
    out = loop.tick(moved_obstacles(w["obstacles"], w["velocity"], k))
    torch.cuda.synchronize()
    log(f"tick {k}: {1e3 * (time.time() - t0):.1f} ms", {a: c for a, c in loop.stats.items()},
        "regrow calls", int((out["regrow"] != 0).sum()), "max its", int(out["regrow_iters"].max()))
log("done")
