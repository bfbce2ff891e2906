"""Time the UNMODIFIED Python reference on the benchmark's workloads — TEST INFRASTRUCTURE, build container only.

    python -m oracle.time_python_reference [--planners 3] [--out profiles/r02_python_reference.json]

The reference cannot travel to the GPU box (`/root/reference` exists only here), so `bench.py --impl reference` times the C
port of its algorithm there. This script supplies the missing figure: the reference's own `RRT_star` / `RRT_star_FN`
(src/algorithm.py:98, :190), driven by Python's global `random` (random.seed, no stream injection, the real cKDTree), on the
problems of `bench.py`'s workloads c3 and c2 — the same obstacles, map, step length, radius, bias, iteration count and the
first planners' start / goal pairs — one planner after the other on one core of THIS container. Only the progress bar and
the two plot calls are switched off (refharness.patched_reference); nothing in the algorithm is touched.

Counted: accepted extensions (Graph.last_id grows by one per accepted node) and wall time around the planner call.
"""
import argparse
import json
import os
import platform
import random
import time

from oracle.refharness import patched_reference
from rrt_star_fnd_algorithm_b200 import workloads as W


def run_one(ref, w, start, goal, seed):
    xy = w["xy_range"]
    m = ref.map.Map((xy[0][1] - xy[0][0], xy[1][1] - xy[1][0]), tuple(start), tuple(goal), w["node_radius"])
    m.add_obstacles([((ox, oy), r) for ox, oy, r in w["obstacles"]])
    g = ref.graph.Graph(tuple(start), tuple(goal), 0, 0, xy)
    random.seed(seed)
    t0 = time.perf_counter()
    if w["mode"] == "star":
        ref.algorithm.RRT_star(g, w["iter_num"], m, w["step_length"], w["radius"], w["node_radius"], bias=w["bias"])
    else:
        ref.algorithm.RRT_star_FN(g, w["iter_num"], m, w["step_length"], w["radius"], w["node_radius"],
                                  max_nodes=w["max_nodes"], bias=w["bias"])
    dt = time.perf_counter() - t0
    return int(g.last_id), dt, len(g.vertices)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planners", type=int, default=3)
    ap.add_argument("--workloads", default="c3,c2")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = dict(host=platform.processor() or platform.machine(), python=platform.python_version(), cores_used=1, runs=[])
    with patched_reference(random, ascending_ball=False) as ref:      # `random` itself: the reference's own generator
        for name in a.workloads.split(","):
            w = W.WORKLOADS[name]()
            starts, goals, _, _ = W.planner_problems(w, 0, a.planners)
            ext = sec = 0
            for p in range(a.planners):
                n, dt, live = run_one(ref, w, starts[p], goals[p], 1000 + p)
                ext, sec = ext + n, sec + dt
                print(f"{name} planner {p}: {n} extensions ({live} nodes live) in {dt:.1f} s = {n / dt:.1f} ext/s", flush=True)
            out["runs"].append(dict(workload=name, planners=a.planners, iter_num=w["iter_num"], extensions=ext,
                                    seconds=round(sec, 2), ext_per_s_per_core=round(ext / sec, 1)))
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
