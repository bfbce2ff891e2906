"""Golden vectors for the drop-in API: the UNMODIFIED reference driven by Python's real global `random`
(random.seed(s), no stream injection) — TEST INFRASTRUCTURE, build container only.

    python -m oracle.make_dropin_golden        # writes tests/golden/dropin_*.npz

Each case is a sequence of planner calls on one Graph (a second call starts from the tree the first one left, which
exercises the upload of an existing tree). Recorded: the Graph dictionaries after the last call, every call's return
value, and the next random.random() — the GPU drop-in must reproduce all of it from the same seed.
"""
import os
import random

import numpy as np

from oracle import cases as K
from oracle.refharness import patched_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DROPIN_CASES = [
    dict(name="star_main", seed=0, case="main_star", calls=[("RRT_star", 300)], node_radius_arg=20),
    dict(name="fn_float", seed=5, case="float_fn", calls=[("RRT_star_FN", 500)], node_radius_arg=1),
    dict(name="rrt_main", seed=3, case="main_rrt", calls=[("RRT", 200)], node_radius_arg=20),
    # two calls on one Graph, and a goal radius different from the map's node radius
    dict(name="rrt_then_star", seed=11, case="int_star", calls=[("RRT", 40), ("RRT_star", 200)], node_radius_arg=8),
    dict(name="star_then_fn", seed=12, case="float_star", calls=[("RRT_star", 120), ("RRT_star_FN", 300)], node_radius_arg=1),
    # RRT*-FN on a Graph that already holds more nodes than the slack of a fresh FN run: RRT_star_FN counts only its own
    # insertions (n_of_nodes starts at 1, src/algorithm.py:209), so the live count peaks at 401 + 350
    dict(name="star_then_fn_big", seed=13, case="float_star", calls=[("RRT_star", 400), ("RRT_star_FN", 900)], node_radius_arg=1,
         max_nodes=350),
]


def build_objects(ref_graph, ref_map, case):
    xy = case["xy_range"]
    start, goal = tuple(case["start"]), tuple(case["goal"])
    m = ref_map.Map((xy[0][1] - xy[0][0], xy[1][1] - xy[1][0]), start, goal, case["node_radius"])
    m.add_obstacles([((ox, oy), r) for ox, oy, r in case["obstacles"]])
    g = ref_graph.Graph(start, goal, 0, 0, xy)
    return g, m


def call_planner(mod, name, G, m, case, iters, node_radius_arg, max_nodes=None):
    if name == "RRT":
        return mod.RRT(G, iters, m, case["step_length"], node_radius_arg, bias=case["bias"])
    if name == "RRT_star":
        return mod.RRT_star(G, iters, m, case["step_length"], case["radius"] or 30, node_radius_arg, bias=case["bias"])
    return mod.RRT_star_FN(G, iters, m, case["step_length"], case["radius"], node_radius_arg,
                           max_nodes=case["max_nodes"] if max_nodes is None else max_nodes, bias=case["bias"])


def graph_arrays(G):
    ids = np.array(sorted(G.vertices), dtype=np.int32)
    return dict(
        ids=ids,
        x=np.array([float(G.vertices[i][0]) for i in ids]), y=np.array([float(G.vertices[i][1]) for i in ids]),
        parent=np.array([-1 if G.parent[i] is None else G.parent[i] for i in ids], dtype=np.int32),
        cost=np.array([float(G.cost[i]) for i in ids]),
        nchild=np.array([len(G.children[i]) for i in ids], dtype=np.int32),
        last_id=np.int64(G.last_id),
    )


def main():
    import contextlib, io, sys
    only = set(sys.argv[1:])                      # optional: names of the cases to (re)generate
    for dc in DROPIN_CASES:
        if only and dc["name"] not in only:
            continue
        case = K.case_by_name(dc["case"])
        with patched_reference(random, ascending_ball=True) as ref:     # rng = the real module: nothing is injected
            G, m = build_objects(ref.graph, ref.map, case)
            random.seed(dc["seed"])
            rets = []
            with contextlib.redirect_stdout(io.StringIO()):
                for name, iters in dc["calls"]:
                    rets.append(call_planner(ref.algorithm, name, G, m, case, iters, dc["node_radius_arg"], dc.get("max_nodes")))
            nxt = random.random()
        out = graph_arrays(G)
        out["next_random"] = np.float64(nxt)
        for i, r in enumerate(rets):
            if isinstance(r, tuple):
                out[f"ret{i}_iter"] = np.int64(r[0])
                out[f"ret{i}_path"] = np.array(r[1]["path"], dtype=np.int32)
                out[f"ret{i}_cost"] = np.float64(r[1]["cost"])
            else:
                out[f"ret{i}_iter"] = np.int64(r)
        np.savez_compressed(os.path.join(OUT, f"dropin_{dc['name']}.npz"), **out)
        print(dc["name"], "nodes", len(out["ids"]), "last_id", int(out["last_id"]), "rets", [int(out[f"ret{i}_iter"]) for i in range(len(rets))])


if __name__ == "__main__":
    main()
