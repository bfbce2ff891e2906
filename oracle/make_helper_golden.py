"""Golden vectors for the per-iteration helper functions (new_node, nearest_node[_kdtree], steer, choose_parent[_kdtree],
rewire[_kdtree], forced_removal, intersection_circle and its pieces calc_delta / delta_solutions / is_between, remove_children,
reconnect_ancestors, delete_ancestors, get_near_nodes, check_for_tree_associativity;
src/algorithm.py:36-51, :328-341, :407-434, :531-551, :555-635, :666-746, :800-935) from the UNMODIFIED reference driven by Python's real
global `random` — TEST INFRASTRUCTURE, build container only.

    python -m oracle.make_helper_golden        # writes tests/golden/helpers_*.npz

`drive(mod, graph_mod, map_mod, line_mod, kdtree_cls, case)` runs the SAME call sequence on whichever implementation it is
handed - the reference's modules here, the product's modules in tests/test_helpers.py - and returns everything observable:
return values, side effects on the Graph, the next random.random(). Two harness shims as in make_fnd_golden.py (SURVEY
App. B12); ball queries go through AscendingKDTree so that choose_parent_kdtree's order-dependent result is pinned.
"""
import contextlib
import copy
import io
import os
import random

import numpy as np

from oracle import cases as K

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
HELPER_CASES = ["float_star", "int_star"]


def _arrays(G, prefix, out):
    ids = sorted(G.vertices)
    out[f"{prefix}_ids"] = np.array(ids, np.int32)
    out[f"{prefix}_xy"] = np.array([[float(G.vertices[i][0]), float(G.vertices[i][1])] for i in ids], np.float64).reshape(-1, 2)
    out[f"{prefix}_parent"] = np.array([-1 if G.parent.get(i) is None else G.parent[i] for i in ids], np.int32)
    out[f"{prefix}_cost"] = np.array([float(G.cost.get(i, 0.0)) for i in ids], np.float64)
    out[f"{prefix}_nchild"] = np.array([len(G.children[i]) for i in ids], np.int32)


def drive(alg, graph_mod, map_mod, line_mod, kdtree_cls, case):
    out = {}
    xy = case["xy_range"]
    start, goal = tuple(case["start"]), tuple(case["goal"])
    m = map_mod.Map((xy[0][1] - xy[0][0], xy[1][1] - xy[1][0]), start, goal, case["node_radius"])
    m.add_obstacles([((ox, oy), r) for ox, oy, r in case["obstacles"]])
    obstacles = m.obstacles_c
    G = graph_mod.Graph(start, goal, 0, 0, xy)
    radius, step = case["radius"], case["step_length"]

    # ---- a tree grown call by call, exactly as the body of RRT_star does it (src/algorithm.py:139-149)
    random.seed(4242)
    rec = dict(id_new=[], id_near=[], dist=[], qx=[], qy=[], cp_parent=[], cp_cost=[], cost_after_cp=[], none=0)
    it = 0
    while it < 140:
        q_new, id_new, id_near, d = alg.new_node(G, m, obstacles, step, case["bias"])
        if q_new is None:
            rec["none"] += 1
            continue
        best_edge = (id_new, id_near, d)
        G.cost[id_new] = d
        G.parent[id_new] = id_near
        kd = kdtree_cls(list(G.vertices.values()))
        ret = alg.choose_parent_kdtree(G, q_new, id_new, best_edge, radius, obstacles, kdtree=kd)
        rec["cost_after_cp"].append(float(G.cost[id_new]))      # the function's side effect (:849)
        G.add_edge(*best_edge)
        alg.rewire_kdtree(G, q_new, id_new, radius, obstacles, kdtree=kd)
        for k, v in zip(("id_new", "id_near", "dist", "qx", "qy", "cp_parent", "cp_cost"),
                        (id_new, id_near, float(d), float(q_new[0]), float(q_new[1]), ret[1], float(ret[2]))):
            rec[k].append(v)
        it += 1
    out["next_random_after_growth"] = np.float64(random.random())
    for k, v in rec.items():
        out[f"grow_{k}"] = np.array(v)
    _arrays(G, "t0", out)
    out["t0_last_id"] = np.int64(G.last_id)

    # ---- nearest_node_kdtree / nearest_node on query points: free space, inside circles, existing vertices
    rng = np.random.default_rng(77)
    qs = [tuple(float(v) for v in rng.uniform([xy[0][0], xy[1][0]], [xy[0][1], xy[1][1]])) for _ in range(48)]
    qs += [tuple(G.vertices[i]) for i in (0, 5, 17)]
    qs += [(float(ox) + 0.01, float(oy) - 0.02) for ox, oy, _ in case["obstacles"][:6]]
    kd = kdtree_cls(np.array(list(G.vertices.values())))
    separate = [i for i in G.vertices if i % 3 == 0]
    nk, nb, nk_pos = [], [], []
    for q in qs:
        pos, i = alg.nearest_node_kdtree(G, q, obstacles, kdtree=kd)
        nk.append(-1 if i is None else i); nk_pos.append([float(pos[0]), float(pos[1])])
        pos2, j = alg.nearest_node(G, q, obstacles, separate_tree_nodes=separate)
        nb.append(-1 if j is None else j)
    out["q"], out["nearest_kd_id"], out["nearest_kd_pos"], out["nearest_brute_id"] = np.array(qs), np.array(nk), np.array(nk_pos), np.array(nb)

    # ---- steer
    st = []
    for k in range(24):
        a, b = qs[k], qs[k + 1]
        r = alg.steer(a, b, step if k % 2 else 1000.0)
        st.append([float(r[0]), float(r[1]), float(r is a)])
    out["steer"] = np.array(st)

    # ---- intersection_circle on Line objects, vertical ones included
    ic, segs = [], []
    for k in range(300):
        p1 = tuple(float(v) for v in rng.uniform(xy[0][0], xy[0][1], 2))
        p2 = tuple(float(v) for v in rng.uniform(xy[0][0], xy[0][1], 2))
        if k % 10 == 0:
            p2 = (p1[0], p2[1])
        ob = obstacles[k % len(obstacles)]
        if k % 3 == 0:                                              # aim at the circle so that hits are common
            p2 = (float(ob[0][0]) + float(rng.uniform(-1, 1)) * ob[1], float(ob[0][1]) + float(rng.uniform(-1, 1)) * ob[1])
        segs.append(p1 + p2 + (k % len(obstacles),))
        ic.append(bool(alg.intersection_circle(line_mod.Line(p1, p2), ob)))
    out["ic_segs"], out["ic_hit"] = np.array(segs), np.array(ic)

    # ---- brute-force twins on copies of the tree: one more node, choose_parent / rewire
    H = copy.deepcopy(G)
    random.seed(99)
    while True:
        q_new, id_new, id_near, d = alg.new_node(H, m, obstacles, step, case["bias"])
        if q_new is not None:
            break
    H.cost[id_new] = d
    H.parent[id_new] = id_near
    ret = alg.choose_parent(H, q_new, id_new, (id_new, id_near, d), radius, obstacles)
    out["brute_cp"] = np.array([ret[1], float(ret[2]), float(H.cost[id_new])])
    H.cost[id_new] = d                      # undo the side effect so that rewire sees the planner's state
    H.add_edge(id_new, id_near, d)
    alg.rewire(H, q_new, id_new, radius * 1.5, obstacles)
    _arrays(H, "brute", out)

    # ---- forced_removal: a handful of draws
    F = copy.deepcopy(G)
    random.seed(5)
    leaf = max(F.vertices)
    out["removed"] = np.array([alg.forced_removal(F, leaf, [leaf - 1, 3, 0]) for _ in range(12)])
    out["next_random_after_removal"] = np.float64(random.random())
    _arrays(F, "fr", out)

    # ---- remove_children / check_for_tree_associativity / get_near_nodes
    R = copy.deepcopy(G)
    node = max(R.vertices, key=lambda i: len(R.children[i]))
    path = [node]
    n = node
    while R.parent[n] is not None:
        n = R.parent[n]
        path.append(n)
    keep_child = R.children[path[1]][0] if len(path) > 1 else None
    with contextlib.redirect_stdout(io.StringIO()):
        alg.remove_children(R, path[-1], path + ([keep_child] if keep_child is not None else []))
    _arrays_present(R, "rc", out)
    out["assoc"] = np.array([bool(alg.check_for_tree_associativity(G, 0, i)) for i in sorted(G.vertices)][:64])
    out["near_nodes"] = np.array(sorted(alg.get_near_nodes(G, 7, radius * 2, 0)))

    # ---- reconnect_ancestors (re-rooting) / delete_ancestors on copies; children ORDER matters to later draws, so keep it
    A = copy.deepcopy(G)
    deep = max(A.vertices, key=lambda i: len(path_to_root(A, i)))
    alg.reconnect_ancestors(A, deep)
    _arrays_present(A, "ra", out)
    out["ra_children"] = np.array([c for i in sorted(A.vertices) for c in A.children[i] + [-1]], np.int32)
    A2 = copy.deepcopy(G)
    alg.reconnect_ancestors(A2, A2.children[0][0])                  # a child of the root: the loop body never runs
    _arrays_present(A2, "ra2", out)
    D = copy.deepcopy(G)
    with contextlib.redirect_stdout(io.StringIO()):
        alg.delete_ancestors(D, deep)
    out["da_ids"] = np.array(sorted(D.vertices), np.int32)

    # ---- the scalar pieces of intersection_circle
    cd = []
    for k in range(120):
        s = out["ic_segs"][k]
        if s[0] == s[2]:
            continue
        ln = line_mod.Line((float(s[0]), float(s[1])), (float(s[2]), float(s[3])))
        delta, a, b = alg.calc_delta(ln, obstacles[int(s[4])])
        row = [float(delta), float(a), float(b), np.nan, np.nan, 0.0, 0.0]
        if delta >= 0:
            x1, x2 = alg.delta_solutions(delta, a, b)
            row[3:] = [float(x1[0]), float(x2[0]), float(alg.is_between(x1, ln.p1, ln.p2)), float(alg.is_between(x2, ln.p1, ln.p2))]
        cd.append(row)
    out["calc_delta"] = np.array(cd)
    return out


def path_to_root(G, i):
    p = [i]
    while G.parent[p[-1]] is not None:
        p.append(G.parent[p[-1]])
    return p


def _arrays_present(G, prefix, out):
    """Like _arrays for a Graph that may hold dangling parents (remove_children leaves them)."""
    ids = sorted(G.vertices)
    out[f"{prefix}_ids"] = np.array(ids, np.int32)
    out[f"{prefix}_parent"] = np.array([-1 if G.parent.get(i) is None else G.parent[i] for i in ids], np.int32)
    out[f"{prefix}_nchild"] = np.array([len(G.children[i]) for i in ids], np.int32)


def main():
    from oracle.refharness import AscendingKDTree, load_reference
    ref = load_reference()
    alg = ref.algorithm
    orig_gdd = alg.get_distance_dict
    alg.get_distance_dict = lambda G_, node, idx=None: orig_gdd(G_, node, idx)      # shim B12
    try:
        for name in HELPER_CASES:
            with contextlib.redirect_stdout(io.StringIO()):
                out = drive(alg, ref.graph, ref.map, ref.line, AscendingKDTree, K.case_by_name(name))
            np.savez_compressed(os.path.join(OUT, f"helpers_{name}.npz"), **out)
            print(name, "nodes", len(out["t0_ids"]), "rejected", int(out["grow_none"]), "nearest none", int((out["nearest_kd_id"] < 0).sum()),
                  "ic hits", int(out["ic_hit"].sum()), "removed", out["removed"][:4], "near", len(out["near_nodes"]))
    finally:
        alg.get_distance_dict = orig_gdd


if __name__ == "__main__":
    main()
