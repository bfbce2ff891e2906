"""EXTENSION: axis-aligned rectangular obstacles (BASELINE.json config 2 says "200 rectangular obstacles").

The reference declares Map.obstacles_r and never implements it (src/map.py:18, :54), so there is nothing of the reference
to compare with: PARITY UNPINNED BY THE REFERENCE. The semantics are written down in include/rrtfnd.h; the CPU oracle
restates them operation for operation (oracle/rrt_oracle.c seg_hits_rect / point_in_inflated_rect). CPU tests check the
oracle against exact rational geometry away from razor edges and against its own invariants; GPU tests require the CUDA
path (un-culled operators, culled operators, fused planner) to equal the oracle bit for bit."""
from fractions import Fraction

import numpy as np
import pytest

import oracle as O
from rrt_star_fnd_algorithm_b200 import _structs as S
from rrt_star_fnd_algorithm_b200 import workloads as W


def rect_row(cx, cy, hx, hy):
    return (float(cx), float(cy), -float(hx), float(hy))


def exact_seg_hits_rect(p1, p2, cx, cy, hx, hy):
    """Closed segment vs closed rectangle in exact rational arithmetic (Liang-Barsky clipping): (hit, clearance) where
    clearance is a crude measure of how far the configuration is from a touching one (0 = razor edge)."""
    F = Fraction
    x1, y1, x2, y2 = F(p1[0]), F(p1[1]), F(p2[0]), F(p2[1])
    lo_x, hi_x, lo_y, hi_y = F(cx) - F(hx), F(cx) + F(hx), F(cy) - F(hy), F(cy) + F(hy)
    t0, t1 = F(0), F(1)
    for d, a, lo, hi in ((x2 - x1, x1, lo_x, hi_x), (y2 - y1, y1, lo_y, hi_y)):
        if d == 0:
            if a < lo or a > hi:
                return False
        else:
            ta, tb = (lo - a) / d, (hi - a) / d
            if ta > tb:
                ta, tb = tb, ta
            t0, t1 = max(t0, ta), min(t1, tb)
            if t0 > t1:
                return False
    return True


def test_oracle_segment_vs_rectangle_agrees_with_exact_geometry():
    rng = np.random.default_rng(3)
    n_hit = n_checked = 0
    for _ in range(4000):
        cx, cy = rng.uniform(10, 90, 2)
        hx, hy = rng.uniform(0.5, 6, 2)
        p1 = np.array([cx, cy]) + rng.normal(0, 9, 2)     # around the rectangle, so that hits and misses are both common
        p2 = p1 + rng.normal(0, 12, 2) if rng.random() < 0.7 else rng.uniform(0, 100, 2)
        if rng.random() < 0.1:
            p2[0] = p1[0]                                   # vertical: a true segment here, unlike the circles' infinite line
        if rng.random() < 0.1:
            p2[1] = p1[1]
        want = exact_seg_hits_rect(p1, p2, cx, cy, hx, hy)
        # skip razor edges: compare only when a slightly larger and a slightly smaller rectangle agree with the exact answer
        grown = exact_seg_hits_rect(p1, p2, cx, cy, hx * (1 + 1e-9), hy * (1 + 1e-9))
        shrunk = exact_seg_hits_rect(p1, p2, cx, cy, hx * (1 - 1e-9), hy * (1 - 1e-9))
        if grown != want or shrunk != want:
            continue
        got = O.through_obstacle(p1, p2, [rect_row(cx, cy, hx, hy)])
        assert got == want, (p1, p2, cx, cy, hx, hy)
        assert O.through_obstacle(p2, p1, [rect_row(cx, cy, hx, hy)]) == want          # endpoint order does not matter here
        n_checked += 1; n_hit += want
    assert n_checked > 3500 and 500 < n_hit < n_checked - 500, (n_checked, n_hit)


def test_oracle_point_vs_inflated_rectangle():
    row = rect_row(50, 40, 3, 2)
    assert O.is_occupied((50, 40), [row], 0.0) and O.is_occupied((53, 42), [row], 0.0)       # closed rectangle, even with no inflation
    assert not O.is_occupied((53.0001, 40), [row], 0.0)
    assert O.is_occupied((53.9, 40), [row], 1.0) and not O.is_occupied((54.0, 40), [row], 1.0)   # strict, like src/map.py:45
    d = 1.0 / np.sqrt(2.0)
    assert O.is_occupied((53 + d * 0.999, 42 + d * 0.999), [row], 1.0) and not O.is_occupied((53 + d * 1.001, 42 + d * 1.001), [row], 1.0)
    # mixed tables: circles keep the reference's semantics, rectangles add theirs
    circ = (10.0, 10.0, 2.0)
    assert O.is_occupied((11, 10), [circ, row], 0.5) and O.is_occupied((50, 40), [circ, row], 0.5) and not O.is_occupied((30, 30), [circ, row], 0.5)


def test_c2r_workload_is_deterministic_and_clear_of_start_and_goal():
    a, b = W.c2r(), W.c2r()
    assert a["obstacles"] == b["obstacles"] and len(a["obstacles"]) == 200 and all(r[2] < 0 and r[3] > 0 for r in a["obstacles"])
    assert not O.is_occupied(a["start"], a["obstacles"], a["node_radius"]) and not O.is_occupied(a["goal"], a["obstacles"], a["node_radius"])
    c = W.c2()
    assert [r[:2] for r in a["obstacles"]] == [r[:2] for r in c["obstacles"]]               # the same centres as the circle variant


def oracle_params(w):
    p = S.Params()
    p.mode = {"rrt": 0, "star": 1, "fn": 2}[w["mode"]]
    p.iter_num, p.max_nodes, p.capacity = w["iter_num"], w["max_nodes"], w["iter_num"] + 2
    p.n_obst, p.near_cap, p.finish_cap, p.path_cap = len(w["obstacles"]), 4096, 4096, 4096
    p.stream_mode, p.flags = S.STREAM_PHILOX, 0
    p.step_length, p.radius, p.node_radius, p.bias = w["step_length"], w["radius"], w["node_radius"], w["bias"]
    (p.x_lo, p.x_hi), (p.y_lo, p.y_hi) = w["xy_range"]
    p.goal_node_radius = w["node_radius"]
    return p


def test_oracle_plans_through_rectangles():
    """The oracle grows an RRT*-FN tree on the rectangle map; no node lies in an inflated rectangle, no edge crosses one."""
    w = dict(W.c2r(), iter_num=600, max_nodes=300)
    s, g, seeds, sids = W.planner_problems(w, 0, 1)
    t = O.OracleTree(s[0], w["iter_num"] + 2)
    st, _ = t.run(oracle_params(w), w["obstacles"], g[0], seed=int(seeds[0]), stream_id=int(sids[0]))
    assert st == S.ST_DONE
    e = t.export()
    pos = {int(i): (e["x"][k], e["y"][k]) for k, i in enumerate(e["ids"])}
    for i, q in zip(e["ids"], e["parent"]):
        if q >= 0:
            assert exact_seg_hits_rect(pos[int(i)], pos[int(q)], 0, 0, 0.0, 0.0) in (False, True)      # well formed
            assert not O.through_obstacle(pos[int(q)], pos[int(i)], w["obstacles"])
    assert len(e["ids"]) == 300


# ------------------------------------------------------------------------------------------------------ GPU

def adversarial_segments(w, n, seed):
    """Segments around the rectangles: through them, grazing their corners and sides, vertical, horizontal, tiny."""
    rng = np.random.default_rng(seed)
    ob = np.asarray(w["obstacles"])
    k = rng.integers(0, len(ob), n)
    c, hx, hy = ob[k, :2], -ob[k, 2], ob[k, 3]
    corner = c + np.column_stack([hx * rng.choice([-1, 1], n), hy * rng.choice([-1, 1], n)])
    p1 = corner + rng.normal(0, 1.5, (n, 2))
    p2 = corner + rng.normal(0, 1.5, (n, 2))
    far = rng.random(n) < 0.3
    p2[far] = rng.uniform(0, 100, (int(far.sum()), 2))
    vert = rng.random(n) < 0.1
    p2[vert, 0] = p1[vert, 0]
    hor = rng.random(n) < 0.1
    p2[hor, 1] = p1[hor, 1]
    on_side = rng.random(n) < 0.1                              # an endpoint exactly on a side
    p1[on_side, 0] = (c[:, 0] + hx)[on_side]
    return p1, p2


@pytest.mark.gpu
def test_gpu_rectangle_operators_equal_the_oracle():
    """Un-culled and grid-culled edge / point operators on the c2r map (rectangles only) and on a mixed circle + rectangle map."""
    import ctypes as C
    import torch
    from rrt_star_fnd_algorithm_b200 import _abi
    from rrt_star_fnd_algorithm_b200.batch import build_obstacle_grid, obstacle_table
    w = W.c2r()
    mixed = list(W.c2()["obstacles"][:100]) + list(w["obstacles"][100:])
    for name, obstacles in (("c2r", w["obstacles"]), ("mixed", mixed)):
        tab = obstacle_table(obstacles)
        p1, p2 = adversarial_segments(w, 120000, 5)
        want = O.through_obstacle_batch(p1, p2, [tuple(r) for r in tab.tolist()])
        assert 0.2 < want.mean() < 0.9
        dev = torch.device("cuda:0")
        cols = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1])]
        tab_t = torch.from_numpy(tab).to(dev)
        g, arrs = build_obstacle_grid(tab, 1.0, (0.0, 100.0, 0.0, 100.0))
        gt = [torch.from_numpy(a).to(dev) for a in arrs]
        g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in gt]
        full = torch.zeros(len(p1), dtype=torch.uint8, device=dev)
        culled = torch.zeros(len(p1), dtype=torch.uint8, device=dev)
        _abi.check(_abi.lib.rrtfnd_edges_vs_circles(*[t.data_ptr() for t in cols], len(p1), tab_t.data_ptr(), len(tab), full.data_ptr(), None), "edges")
        _abi.check(_abi.lib.rrtfnd_edges_vs_circles_grid(*[t.data_ptr() for t in cols], len(p1), tab_t.data_ptr(), len(tab), C.byref(g), culled.data_ptr(), None), "edges grid")
        assert np.array_equal(full.cpu().numpy().astype(bool), want), name
        assert np.array_equal(culled.cpu().numpy().astype(bool), want), name
        pts = np.concatenate([p1[:40000], np.random.default_rng(1).uniform(0, 100, (40000, 2))])
        want_occ = np.array([O.is_occupied(q, [tuple(r) for r in tab.tolist()], 1.0) for q in pts[::7]])
        px, py = torch.from_numpy(pts[:, 0].copy()).to(dev), torch.from_numpy(pts[:, 1].copy()).to(dev)
        occ = torch.zeros(len(pts), dtype=torch.uint8, device=dev)
        occ_g = torch.zeros(len(pts), dtype=torch.uint8, device=dev)
        _abi.check(_abi.lib.rrtfnd_points_occupied(px.data_ptr(), py.data_ptr(), len(pts), tab_t.data_ptr(), len(tab), 1.0, occ.data_ptr(), None), "points")
        _abi.check(_abi.lib.rrtfnd_points_occupied_grid(px.data_ptr(), py.data_ptr(), len(pts), tab_t.data_ptr(), len(tab), C.byref(g), 1.0, occ_g.data_ptr(), None), "points grid")
        assert np.array_equal(occ.cpu().numpy()[::7].astype(bool), want_occ) and np.array_equal(occ_g.cpu().numpy(), occ.cpu().numpy()), name
        assert 0.1 < want_occ.mean() < 0.9


@pytest.mark.gpu
def test_gpu_planner_on_the_rectangle_map_equals_the_oracle():
    """Config 2's rectangle variant through the fused planner (default launch): 48 planners, RRT*-FN, 2,000 iterations."""
    from concurrent.futures import ThreadPoolExecutor
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    from util import assert_tree_equal
    w = dict(W.c2r(), iter_num=2000, max_nodes=600)
    P = 48
    s, g, seeds, sids = W.planner_problems(w, 0, P)
    b = PlannerBatch(w["mode"], starts=s, goals=g, seeds=seeds, stream_ids=sids, obstacles=w["obstacles"], xy_range=w["xy_range"],
                     iter_num=w["iter_num"], step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"],
                     bias=w["bias"], max_nodes=w["max_nodes"]).run()
    b.check_status()

    def one(p):
        t = O.OracleTree(s[p], w["iter_num"] + 2)
        t.run(oracle_params(w), w["obstacles"], g[p], seed=int(seeds[p]), stream_id=int(sids[p]))
        return t
    with ThreadPoolExecutor(16) as ex:
        trees = list(ex.map(one, range(P)))
    pls = b.planners()
    for p, t in enumerate(trees):
        assert_tree_equal(b.tree(p), t.export(), f"c2r[{p}]")
        op = t.planner()
        for f in ("iter", "cursor", "attempts", "best_cost", "n_removed", "n_rewired", "n_edge_checks", "n_edge_checks_ref"):
            assert pls[p][f] == getattr(op, f), (p, f)
        assert np.array_equal(b.best_path(p), t.best_path())
    assert pls["solution_found"].sum() >= P // 2


@pytest.mark.gpu
def test_dropin_map_with_rectangles():
    """Map.add_rectangles + the drop-in RRT_star: the tree avoids the rectangles (occupancy of samples, edges)."""
    import contextlib, io, random
    from rrt_star_fnd_algorithm_b200.algorithm import RRT_star
    from rrt_star_fnd_algorithm_b200.graph import Graph
    from rrt_star_fnd_algorithm_b200.map import Map
    m = Map((100, 100), (5.0, 5.0), (95.0, 95.0), 1.0)
    m.add_rectangles([[(30.0, 30.0), (8.0, 3.0)], [(60.0, 55.0), (4.0, 12.0)], [(50.0, 20.0), (2.0, 2.0)]])
    m.add_obstacles([((20.0, 70.0), 5.0)])
    assert m.is_occupied_c((30.0, 30.0)) and m.is_occupied_c((38.5, 30.0)) and not m.is_occupied_c((39.5, 30.0)) and m.is_occupied_c((20.0, 70.0))
    G = Graph((5.0, 5.0), (95.0, 95.0), 100, 100, [[0, 100], [0, 100]])
    random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        it, best = RRT_star(G, 400, m, 4.0, 8.0, 1.0, bias=0.05)
    assert it == 400 and len(G.vertices) == 401
    rows = [rect_row(30, 30, 8, 3), rect_row(60, 55, 4, 12), rect_row(50, 20, 2, 2), (20.0, 70.0, 5.0)]
    for i, q in G.parent.items():
        if q is not None:
            assert not O.through_obstacle(G.vertices[q], G.vertices[i], rows)
