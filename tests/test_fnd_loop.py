"""RRT*-FND replanning loop (BASELINE.json config 5; rrt_star_fnd_algorithm_b200/fnd_loop.py).

The reference ships the steps (select_branch / valid_path / reconnect / regrow, pinned on reference goldens in
tests/test_fnd.py) but not the loop (RRT_STAR_FND stops at a stub, src/algorithm.py:283-291). Here the device loop is
compared, tick by tick and bit for bit, with the same composition of the CPU oracle's steps (oracle/fnd_loop.py)."""
import numpy as np
import pytest

import oracle as O
from oracle import cases as K
from oracle.fnd_loop import OracleReplanner, MOVING, ARRIVED, LOST
from rrt_star_fnd_algorithm_b200 import _structs as S
from rrt_star_fnd_algorithm_b200 import workloads as W
from rrt_star_fnd_algorithm_b200.fnd_loop import moved_obstacles

from util import assert_tree_equal


def small_c5(iter_num=1500, max_nodes=400):
    w = W.c5()
    w.update(iter_num=iter_num, max_nodes=max_nodes)
    return w


def oracle_planners(w, P, ticks, edges=False):
    s, g, seeds, sids = W.planner_problems(w, 0, P)
    prm = K.params_for(w, stream_mode=S.STREAM_PHILOX, flags=0)
    out = []
    for p in range(P):
        t = O.OracleTree(s[p], w["iter_num"] + 2 + ticks * w["regrow_iters"])
        assert t.run(prm, w["obstacles"], g[p], seed=int(seeds[p]), stream_id=int(sids[p]))[0] == S.ST_DONE
        out.append(OracleReplanner(t, prm, g[p], seeds[p], sids[p], w["node_radius"], w["reconnect_radius"], w["regrow_iters"],
                                   edge_invalidation=edges))
    return out


def check_forest(e, path, root):
    """Child counts agree with the parents; the stored path follows parent links from the goal-side leaf to the root."""
    ids = [int(i) for i in e["ids"]]
    par = dict(zip(ids, [int(q) for q in e["parent"]]))
    cnt = {i: 0 for i in ids}
    for i, q in par.items():
        if q >= 0:
            assert q in par, f"node {i} hangs from the deleted node {q}"
            cnt[q] += 1
    assert [cnt[i] for i in ids] == [int(c) for c in e["nchild"]]
    for a, b in zip(path[:-1], path[1:]):
        assert par[int(a)] == int(b)
    assert int(path[-1]) == root and par[root] == -1


def test_c5_workload_keeps_the_goal_region_clear_and_is_deterministic():
    a, b = W.c5(), W.c5()
    assert a["obstacles"] == b["obstacles"] and np.array_equal(a["velocity"], b["velocity"]) and len(a["obstacles"]) == 50
    for k in range(a["ticks"] + 1):
        ob = np.asarray(moved_obstacles(a["obstacles"], a["velocity"], k))
        d = np.hypot(ob[:, 0] - a["goal"][0], ob[:, 1] - a["goal"][1])
        assert (d >= ob[:, 2] + 3 * a["node_radius"]).all()
    ob1 = np.asarray(moved_obstacles(a["obstacles"], a["velocity"], 1))
    assert np.array_equal(ob1[:, 0], np.asarray(a["obstacles"])[:, 0] + a["velocity"][:, 0])


def test_oracle_loop_exercises_every_branch_and_keeps_a_consistent_forest():
    w = small_c5()
    P, T = 12, w["ticks"]
    reps = oracle_planners(w, P, T)
    n = dict(split=0, linked=0, regrown=0, lost=0, arrived=0, multi=0)
    for k in range(1, T + 1):
        ob = moved_obstacles(w["obstacles"], w["velocity"], k)
        for r in reps:
            before = r.phase
            o = r.tick(ob)
            if before != MOVING:
                assert o["current"] == -1
                continue
            if o["separate_root"] >= 0 and o["separate_root"] != o["current"]:
                n["split"] += 1
                n["linked"] += sum(v >= 0 for v in o["linked"])
                n["regrown"] += sum(v == 1 for v in o["regrow"])
                n["multi"] += (o["linked"][1] >= 0) or (o["regrow"][1] != 0)
            e = r.t.export()
            assert r.t.planner().n_live == len(e["ids"])
            if r.phase != LOST:
                check_forest(e, r.t.best_path(), o["current"])
    n["lost"] = sum(r.phase == LOST for r in reps)
    assert n["split"] >= 20 and n["linked"] >= 10 and n["regrown"] >= 3 and n["multi"] >= 1 and n["lost"] < P, n


class OracleBackedBatch:
    """The slice of PlannerBatch that FNDReplanner drives, answered by oracle trees: lets the host logic of the product
    loop run on the CPU against the oracle loop (the device kernels themselves are covered by the gpu test below)."""

    def __init__(self, w, P, ticks):
        import torch
        self.torch, self.w, self.P, self.device = torch, w, P, torch.device("cpu")
        self.reps = oracle_planners(w, P, ticks)          # only their trees are used
        self.obstacles = w["obstacles"]
        self._refresh()

    def _refresh(self):
        torch = self.torch
        rows, paths, rec = [], [], np.zeros(self.P, dtype=S.PLANNER_DTYPE)
        for p, r in enumerate(self.reps):
            ids = [int(i) for i in r.t.export()["ids"]] + [-1]
            root = r.t.root()
            rec[p]["root_slot"] = ids.index(root) if root in ids[:-1] else len(ids) - 1
            path = [int(i) for i in r.t.best_path()]
            rec[p]["best_len"] = len(path)
            rows.append(ids); paths.append(path)
        n = max(len(r) for r in rows)
        self.id = torch.tensor([r + [-1] * (n - len(r)) for r in rows], dtype=torch.int32)
        m = max(2, max(len(q) for q in paths))
        self.best_path_t = torch.tensor([q + [0] * (m - len(q)) for q in paths], dtype=torch.int32)
        self._rec = rec

    def planners(self):
        return self._rec.copy()

    def set_obstacles(self, obstacles):
        self.obstacles = obstacles

    def select_branch(self, cur):
        out = np.array([r.t.select_branch(int(c)) if c >= 0 else -1 for r, c in zip(self.reps, cur)], np.int32)
        self._refresh()
        return out

    def valid_path(self, prev):
        sep, err = np.full(self.P, -1, np.int32), np.zeros(self.P, np.int32)
        for p, (r, c) in enumerate(zip(self.reps, prev)):
            if c >= 0:
                s_, e_ = r.t.valid_path(self.obstacles, self.w["node_radius"], int(c))
                sep[p], err[p] = (-1 if e_ else s_), int(e_)
        self._refresh()
        return sep, err

    def cut_blocked_edges(self, sep):
        out, cuts = np.array(sep, np.int32).copy(), np.zeros(self.P, np.int32)
        for p, (r, s_) in enumerate(zip(self.reps, sep)):
            if s_ >= 0:
                out[p], cuts[p] = r.t.cut_blocked_edges(self.obstacles, int(s_))
        self._refresh()
        return out, cuts

    def reconnect(self, ids, radius):
        out = np.array([r.t.reconnect(self.obstacles, int(i), radius, int(r.t.best_path()[0])) if i >= 0 else -1
                        for r, i in zip(self.reps, ids)], np.int32)
        self._refresh()
        return out

    def regrow(self, ids, max_iter):
        res, its = np.full(self.P, S.E_BADARG if hasattr(S, "E_BADARG") else -100, np.int32), np.zeros(self.P, np.int32)
        for p, (r, i) in enumerate(zip(self.reps, ids)):
            if i >= 0:
                res[p], its[p] = r.t.regrow(r.prm, self.obstacles, r.goal, int(i), seed=r.seed, stream_id=r.stream_id, max_iter=max_iter)
        self._refresh()
        return res, its


@pytest.mark.parametrize("edges", [False, True], ids=["nodes", "nodes+edges"])
def test_product_loop_host_logic_equals_the_oracle_loop_on_oracle_trees(edges):
    from rrt_star_fnd_algorithm_b200.fnd_loop import FNDReplanner
    w = small_c5()
    P, T = 10, w["ticks"]
    fake = OracleBackedBatch(w, P, T)
    loop = FNDReplanner(fake, w["reconnect_radius"], w["regrow_iters"], edge_invalidation=edges)
    reps = oracle_planners(w, P, T, edges)
    assert list(loop.phase) == [r.phase for r in reps]
    for k in range(1, T + 1):
        ob = moved_obstacles(w["obstacles"], w["velocity"], k)
        got = loop.tick(ob)
        for p, r in enumerate(reps):
            want = r.tick(ob)
            what = f"tick {k}, planner {p}"
            assert int(got["current"][p]) == want["current"] and int(got["phase"][p]) == want["phase"], what
            if want["current"] < 0:
                continue
            assert int(got["status"][p]) == want["status"] and int(got["separate_root"][p]) == want["separate_root"], what
            assert list(got["linked"][p]) == want["linked"] and list(got["regrow"][p]) == want["regrow"], what
            assert list(got["regrow_iters"][p]) == want["regrow_iters"], what
            assert_tree_equal(fake.reps[p].t.export(), r.t.export(), what)
    assert loop.stats["splits"] >= 15 and loop.stats["regrown"] >= 2 and loop.stats["moves"] > 5 * P
    if edges:
        assert loop.stats["edge_cuts"] >= 3, loop.stats       # the extension really cut something the node test missed


def test_edge_invalidation_cuts_what_the_node_test_misses():
    """An obstacle dropped BETWEEN two path nodes: valid_path (nodes only, src/algorithm.py:356) sees nothing, the edge
    extension cuts that edge and returns the goal-side node as the separate root."""
    case = K.case_by_name("float_star")
    prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
    t = O.OracleTree(case["start"], case["iter_num"] + 2)
    assert t.run(prm, case["obstacles"], case["goal"], seed=case["seed"], stream_id=case["stream_id"])[0] == S.ST_DONE
    path = [int(i) for i in t.best_path()]
    e = t.export()
    pos = {int(i): (e["x"][k], e["y"][k]) for k, i in enumerate(e["ids"])}
    k = len(path) // 2
    a, b = pos[path[k]], pos[path[k + 1]]
    mid = (0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1]), 0.05)            # a small circle on the edge, clear of both nodes
    obst = list(case["obstacles"]) + [mid]
    sep, err = t.valid_path(obst, 0.0, path[-1])
    assert not err and sep == path[-1]                                # node test: nothing is occupied (node_radius 0 here)
    sep2, cuts = t.cut_blocked_edges(obst, sep)
    assert cuts == 1 and sep2 == path[k]
    e2 = t.export()
    par = dict(zip([int(i) for i in e2["ids"]], [int(q) for q in e2["parent"]]))
    assert par[path[k]] == -1 and len(e2["ids"]) == len(e["ids"])     # cut, nothing deleted


@pytest.mark.gpu
@pytest.mark.parametrize("kind,edges", [("host", False), ("device", False), ("host", True), ("device", True)],
                         ids=["host-loop", "device-loop", "host-loop+edges", "device-loop+edges"])
def test_gpu_fnd_loop_matches_the_oracle_loop_tick_by_tick(kind, edges):
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch, fn_capacity
    from rrt_star_fnd_algorithm_b200.fnd_loop import DeviceFNDReplanner, FNDReplanner
    w = small_c5()
    P, T = 16, w["ticks"]
    s, g, seeds, sids = W.planner_problems(w, 0, P)
    b = PlannerBatch("fn", starts=s, goals=g, obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=w["iter_num"],
                     step_length=w["step_length"], radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"],
                     max_nodes=w["max_nodes"], seeds=seeds, stream_ids=sids).run()
    b.check_status()
    b.grow_capacity(fn_capacity(w["max_nodes"]) + T * w["regrow_iters"])      # room for the regrow() extensions
    reps = oracle_planners(w, P, T, edges)
    for p in range(P):
        assert_tree_equal(b.tree(p), reps[p].t.export(), f"planned[{p}]")
    loop = (DeviceFNDReplanner if kind == "device" else FNDReplanner)(b, w["reconnect_radius"], w["regrow_iters"], edge_invalidation=edges)
    assert list(loop.phase) == [r.phase for r in reps]
    seen = dict(linked=0, regrown=0)
    for k in range(1, T + 1):
        ob = moved_obstacles(w["obstacles"], w["velocity"], k)
        got = loop.tick(ob)
        pl = b.planners()
        for p, r in enumerate(reps):
            was = r.phase
            want = r.tick(ob)
            what = f"tick {k}, planner {p}"
            assert int(got["current"][p]) == want["current"], what
            assert int(got["phase"][p]) == want["phase"], what
            if was != MOVING or want["current"] < 0:
                continue
            assert int(got["status"][p]) == want["status"], what
            assert int(got["index_error"][p]) == want["index_error"] and int(got["separate_root"][p]) == want["separate_root"], what
            assert list(got["linked"][p]) == want["linked"], what
            assert list(got["regrow"][p]) == want["regrow"] and list(got["regrow_iters"][p]) == want["regrow_iters"], what
            seen["linked"] += sum(v >= 0 for v in want["linked"])
            seen["regrown"] += sum(v == 1 for v in want["regrow"])
            assert_tree_equal(b.tree(p), r.t.export(), what)
            op = r.t.planner()
            assert int(pl[p]["cursor"]) == op.cursor and int(pl[p]["last_id"]) == op.last_id and int(pl[p]["n_live"]) == op.n_live, what
            if want["phase"] != LOST:
                assert np.array_equal(b.best_path(p), r.t.best_path()), what
                assert pl[p]["best_cost"] == op.best_cost, what
    assert seen["linked"] >= 10 and seen["regrown"] >= 3, seen
    assert loop.stats["splits"] >= 20 and loop.stats["reconnected"] == seen["linked"] and loop.stats["regrown"] == seen["regrown"]


# ---- the composition against the UNMODIFIED reference functions (oracle/make_fndloop_golden.py) ------------------------

from oracle.make_fndloop_golden import LOOP_CASES, REGROW_ITERS  # noqa: E402
from oracle.fnd_loop import MAX_ROUNDS  # noqa: E402
from util import load_npz  # noqa: E402


def golden_tree(g, k, regrown):
    t = {key: g[f"t{k}_{key}"] for key in ("ids", "x", "y", "parent", "cost", "nchild")}
    if regrown:      # the reference's children lists are inconsistent after reconnect_ancestors
        ids = list(t["ids"])
        nc = np.zeros(len(ids), np.int32)
        for q in t["parent"]:
            if q >= 0:
                nc[ids.index(int(q))] += 1
        t["nchild"] = nc
    return t


def check_loop_against_golden(lc, g, tick_fn, tree_fn, path_fn, cost_fn, cursor_fn):
    n = int(g["n_ticks"])
    assert n >= 1
    seen = dict(splits=0, linked=0, regrown=0, multi=0)
    for k in range(1, n + 1):
        got = tick_fn(g[f"obst_{k}"])
        what = f"{lc['name']} tick {k}"
        assert got["current"] == int(g[f"t{k}_current"]), what
        assert got["index_error"] == int(g[f"t{k}_index_error"]), what
        lost = int(g[f"t{k}_lost"])
        assert (got["phase"] == LOST) == bool(lost), what
        if not got["index_error"]:
            assert got["separate_root"] == int(g[f"t{k}_separate_root"]), what
        assert list(got["linked"]) == [int(v) for v in g[f"t{k}_linked"]], what
        assert [int(v == 1) for v in got["regrow"]] == [int(v) for v in g[f"t{k}_regrow"]], what
        regrown = bool(g[f"t{k}_regrow"].any())
        assert_tree_equal(tree_fn(), golden_tree(g, k, regrown), what)
        assert cursor_fn() == int(g[f"t{k}_cursor"]), what
        if not lost:
            assert [int(i) for i in path_fn()] == [int(i) for i in g[f"t{k}_path"]], what
            if f"t{k}_path_cost" in g:
                assert cost_fn() == float(g[f"t{k}_path_cost"]), what
        split = got["separate_root"] not in (-1, got["current"])
        seen["splits"] += split; seen["linked"] += sum(v >= 0 for v in got["linked"]); seen["regrown"] += regrown
        seen["multi"] += got["linked"][1] >= 0 or got["regrow"][1] != 0
    return seen


@pytest.mark.parametrize("lc", LOOP_CASES, ids=[c["name"] for c in LOOP_CASES])
def test_oracle_loop_equals_the_reference_functions_composed(lc):
    case, g = K.case_by_name(lc["case"]), load_npz(f"fndloop_{lc['name']}.npz")
    prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
    t = O.OracleTree(case["start"], case["iter_num"] + 2 + (lc["ticks"] + 1) * REGROW_ITERS)
    assert t.run(prm, case["obstacles"], case["goal"], seed=case["seed"], stream_id=case["stream_id"])[0] == S.ST_DONE
    assert np.array_equal(t.best_path(), g["path0"]) and t.planner().cursor == int(g["cursor0"])
    rep = OracleReplanner(t, prm, case["goal"], case["seed"], case["stream_id"], case["node_radius"], lc["radius"], REGROW_ITERS)
    check_loop_against_golden(lc, g, rep.tick, t.export, t.best_path, lambda: t.planner().best_cost, lambda: t.planner().cursor)


def test_loop_goldens_cover_the_interesting_states():
    tot = dict(ticks=0, splits=0, linked=0, regrown=0, multi=0, lost=0)
    for lc in LOOP_CASES:
        g = load_npz(f"fndloop_{lc['name']}.npz")
        n = int(g["n_ticks"])
        tot["ticks"] += n
        for k in range(1, n + 1):
            tot["splits"] += int(g[f"t{k}_separate_root"]) not in (-1, int(g[f"t{k}_current"]))
            tot["linked"] += int((g[f"t{k}_linked"] >= 0).sum()); tot["regrown"] += int(g[f"t{k}_regrow"].sum())
            tot["multi"] += int(g[f"t{k}_linked"][1] >= 0 or g[f"t{k}_regrow"][1] != 0)
        tot["lost"] += int(g[f"t{n}_lost"])
    assert tot["ticks"] >= 40 and tot["splits"] >= 15 and tot["linked"] >= 12 and tot["regrown"] >= 2 and tot["multi"] >= 1, tot


@pytest.mark.gpu
@pytest.mark.parametrize("lc", LOOP_CASES, ids=[c["name"] for c in LOOP_CASES])
def test_gpu_loop_equals_the_reference_functions_composed(lc):
    """The device loop against the goldens recorded from the unmodified reference's own functions."""
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    from rrt_star_fnd_algorithm_b200.fnd_loop import FNDReplanner
    case, g = K.case_by_name(lc["case"]), load_npz(f"fndloop_{lc['name']}.npz")
    b = PlannerBatch(case["mode"], starts=[case["start"]], goals=[case["goal"]], obstacles=case["obstacles"],
                     xy_range=case["xy_range"], iter_num=case["iter_num"], step_length=case["step_length"],
                     radius=case["radius"], node_radius=case["node_radius"], bias=case["bias"], max_nodes=case["max_nodes"],
                     seeds=[case["seed"]], stream_ids=[case["stream_id"]]).run()
    b.check_status()
    assert np.array_equal(b.best_path(0), g["path0"]) and int(b.planners()[0]["cursor"]) == int(g["cursor0"])
    b.grow_capacity(b.prm.capacity + (int(g["n_ticks"]) + 1) * REGROW_ITERS)
    loop = FNDReplanner(b, lc["radius"], REGROW_ITERS)

    def tick(obst):
        out = loop.tick([(float(r[0]), float(r[1]), float(r[2])) for r in obst])
        return dict(current=int(out["current"][0]), index_error=int(out["index_error"][0]), phase=int(out["phase"][0]),
                    separate_root=int(out["separate_root"][0]), linked=[int(v) for v in out["linked"][0]],
                    regrow=[int(v) for v in out["regrow"][0]])

    check_loop_against_golden(lc, g, tick, lambda: b.tree(0), lambda: b.best_path(0),
                              lambda: b.planners()[0]["best_cost"], lambda: int(b.planners()[0]["cursor"]))
