"""GPU: the paths that carry the headline numbers, at the headline configurations, against the CPU oracle.

* `rrtfnd_plan_host` — the host-buffer entry point bench.py's `e2e` times and INTEGRATION.md binds — on the c3 and c2
  workloads, full length, planners spread over the whole index range;
* the DEFAULT launch of `rrtfnd_plan_batch` (launch word 0: lean instance, TMA ring when capacity > 2048, plain loads
  otherwise) on the c3, c2 and c5 workloads exactly as bench.py constructs them (reference: src/algorithm.py:97-265);
* the stand-alone operators `rrtfnd_nearest_visible` (src/algorithm.py:666-699), `rrtfnd_near_set` (:837, :893) and
  `rrtfnd_chain_cost` (src/graph.py:34-46) against the oracle's `orc_nearest_visible` / `orc_near_set` / `orc_get_cost`.

Everything is compared bit for bit: trees (ids, coordinates, parents, edge costs, child counts), planner records,
best paths and costs.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import pytest

import oracle as O
from rrt_star_fnd_algorithm_b200 import _structs as S
from rrt_star_fnd_algorithm_b200 import workloads as W

from util import assert_tree_equal

RECORD_FIELDS = ("status", "iter", "cursor", "attempts", "best_cost", "n_live", "last_id", "n_rewired", "n_removed",
                 "n_edge_checks", "n_edge_checks_ref", "solution_found", "best_len")
N_THREADS = max(1, min(32, len(os.sched_getaffinity(0))))


def oracle_params(w, flags=0):
    p = S.Params()
    p.mode = {"rrt": 0, "star": 1, "fn": 2}[w["mode"]]
    p.iter_num, p.max_nodes, p.capacity = w["iter_num"], w["max_nodes"], w["iter_num"] + 2
    p.n_obst, p.near_cap, p.finish_cap, p.path_cap = len(w["obstacles"]), 4096, 4096, 4096
    p.stream_mode, p.flags = S.STREAM_PHILOX, flags
    p.step_length, p.radius, p.node_radius, p.bias = w["step_length"], w["radius"], w["node_radius"], w["bias"]
    (p.x_lo, p.x_hi), (p.y_lo, p.y_hi) = w["xy_range"]
    p.goal_node_radius = w["node_radius"]
    return p


def spread_problems(w, total, count):
    """`count` planners of the workload spread evenly over global indices 0 .. total-1 (first and last included)."""
    idx = np.unique(np.linspace(0, total - 1, count).astype(np.int64))
    rows = [W.planner_problems(w, int(g), 1) for g in idx]
    starts, goals = np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows])
    seeds, sids = np.concatenate([r[2] for r in rows]), np.concatenate([r[3] for r in rows])
    return idx, starts, goals, seeds, sids


def oracle_run(w, starts, goals, seeds, sids, extra_ids=0):
    """The CPU oracle on every problem (threads; ctypes releases the GIL). Returns the OracleTree objects."""
    def one(p):
        t = O.OracleTree(starts[p], w["iter_num"] + 2 + extra_ids)
        t.run(oracle_params(w), w["obstacles"], goals[p], seed=int(seeds[p]), stream_id=int(sids[p]))
        return t
    with ThreadPoolExecutor(N_THREADS) as ex:
        return list(ex.map(one, range(len(starts))))


def assert_record_equal(pl, opl, what):
    for f in RECORD_FIELDS:
        got, want = pl[f], getattr(opl, f)
        assert got == want, f"{what}: planner record field {f}: got {got}, want {want}"


def canonical_tree(ids, xy, ecost, parent_slot, nchild, n_slots):
    ids, par = ids[:n_slots], parent_slot[:n_slots]
    live = ids >= 0
    parent = np.where(par >= 0, ids[np.clip(par, 0, None)], -1).astype(np.int32)
    return dict(ids=ids[live].copy(), x=xy[:n_slots, 0][live].copy(), y=xy[:n_slots, 1][live].copy(), parent=parent[live],
                cost=ecost[:n_slots][live].copy(), nchild=nchild[:n_slots][live].copy())


# ---------------------------------------------------------------------------------------------- rrtfnd_plan_host

@pytest.mark.gpu
@pytest.mark.parametrize("name,total,count", [("c3", 16384, 32), ("c2", 4096, 32)])
def test_plan_host_matches_oracle_at_full_length(name, total, count):
    """Host buffers in, host buffers out, through the C ABI call bench.py's e2e leg times; [M][3] obstacle rows, so the
    per-circle constant K is evaluated inside the library (host pow()), not by the Python host."""
    from rrt_star_fnd_algorithm_b200 import _abi
    from rrt_star_fnd_algorithm_b200.batch import fn_capacity, obstacle_table
    w = W.WORKLOADS[name]()
    idx, starts, goals, seeds, sids = spread_problems(w, total, count)
    P = len(idx)
    prm = oracle_params(w, S.FLAG_EVAL_CHOOSE_PARENT)
    prm.capacity = fn_capacity(w["max_nodes"]) if w["mode"] == "fn" else w["iter_num"] + 2     # as PlannerBatch / bench.py size it
    prm.near_cap, prm.finish_cap, prm.path_cap = 64, 1024, 1024
    Cn = prm.capacity
    sg = np.ascontiguousarray(np.concatenate([starts, goals], 1))
    sd = np.ascontiguousarray(np.stack([seeds, sids], 1).astype(np.uint64))
    ob3 = np.ascontiguousarray(np.asarray(w["obstacles"], np.float64).reshape(-1, 3))
    o_pl = np.zeros(P, dtype=S.PLANNER_DTYPE)
    o_xy, o_e = np.zeros((P, Cn, 2)), np.zeros((P, Cn))
    o_par, o_nc, o_id = np.zeros((P, Cn), np.int32), np.zeros((P, Cn), np.int32), np.zeros((P, Cn), np.int32)
    o_bp = np.zeros((P, prm.path_cap), np.int32)
    rc = _abi.lib.rrtfnd_plan_host(C.byref(prm), P, sg.ctypes.data, sd.ctypes.data, ob3.ctypes.data, None, o_pl.ctypes.data,
                                   o_xy.ctypes.data, o_e.ctypes.data, o_par.ctypes.data, o_nc.ctypes.data, o_id.ctypes.data,
                                   o_bp.ctypes.data, 0)
    _abi.check(rc, "rrtfnd_plan_host")
    trees = oracle_run(w, starts, goals, seeds, sids)
    for p, t in enumerate(trees):
        what = f"{name} planner {int(idx[p])} (plan_host)"
        assert_record_equal(o_pl[p], t.planner(), what)
        assert_tree_equal(canonical_tree(o_id[p], o_xy[p], o_e[p], o_par[p], o_nc[p], int(o_pl[p]["n_slots"])), t.export(), what)
        assert np.array_equal(o_bp[p, :int(o_pl[p]["best_len"])], t.best_path()), what

    # the caller's own K (RRTFND_FLAG_OBST_HAS_K): same results, and the table the library derives from [M][3] rows
    # equals the one the Python host evaluates with the reference's `**` expression
    tab4 = obstacle_table(w["obstacles"])
    prm4 = oracle_params(w, S.FLAG_EVAL_CHOOSE_PARENT | S.FLAG_OBST_HAS_K)
    prm4.capacity, prm4.near_cap, prm4.finish_cap, prm4.path_cap = Cn, 64, 1024, 1024
    q_pl = np.zeros(P, dtype=S.PLANNER_DTYPE)
    q_xy = np.zeros((P, Cn, 2))
    rc = _abi.lib.rrtfnd_plan_host(C.byref(prm4), P, sg.ctypes.data, sd.ctypes.data, tab4.ctypes.data, None, q_pl.ctypes.data,
                                   q_xy.ctypes.data, None, None, None, None, None, 0)
    _abi.check(rc, "rrtfnd_plan_host (OBST_HAS_K)")
    assert np.array_equal(q_pl.view(np.uint8), o_pl.view(np.uint8))
    for p in range(P):
        n = int(o_pl[p]["n_slots"])
        assert np.array_equal(q_xy[p, :n].view(np.uint64), o_xy[p, :n].view(np.uint64))


def test_host_pow_table_equals_the_python_expression():
    """K from C pow() (what rrtfnd_plan_host evaluates for [M][3] rows) == the reference's `-r ** 2 + x0 ** 2 + y0 ** 2` on
    Python floats, for every circle of the bench workloads; and it is NOT always -(r*r) + x0*x0 + y0*y0."""
    libm = C.CDLL("libm.so.6")
    libm.pow.restype, libm.pow.argtypes = C.c_double, [C.c_double, C.c_double]
    differs = 0
    for name in ("c2", "c3", "c4", "c5"):
        for (x0, y0, r) in W.WORKLOADS[name]()["obstacles"]:
            want = -r ** 2 + x0 ** 2 + y0 ** 2
            got = (-libm.pow(r, 2.0) + libm.pow(x0, 2.0)) + libm.pow(y0, 2.0)
            assert got == want
            differs += (-(r * r) + x0 * x0) + y0 * y0 != want
    assert differs >= 1          # the reason plan_host must not square by multiplication


# ---------------------------------------------------------------------------------------------- default bench launch

def make_workload_batch(w, starts, goals, seeds, sids, **kw):
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    return PlannerBatch(w["mode"], starts=starts, goals=goals, seeds=seeds, stream_ids=sids, obstacles=w["obstacles"],
                        xy_range=w["xy_range"], iter_num=w["iter_num"], step_length=w["step_length"], radius=w["radius"],
                        node_radius=w["node_radius"], bias=w["bias"], max_nodes=w["max_nodes"], near_cap=64, launch=0, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name,total,count,slices", [("c3", 16384, 64, 0), ("c2", 4096, 64, 0), ("c3", 16384, 64, 8), ("c2", 4096, 64, 8)],
                         ids=["c3-one-wave", "c2-one-wave", "c3-as-benchmarked", "c2-as-benchmarked"])
def test_default_bench_launch_matches_oracle_at_full_length(name, total, count, slices, monkeypatch):
    """The kernel instances bench.py runs (launch word 0 at the workload's own capacity), 5,000 iterations per planner. A batch
    that fits the warp slots takes one launch of the single-wave instance (what a rank runs when config 3 is sharded over 8
    GPUs); `slices = 8` forces what the library does for the benchmarked 16,384- / 4,096-planner batches: eight overlapped slice
    launches of the multi-wave instance (circles of a grid row in list order)."""
    import ctypes as C
    from rrt_star_fnd_algorithm_b200 import _abi
    w = W.WORKLOADS[name]()
    idx, starts, goals, seeds, sids = spread_problems(w, total, count)
    if slices:
        monkeypatch.setenv("RRTFND_SLICES", str(slices))
    b = make_workload_batch(w, starts, goals, seeds, sids)
    assert _abi.lib.rrtfnd_plan_launches(C.byref(b.prm), len(idx)) == max(1, slices)
    b.run()
    assert (b.prm.capacity > 2048) == (name == "c3")         # c3 takes the TMA ring, c2 the plain-load scans
    pls = b.planners()
    trees = oracle_run(w, starts, goals, seeds, sids)
    for p, t in enumerate(trees):
        what = f"{name} planner {int(idx[p])} (default launch)"
        assert_record_equal(pls[p], t.planner(), what)
        assert_tree_equal(b.tree(p), t.export(), what)
        assert np.array_equal(b.best_path(p), t.best_path()), what


@pytest.mark.gpu
def test_default_bench_launch_c5_planning_and_ticks_match_oracle():
    """Config 5 as bench.py runs it: RRT*-FN planning with the default launch, then 24 replanning ticks through the device-side
    loop (rrtfnd_fnd_tick) with node AND edge invalidation."""
    from oracle.fnd_loop import OracleReplanner
    from rrt_star_fnd_algorithm_b200.fnd_loop import DeviceFNDReplanner, LOST
    w = W.WORKLOADS["c5"]()
    idx, starts, goals, seeds, sids = spread_problems(w, 4096, 48)
    b = make_workload_batch(w, starts, goals, seeds, sids).run()
    extra = w["ticks"] * w["regrow_iters"]
    trees = oracle_run(w, starts, goals, seeds, sids, extra_ids=extra)
    pls = b.planners()
    for p, t in enumerate(trees):
        assert_record_equal(pls[p], t.planner(), f"c5 planner {int(idx[p])} (planning)")
        assert_tree_equal(b.tree(p), t.export(), f"c5 planner {int(idx[p])} (planning)")
    b.grow_capacity(b.prm.capacity + extra)
    loop = DeviceFNDReplanner(b, w["reconnect_radius"], w["regrow_iters"], edge_invalidation=w["edge_invalidation"])
    reps = [OracleReplanner(t, oracle_params(w), goals[p], seeds[p], sids[p], w["node_radius"], w["reconnect_radius"],
                            w["regrow_iters"], edge_invalidation=w["edge_invalidation"]) for p, t in enumerate(trees)]
    for k in range(1, w["ticks"] + 1):
        ob = W.moved_obstacles(w["obstacles"], w["velocity"], k)
        got = loop.tick(ob)
        with ThreadPoolExecutor(N_THREADS) as ex:
            wants = list(ex.map(lambda r: r.tick(ob), reps))
        pls = b.planners()
        for p, (r, want) in enumerate(zip(reps, wants)):
            what = f"c5 planner {int(idx[p])} tick {k}"
            assert int(got["current"][p]) == want["current"] and int(got["phase"][p]) == want["phase"], what
            if want["current"] < 0:
                continue
            assert list(got["linked"][p]) == want["linked"] and list(got["regrow"][p]) == want["regrow"], what
            assert list(got["regrow_iters"][p]) == want["regrow_iters"], what
            assert_tree_equal(b.tree(p), r.t.export(), what)
            op = r.t.planner()
            assert int(pls[p]["cursor"]) == op.cursor and int(pls[p]["last_id"]) == op.last_id, what
            if want["phase"] != LOST:
                assert np.array_equal(b.best_path(p), r.t.best_path()) and pls[p]["best_cost"] == op.best_cost, what
    assert loop.stats["splits"] >= 5 and loop.stats["moves"] >= 200, loop.stats


# ---------------------------------------------------------------------------------------------- stand-alone operators

@pytest.mark.gpu
def test_nearest_visible_near_set_chain_cost_match_oracle():
    """rrtfnd_nearest_visible / rrtfnd_near_set / rrtfnd_chain_cost on grown c3-like trees (RRT*, so no tombstones: slot ==
    id) against orc_nearest_visible / orc_near_set / orc_get_cost: blocked nearest nodes (rounds > 1), exact vertex hits
    (-2), queries nobody sees (-1), balls larger than near_cap."""
    import torch
    from rrt_star_fnd_algorithm_b200 import _abi
    w = dict(W.WORKLOADS["c3"](), iter_num=1200)
    idx, starts, goals, seeds, sids = spread_problems(w, 16384, 12)
    P = len(idx)
    b = make_workload_batch(w, starts, goals, seeds, sids).run()
    b.check_status()
    trees = oracle_run(w, starts, goals, seeds, sids)
    dev = b.device
    rng = np.random.default_rng(12)
    near_cap = 48
    rounds_seen, special = 0, {"exact": 0, "none": 0, "overflow": 0}
    for rep in range(40):
        q = rng.uniform(0, 200, (P, 2))
        if rep % 8 == 0:                          # a query that coincides with a vertex (src/algorithm.py:676-678)
            for p in range(0, P, 2):
                e = trees[p].export()
                k = int(rng.integers(0, len(e["ids"])))
                q[p] = (e["x"][k], e["y"][k])
        if rep % 8 == 1:                          # inside a circle: few or no nodes see it
            ob = np.asarray(w["obstacles"])
            big = ob[np.argsort(-ob[:, 2])[:P]]
            q = big[:, :2] + rng.uniform(-0.05, 0.05, (P, 2))
        radius = float(rng.choice([2.0, 10.0, 25.0, 60.0]))
        qt = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        slot = torch.zeros(P, dtype=torch.int32, device=dev)
        rounds = torch.zeros(P, dtype=torch.int32, device=dev)
        cnt = torch.zeros(P, dtype=torch.int32, device=dev)
        near = torch.full((P, near_cap), -1, dtype=torch.int32, device=dev)
        bt = b._batch()
        st = b._stream()
        _abi.check(_abi.lib.rrtfnd_nearest_visible(C.byref(bt), b.prm.capacity, qt.data_ptr(), b.prm.n_obst, slot.data_ptr(),
                                                   rounds.data_ptr(), st), "rrtfnd_nearest_visible")
        _abi.check(_abi.lib.rrtfnd_near_set(C.byref(bt), b.prm.capacity, qt.data_ptr(), radius, near_cap, near.data_ptr(),
                                            cnt.data_ptr(), st), "rrtfnd_near_set")
        slot, rounds, cnt, near = slot.cpu().numpy(), rounds.cpu().numpy(), cnt.cpu().numpy(), near.cpu().numpy()
        for p in range(P):
            want_id, want_rounds = trees[p].nearest_visible(q[p], w["obstacles"])
            what = f"rep {rep} planner {p} q={tuple(q[p])}"
            assert int(slot[p]) == want_id, what                         # RRT* trees: slot == id; -1 / -2 pass through
            if want_id >= 0:
                assert int(rounds[p]) == want_rounds, what
                rounds_seen = max(rounds_seen, want_rounds)
            special["exact"] += want_id == -2; special["none"] += want_id == -1
            ball = trees[p].near_set(q[p], radius)
            assert int(cnt[p]) == len(ball), what
            assert np.array_equal(near[p, :min(len(ball), near_cap)], ball[:near_cap]), what
            special["overflow"] += len(ball) > near_cap
    assert rounds_seen >= 3 and special["exact"] >= 5 and special["overflow"] >= 5, (rounds_seen, special)

    # Graph.get_cost for every node of every tree
    n_slots = b.planners()["n_slots"]
    pi = np.concatenate([np.full(int(n), p, np.int32) for p, n in enumerate(n_slots)])
    sl = np.concatenate([np.arange(int(n), dtype=np.int32) for n in n_slots])
    out = torch.zeros(len(pi), dtype=torch.float64, device=dev)
    pi_t, sl_t = torch.from_numpy(pi).to(dev), torch.from_numpy(sl).to(dev)
    bt = b._batch()
    _abi.check(_abi.lib.rrtfnd_chain_cost(C.byref(bt), b.prm.capacity, pi_t.data_ptr(), sl_t.data_ptr(), len(pi), out.data_ptr(),
                                          b._stream()), "rrtfnd_chain_cost")
    out = out.cpu().numpy()
    want = np.concatenate([[trees[p].get_cost(int(s)) for s in range(int(n))] for p, n in enumerate(n_slots)])
    assert np.array_equal(out.view(np.uint64), want.view(np.uint64))
