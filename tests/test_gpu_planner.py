"""GPU: the fused planner kernel (through the C ABI) against the golden vectors of the reference and the oracle."""
import numpy as np
import pytest

import oracle as O
from oracle import cases as K
from oracle.philox import philox_words
from rrt_star_fnd_algorithm_b200 import _structs as S

from util import assert_trace_equal, assert_tree_equal, load_golden

pytestmark = pytest.mark.gpu
CASES = [c["name"] for c in K.all_cases()]


def make_batch(case, *, words=None, seeds=None, stream_ids=None, starts=None, goals=None, trace=True, **kw):
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    P = 1 if starts is None else len(starts)
    return PlannerBatch(
        case["mode"], starts=[case["start"]] * P if starts is None else starts,
        goals=[case["goal"]] * P if goals is None else goals, obstacles=case["obstacles"], xy_range=case["xy_range"],
        iter_num=case["iter_num"], step_length=case["step_length"], radius=case["radius"],
        node_radius=case["node_radius"], bias=case["bias"], max_nodes=case["max_nodes"], words=words, seeds=seeds,
        stream_ids=stream_ids, trace=trace, **kw)


def check_against_golden(b, case, g, p=0):
    pl = b.planners()[p]
    assert pl["status"] == S.ST_DONE, S.STATUS_NAMES.get(int(pl["status"]))
    assert_tree_equal(b.tree(p), g, case["name"])
    assert pl["iter"] == int(g["iter"])
    assert pl["cursor"] == int(g["words_used"])
    assert pl["attempts"] == int(g["attempts"])
    assert pl["last_id"] == int(g["last_id"])
    assert pl["n_edge_checks_ref"] == int(g["edge_calls"])
    assert np.array_equal(b.best_path(p), g["best_path"])
    assert pl["best_cost"] == float(g["best_cost"])
    n = int(pl["iter"])
    tr = b.trace(p)
    assert_trace_equal(tr, g, n, case["mode"], case["name"])
    if case["mode"] != "rrt":
        assert np.array_equal(tr["cp_parent"][:n], g["rec_cp_parent"][:n])
        assert np.array_equal(tr["cp_cost"][:n], g["rec_cp_cost"][:n])


@pytest.mark.parametrize("name", CASES)
def test_planner_word_buffer_matches_reference_golden(name):
    case, g = K.case_by_name(name), load_golden(name)
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    b = make_batch(case, words=words[None, :]).run()
    check_against_golden(b, case, g)


@pytest.mark.parametrize("name", CASES)
def test_planner_philox_matches_reference_golden(name):
    case, g = K.case_by_name(name), load_golden(name)
    b = make_batch(case, seeds=[case["seed"]], stream_ids=[case["stream_id"]]).run()
    check_against_golden(b, case, g)


@pytest.mark.parametrize("name", ["float_fn", "float_star", "main_rrt_nobias"])
def test_batched_planners_match_oracle(name):
    """64 planners with their own start/goal and stream: every tree equals the CPU oracle's."""
    case = K.case_by_name(name)
    rng = np.random.default_rng(5)
    P = 64
    (x0, x1), (y0, y1) = case["xy_range"]
    starts, goals = [], []
    while len(starts) < P:
        s = (rng.uniform(x0, x1), rng.uniform(y0, y1)); t = (rng.uniform(x0, x1), rng.uniform(y0, y1))
        if O.is_occupied(s, case["obstacles"], case["node_radius"]) or O.is_occupied(t, case["obstacles"], case["node_radius"]):
            continue
        starts.append(s); goals.append(t)
    it = 300 if case["mode"] != "rrt" else case["iter_num"]
    case = dict(case, iter_num=it)
    b = make_batch(case, seeds=np.full(P, 77), stream_ids=np.arange(P), starts=starts, goals=goals, trace=False).run()
    pls = b.planners()
    for p in range(P):
        prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
        t = O.OracleTree(starts[p], it + 2)
        st, _ = t.run(prm, case["obstacles"], goals[p], seed=77, stream_id=p)
        assert pls[p]["status"] == st
        assert_tree_equal(b.tree(p), t.export(), f"{name}[{p}]")
        opl = t.planner()
        for f in ("iter", "cursor", "attempts", "best_cost", "n_live", "n_rewired", "n_removed", "n_edge_checks",
                  "n_edge_checks_ref", "solution_found"):
            assert pls[p][f] == getattr(opl, f), (p, f)
        assert np.array_equal(b.best_path(p), t.best_path())


def test_tiny_capacity_forces_compactions():
    """FN with the minimum slot capacity compacts constantly; results must not change."""
    case, g = K.case_by_name("float_fn"), load_golden("float_fn")
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    b = make_batch(case, words=words[None, :], capacity=case["max_nodes"] + 8).run()
    check_against_golden(b, case, g)
    assert b.planners()[0]["n_compactions"] > 50


def test_resume_after_need_words():
    case, g = K.case_by_name("main_fn"), load_golden("main_fn")
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    n, calls = 600, 0
    b = make_batch(case, words=words[None, :n])
    import torch
    while True:
        b.run(); calls += 1
        st = int(b.planners()[0]["status"])
        if st != S.ST_NEED_WORDS:
            break
        n += 900
        b.words = torch.from_numpy(np.ascontiguousarray(words[None, :n]).view(np.int32)).to(b.device)
        b.prm.words_per_planner = n
    assert calls > 3
    check_against_golden(b, case, g)


@pytest.mark.parametrize("launch", [1 + 256 * 16, 2 + 16 * 2 + 256 * 32, 4 + 16 * 4 + 256 * 24, 1 + 16 * 2 + 256 * 32])
@pytest.mark.parametrize("grid", [True, False])
@pytest.mark.parametrize("name", ["dense_fn", "int_star"])
def test_every_launch_geometry_gives_the_same_tree(name, launch, grid):
    """planners per CTA and obstacle culling are performance knobs only (3 planners: a partially filled last CTA)."""
    case, g = K.case_by_name(name), load_golden(name)
    b = make_batch(case, seeds=[case["seed"]] * 3, stream_ids=[case["stream_id"]] * 3, starts=[case["start"]] * 3,
                   goals=[case["goal"]] * 3, launch=launch, grid=grid, near_cap=32).run()
    for p in range(3):
        check_against_golden(b, case, g, p)


def test_grid_cell_size_does_not_matter():
    case, g = K.case_by_name("dense_fn"), load_golden("dense_fn")
    for cell in (0.7, 3.0, 11.0, 250.0):
        b = make_batch(case, seeds=[case["seed"]], stream_ids=[case["stream_id"]], grid_cell=cell).run()
        check_against_golden(b, case, g)


@pytest.mark.parametrize("name", ["main_star", "float_fn"])
def test_rewire_list_overflow_path(name):
    """near_cap only sizes the shared-memory list of re-parented nodes; the overflow path gives the same tree."""
    case, g = K.case_by_name(name), load_golden(name)
    assert int(g["rec_n_rewired"].max()) >= 2
    b = make_batch(case, seeds=[case["seed"]], stream_ids=[case["stream_id"]], near_cap=1).run()
    check_against_golden(b, case, g)


@pytest.mark.parametrize("name,cell", [("main_star_long", 0.0), ("float_star", 1.5), ("int_star", 40.0), ("main_rrt", 30.0),
                                       ("float_star", 0.4)])
def test_node_grid_gives_the_same_tree_as_the_scans(name, cell):
    """Config 4's node grid answers nearest / near-set queries from a window of cells: identical trees, records, counts."""
    case, g = K.case_by_name(name), load_golden(name)
    b = make_batch(case, seeds=[case["seed"]] * 2, stream_ids=[case["stream_id"]] * 2, starts=[case["start"]] * 2,
                   goals=[case["goal"]] * 2, node_grid=True, node_grid_cell=cell, node_grid_min_nodes=8,
                   node_grid_cand_cap=64 if cell == 0.4 else 4096).run()     # the tiny cap forces fallbacks to the scan
    for p in range(2):
        check_against_golden(b, case, g, p)


def test_node_grid_large_tree_matches_the_oracle():
    """40,000 nodes (RRTFND_LONG_TESTS=1: 200,000 - minutes of brute-force oracle) in a 400 x 400 map, 60 obstacles: grid from 256
    nodes on, walk records with grandparent hints, every node equal to the CPU oracle's."""
    import os
    import oracle as O
    n_nodes = 200000 if os.environ.get("RRTFND_LONG_TESTS") == "1" else 40000
    rng = np.random.default_rng(8)
    obst = [(float(x), float(y), float(r)) for x, y, r in zip(rng.uniform(20, 380, 60), rng.uniform(20, 380, 60), rng.uniform(3, 12, 60))]
    case = dict(name="big", mode="star", seed=21, stream_id=3, start=(5.0, 5.0), goal=(390.0, 390.0), xy_range=[[0, 400], [0, 400]],
                obstacles=obst, iter_num=n_nodes, step_length=2.0, radius=4.0, node_radius=1.0, bias=0.01, max_nodes=0)
    b = make_batch(case, seeds=[21], stream_ids=[3], trace=False, node_grid=True, node_grid_min_nodes=256,
                   finish_cap=32768, path_cap=32768).run()
    b.check_status()
    t = O.OracleTree(case["start"], case["iter_num"] + 2, finish_cap=32768, path_cap=32768)
    prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0, finish_cap=32768, path_cap=32768)
    st, _ = t.run(prm, obst, case["goal"], seed=21, stream_id=3)
    assert st == S.ST_DONE
    assert_tree_equal(b.tree(0), t.export(), "big")
    pl, opl = b.planners()[0], t.planner()
    for f in ("cursor", "attempts", "best_cost", "n_rewired", "n_edge_checks", "n_edge_checks_ref"):
        assert pl[f] == getattr(opl, f), f
    assert pl["n_scan_nodes"] < 0.05 * opl.n_scan_nodes           # the point of the grid


def test_batch_planner_as_a_graph_object():
    """PlannerBatch.graph(p): the reference's Graph dictionaries built on demand from the device arrays."""
    from rrt_star_fnd_algorithm_b200.algorithm import find_path
    case, g = K.case_by_name("float_star"), load_golden("float_star")
    b = make_batch(case, seeds=[case["seed"]], stream_ids=[case["stream_id"]], trace=False).run()
    G = b.graph(0)
    assert sorted(G.vertices) == [int(i) for i in g["ids"]] and G.last_id == int(g["last_id"])
    assert G.start == (case["start"][0], case["start"][1]) and G.parent[0] is None
    path, cost = find_path(G, int(g["best_path"][0]), 0)
    assert path == [int(i) for i in g["best_path"]] and cost == float(g["best_cost"])
    assert G.get_cost(int(g["best_path"][0])) == float(g["best_cost"])
    assert sum(len(c) for c in G.children.values()) == len(G.vertices) - 1


@pytest.mark.parametrize("name,slices", [("float_star", 5), ("float_fn", 3), ("float_star", 15)])
def test_sliced_overlapping_launches_give_the_same_trees(name, slices, monkeypatch):
    """Large batches are grown in slices whose launches overlap (programmatic dependent launch; a planner's slices are ordered
    by a ticket in its record, csrc/planner.cu launch_plan). Forced here on a small batch, where every slice's warps are
    resident at once and really wait for their tickets: same trees, records and best paths as the oracle."""
    from rrt_star_fnd_algorithm_b200 import _abi
    import ctypes as C
    case = dict(K.case_by_name(name), iter_num=400)
    rng = np.random.default_rng(9)
    P = 96
    (x0, x1), (y0, y1) = case["xy_range"]
    starts, goals = [], []
    while len(starts) < P:
        s_ = (rng.uniform(x0, x1), rng.uniform(y0, y1)); t_ = (rng.uniform(x0, x1), rng.uniform(y0, y1))
        if O.is_occupied(s_, case["obstacles"], case["node_radius"]) or O.is_occupied(t_, case["obstacles"], case["node_radius"]):
            continue
        starts.append(s_); goals.append(t_)
    monkeypatch.setenv("RRTFND_SLICES", str(slices))
    b = make_batch(case, seeds=np.full(P, 31), stream_ids=np.arange(P), starts=starts, goals=goals, trace=False)
    assert _abi.lib.rrtfnd_plan_launches(C.byref(b.prm), P) == slices
    b.run()
    b.run()                                   # a second call on finished planners: every slice is a no-op, tickets move on
    pls = b.planners()
    for p in range(P):
        prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
        t = O.OracleTree(starts[p], case["iter_num"] + 2)
        st, _ = t.run(prm, case["obstacles"], goals[p], seed=31, stream_id=p)
        assert pls[p]["status"] == st == S.ST_DONE
        assert_tree_equal(b.tree(p), t.export(), f"{name}[{p}] in {slices} slices")
        opl = t.planner()
        for f in ("iter", "cursor", "attempts", "best_cost", "n_live", "n_rewired", "n_removed", "n_edge_checks", "n_edge_checks_ref"):
            assert pls[p][f] == getattr(opl, f), (p, f)
        assert np.array_equal(b.best_path(p), t.best_path())
