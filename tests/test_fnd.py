"""FND tree surgery (select_branch / valid_path / reconnect, src/algorithm.py:294-434).

Goldens (tests/golden/fnd_*.npz) come from the unmodified reference (oracle/make_fnd_golden.py, the call sequence of the
reference's own usage example test_select_branch). CPU: the C oracle against them. GPU: the CUDA operators through the
C ABI against them, batched, and the drop-in functions on Graph objects."""
import numpy as np
import pytest

import oracle as O
from oracle import cases as K
from oracle.make_fnd_golden import FND_CASES
from oracle.philox import philox_words
from rrt_star_fnd_algorithm_b200 import _structs as S

from util import assert_tree_equal, load_npz

IDS = [fc["name"] for fc in FND_CASES]


def stage_tree(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in ("ids", "x", "y", "parent", "cost", "nchild")}


def grown_oracle(case, extra=0):
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    prm = K.params_for(case, stream_mode=S.STREAM_WORDS, n_words=len(words), flags=0)
    t = O.OracleTree(case["start"], case["iter_num"] + extra + 2)
    assert t.run(prm, case["obstacles"], case["goal"], words=words)[0] == S.ST_DONE
    return t


@pytest.mark.parametrize("fc", FND_CASES, ids=IDS)
def test_oracle_fnd_matches_reference_golden(fc):
    case, g = K.case_by_name(fc["case"]), load_npz(f"fnd_{fc['name']}.npz")
    t = grown_oracle(case)
    assert np.array_equal(t.best_path(), g["path0"])
    assert t.select_branch(int(g["current"])) == 0
    assert_tree_equal(t.export(), stage_tree(g, "sb"), "select_branch")
    assert np.array_equal(t.best_path(), g["path1"]) and t.root() == int(g["sb_root"])
    sep, err = t.valid_path(g["obstacles2"], case["node_radius"], int(g["current"]))
    assert err == bool(g["vp_index_error"])
    assert_tree_equal(t.export(), stage_tree(g, "vp"), "valid_path")
    if err:
        return
    assert sep == int(g["vp_sep"])
    linked = t.reconnect(g["obstacles2"], sep, fc["radius"], int(g["path0"][0]))
    assert linked == int(g["rc_linked"])
    assert_tree_equal(t.export(), stage_tree(g, "rc"), "reconnect")
    if linked >= 0:
        assert np.array_equal(t.best_path(), g["rcret_path"]) and t.planner().best_cost == float(g["rcret_cost"])


REGROW = [fc for fc in FND_CASES if fc["name"] in ("float_c", "float_b", "star_b", "int_a")]


def regrow_tree(g):
    """regrow's golden tree: child counts recomputed from the parents (the reference's children lists are inconsistent
    after reconnect_ancestors; parents, positions and edge costs are what is pinned)."""
    t = stage_tree(g, "rg")
    ids = list(t["ids"])
    nc = np.zeros(len(ids), np.int32)
    for p in t["parent"]:
        if p >= 0:
            nc[ids.index(int(p))] += 1
    t["nchild"] = nc
    return t


def regrow_params(case):
    prm = K.params_for(dict(case, bias=0.02), stream_mode=S.STREAM_WORDS, n_words=80000, flags=0)   # test_select_branch passes bias=0.02
    return prm


@pytest.mark.parametrize("fc", REGROW, ids=[f["name"] for f in REGROW])
def test_oracle_regrow_matches_reference_golden(fc):
    case, g = K.case_by_name(fc["case"]), load_npz(f"fnd_{fc['name']}.npz")
    t = grown_oracle(case, extra=600)
    assert t.select_branch(int(g["current"])) == 0
    sep, err = t.valid_path(g["obstacles2"], case["node_radius"], int(g["current"]))
    assert not err and t.reconnect(g["obstacles2"], sep, fc["radius"], int(g["path0"][0])) == -1
    words = philox_words(777, FND_CASES.index(fc), 80000)
    r, iters = t.regrow(regrow_params(case), g["obstacles2"], case["goal"], sep, words=words, cursor=0)
    assert r == int(g["rg_reconnected"]) == 1
    assert t.planner().cursor == int(g["rg_words_used"])
    assert_tree_equal(t.export(), regrow_tree(g), "regrow")


def test_select_branch_rejects_a_node_off_the_path():
    t = grown_oracle(K.case_by_name("float_fn"))
    off = next(int(i) for i in t.export()["ids"][::-1] if int(i) not in set(t.best_path().tolist()))
    assert t.select_branch(off) == -1


# ------------------------------------------------------------------------------------------------------------- GPU
def grown_batch(case, P=1, **kw):
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    b = PlannerBatch(case["mode"], starts=[case["start"]] * P, goals=[case["goal"]] * P, obstacles=case["obstacles"],
                     xy_range=case["xy_range"], iter_num=case["iter_num"], step_length=case["step_length"],
                     radius=case["radius"], node_radius=case["node_radius"], bias=case["bias"], max_nodes=case["max_nodes"],
                     seeds=[case["seed"]] * P, stream_ids=[case["stream_id"]] * P, **kw).run()
    b.check_status()
    return b


@pytest.mark.gpu
@pytest.mark.parametrize("fc", FND_CASES, ids=IDS)
def test_gpu_fnd_matches_reference_golden(fc):
    case, g = K.case_by_name(fc["case"]), load_npz(f"fnd_{fc['name']}.npz")
    P = 3                                                   # identical planners: the operators are batched
    b = grown_batch(case, P)
    for p in range(P):
        assert np.array_equal(b.best_path(p), g["path0"])
    st = b.select_branch([int(g["current"])] * P)
    assert list(st) == [0] * P
    for p in range(P):
        assert_tree_equal(b.tree(p), stage_tree(g, "sb"), f"select_branch[{p}]")
        assert np.array_equal(b.best_path(p), g["path1"])
        pl = b.planners()[p]
        assert b.slot_id(p, int(pl["root_slot"])) == int(g["sb_root"])
        assert (pl["start_x"], pl["start_y"]) == (g["sb_x"][list(g["sb_ids"]).index(int(g["sb_root"]))],
                                                  g["sb_y"][list(g["sb_ids"]).index(int(g["sb_root"]))])
    b.set_obstacles([tuple(r) for r in g["obstacles2"]])
    sep, err = b.valid_path([int(g["current"])] * P)
    for p in range(P):
        assert bool(err[p]) == bool(g["vp_index_error"])
        assert_tree_equal(b.tree(p), stage_tree(g, "vp"), f"valid_path[{p}]")
    if bool(g["vp_index_error"]):
        return
    assert list(sep) == [int(g["vp_sep"])] * P
    linked = b.reconnect(sep, fc["radius"])
    assert list(linked) == [int(g["rc_linked"])] * P
    for p in range(P):
        assert_tree_equal(b.tree(p), stage_tree(g, "rc"), f"reconnect[{p}]")
        if int(g["rc_linked"]) >= 0:
            assert np.array_equal(b.best_path(p), g["rcret_path"])
            assert b.planners()[p]["best_cost"] == float(g["rcret_cost"])


@pytest.mark.gpu
@pytest.mark.parametrize("fc", REGROW, ids=[f["name"] for f in REGROW])
def test_gpu_regrow_matches_reference_golden(fc):
    case, g = K.case_by_name(fc["case"]), load_npz(f"fnd_{fc['name']}.npz")
    P = 2
    b = grown_batch(case, P, capacity=case["iter_num"] + 700)
    cur = [int(g["current"])] * P
    assert list(b.select_branch(cur)) == [0] * P
    b.set_obstacles([tuple(r) for r in g["obstacles2"]])
    sep, err = b.valid_path(cur)
    assert not err.any() and list(b.reconnect(sep, fc["radius"])) == [-1] * P
    words = philox_words(777, FND_CASES.index(fc), 80000)
    b.set_stream_words(np.stack([words] * P))
    b.prm.bias = 0.02                                           # test_select_branch calls regrow with bias=0.02
    r, iters = b.regrow(sep)
    assert list(r) == [1] * P and iters[0] == iters[1] > 0
    for p in range(P):
        assert_tree_equal(b.tree(p), regrow_tree(g), f"regrow[{p}]")
        assert int(b.planners()[p]["cursor"]) == int(g["rg_words_used"])
    # the same against the oracle, including the re-derived best path
    t = grown_oracle(case, extra=700)
    t.select_branch(cur[0]); t.valid_path(g["obstacles2"], case["node_radius"], cur[0])
    t.regrow(regrow_params(case), g["obstacles2"], case["goal"], int(sep[0]), words=words, cursor=0)
    assert np.array_equal(b.best_path(0), t.best_path()) and b.planners()[0]["best_cost"] == t.planner().best_cost


@pytest.mark.gpu
def test_gpu_fnd_then_regrow_with_the_planner():
    """After the surgery the batch is a valid planner state: growing on from it equals the oracle doing the same."""
    fc = next(f for f in FND_CASES if f["name"] == "float_a")
    case, g = K.case_by_name(fc["case"]), load_npz("fnd_float_a.npz")
    b = grown_batch(case, 2)
    t = grown_oracle(case, extra=200)
    cur = int(g["current"])
    assert list(b.select_branch([cur, cur])) == [0, 0] and t.select_branch(cur) == 0
    assert_tree_equal(b.tree(1), t.export(), "after select_branch")
    # 200 more FN iterations from the pruned, re-rooted tree (Philox stream continues at the saved cursor)
    more = dict(case, iter_num=case["iter_num"] + 200)
    b.prm.iter_num = more["iter_num"]
    b.run(); b.check_status()
    prm = K.params_for(more, stream_mode=S.STREAM_PHILOX, flags=0)
    st, _ = t.run(prm, case["obstacles"], case["goal"], seed=case["seed"], stream_id=case["stream_id"])
    assert st == S.ST_DONE
    assert_tree_equal(b.tree(0), t.export(), "regrown")
    assert b.planners()[0]["best_cost"] == t.planner().best_cost


@pytest.mark.gpu
def test_dropin_fnd_functions_on_graph_objects():
    import contextlib, io, random
    from rrt_star_fnd_algorithm_b200 import algorithm as A, graph, map as map_mod
    from oracle.make_dropin_golden import build_objects
    fc = next(f for f in FND_CASES if f["name"] == "star_a")
    case, g = K.case_by_name(fc["case"]), load_npz("fnd_star_a.npz")
    # rebuild the grown tree as a Graph from the oracle, then run the drop-in surgery on it
    t = grown_oracle(case)
    G, m = build_objects(graph, map_mod, case)
    e = t.export()
    G._import(e["ids"], np.column_stack([e["x"], e["y"]]), e["parent"], e["cost"], int(e["ids"].max()), {0: G.start})
    path = [int(i) for i in g["path0"]]
    new_path = A.select_branch(G, int(g["current"]), path, m)
    assert new_path == [int(i) for i in g["path1"]] and G.id_vertex[G.start] == int(g["sb_root"])
    from oracle.make_dropin_golden import graph_arrays
    assert_tree_equal(graph_arrays(G), stage_tree(g, "sb"), "drop-in select_branch")
    m.add_obstacles([[G.vertices[new_path[i]], fc["block_r"]] for i in fc["block"]])
    sep = A.valid_path(G, new_path, m, int(g["current"]))
    assert sep == int(g["vp_sep"])
    assert_tree_equal(graph_arrays(G), stage_tree(g, "vp"), "drop-in valid_path")
    ret = A.reconnect(G, new_path, m, fc["radius"], G.id_vertex[G.start], sep, path[0])
    assert ret is not None and list(ret[0]) == [int(i) for i in g["rcret_path"]] and float(ret[1]) == float(g["rcret_cost"])
    assert_tree_equal(graph_arrays(G), stage_tree(g, "rc"), "drop-in reconnect")


@pytest.mark.gpu
def test_dropin_regrow_refills_the_word_buffer_and_matches_the_oracle():
    """The drop-in regrow() hands the global generator's words to the device in chunks and refills when rejected samples
    eat them: a tiny chunk (many refills) and the default give the same Graph, the same following random.random(), and
    both equal the oracle's regrow fed the Mersenne-Twister words directly."""
    import copy, random
    from rrt_star_fnd_algorithm_b200 import algorithm as A, graph, map as map_mod
    from rrt_star_fnd_algorithm_b200.algorithm import _draw_words
    from oracle.make_dropin_golden import build_objects, graph_arrays
    fc = next(f for f in FND_CASES if f["name"] == "star_a")
    case, g = K.case_by_name(fc["case"]), load_npz("fnd_star_a.npz")
    t = grown_oracle(case)
    G, m = build_objects(graph, map_mod, case)
    e = t.export()
    G._import(e["ids"], np.column_stack([e["x"], e["y"]]), e["parent"], e["cost"], int(e["ids"].max()), {0: G.start})
    path = [int(i) for i in g["path0"]]
    new_path = A.select_branch(G, int(g["current"]), path, m)
    m.add_obstacles([[G.vertices[new_path[i]], fc["block_r"]] for i in fc["block"]])
    sep = A.valid_path(G, new_path, m, int(g["current"]))
    assert sep != int(g["current"])
    # oracle side of the same surgery, then regrow on the MT words of seed 99
    assert t.select_branch(int(g["current"])) == 0
    obst2 = [(float(o[0][0]), float(o[0][1]), float(o[1])) for o in m.obstacles_c]
    wsep, werr = t.valid_path(obst2, case["node_radius"], int(g["current"]))
    assert wsep == sep and not werr
    random.seed(99)
    words = _draw_words(40000)
    prm = regrow_params(case)
    prm.bias = 0.02
    big = O.OracleTree(case["start"], int(e["ids"].max()) + 600)
    ex = t.export()
    big.load(ex["ids"], ex["x"], ex["y"], ex["parent"], ex["cost"], t.root(), int(e["ids"].max()))
    big.set_best_path([path[0]])
    r, iters = big.regrow(prm, obst2, case["goal"], sep, words=words, cursor=0)
    assert r in (0, 1) and iters >= 1
    results = []
    for chunk in (128, 16 * 500 + 256):
        H = copy.deepcopy(G)
        random.seed(99)
        A.regrow(H, m, case["step_length"], case["radius"], H.id_vertex[H.start], sep, path[0], 0.02, _n_words=chunk)
        results.append((graph_arrays(H), random.random()))
    random.seed(99)
    random.getrandbits(32 * int(big.planner().cursor))
    want_next = random.random()
    for arr, nxt in results:
        assert nxt == want_next
        got = dict(arr); want = big.export()
        nc = np.zeros(len(want["ids"]), np.int32)            # children lists after regrow are inconsistent in the reference (B14)
        for k in ("ids", "parent"):
            assert np.array_equal(got[k], want[k]), k
        for k in ("x", "y", "cost"):
            assert np.array_equal(np.asarray(got[k]).view(np.uint64), want[k].view(np.uint64)), k
    assert int(big.planner().cursor) > 128                   # the small chunk really was refilled


@pytest.mark.gpu
def test_gpu_fnd_batch_of_different_planners_matches_the_oracle():
    """48 planners with their own start / goal / stream: re-root each at its own path node, invalidate, reconnect."""
    case = dict(K.case_by_name("float_star"), iter_num=400)
    rng = np.random.default_rng(12)
    P = 48
    starts, goals = [], []
    while len(starts) < P:
        s, t = rng.uniform(2, 98, 2), rng.uniform(2, 98, 2)
        if O.is_occupied(s, case["obstacles"], 1) or O.is_occupied(t, case["obstacles"], 1) or np.hypot(*(s - t)) < 40:
            continue
        starts.append(tuple(s)); goals.append(tuple(t))
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    b = PlannerBatch("star", starts=starts, goals=goals, obstacles=case["obstacles"], xy_range=case["xy_range"], iter_num=400,
                     step_length=3, radius=6, node_radius=1, bias=0.05, seeds=[5] * P, stream_ids=list(range(P))).run()
    b.check_status()
    trees = []
    prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
    prm.bias = 0.05
    for p in range(P):
        t = O.OracleTree(starts[p], 402)
        assert t.run(prm, case["obstacles"], goals[p], seed=5, stream_id=p)[0] == S.ST_DONE
        trees.append(t)
    cur, extra = [], list(case["obstacles"])
    for p in range(P):
        path = trees[p].best_path()
        assert np.array_equal(path, b.best_path(p))
        cur.append(int(path[-3]) if len(path) >= 6 else -7)          # planners without a (long enough) path get a bad id
        if len(path) >= 6:
            e = trees[p].export(); i = list(e["ids"]).index(int(path[len(path) // 2]))
            if p % 3 == 0:
                extra.append((float(e["x"][i]), float(e["y"][i]), 0.6))   # drop an obstacle on a mid-path node of every third planner
    st = b.select_branch(cur)
    b.set_obstacles(extra)
    sep, err = b.valid_path(cur)
    # reconnect only where a tree was really split off (linking a root to its own descendant would make a cycle)
    linked = b.reconnect([int(sep[p]) if int(sep[p]) != cur[p] else -1 for p in range(P)], 9.0)
    n_ok = n_split = 0
    for p in range(P):
        t = trees[p]
        want = t.select_branch(cur[p])
        assert int(st[p]) == want
        if want != 0:
            assert_tree_equal(b.tree(p), t.export(), f"untouched[{p}]")
            continue
        wsep, werr = t.valid_path(extra, 1, cur[p])
        assert bool(err[p]) == werr
        if werr:
            assert_tree_equal(b.tree(p), t.export(), f"index-error[{p}]")
            continue
        assert int(sep[p]) == wsep
        if wsep == cur[p]:
            assert int(linked[p]) == -1
            assert_tree_equal(b.tree(p), t.export(), f"not-split[{p}]")
            n_ok += 1
            continue
        n_split += 1
        wl = t.reconnect(extra, wsep, 9.0, int(t.best_path()[0]))
        assert int(linked[p]) == wl
        assert_tree_equal(b.tree(p), t.export(), f"reconnected[{p}]")
        assert np.array_equal(b.best_path(p), t.best_path()) and b.planners()[p]["best_cost"] == t.planner().best_cost
        n_ok += 1
    assert n_ok >= P // 2 and n_split >= 4
