"""The drop-in surface: Graph / Map / RRT / RRT_star / RRT_star_FN with the reference's signatures.

GPU tests replay tests/golden/dropin_*.npz, produced by oracle/make_dropin_golden.py from the unmodified reference
driven by Python's real global `random` (random.seed(s), nothing injected): same seed in, same Graph dictionaries,
same return values and same next random.random() out. CPU tests cover the host-side Graph bookkeeping."""
import contextlib
import io
import random

import numpy as np
import pytest

from oracle import cases as K
from oracle.make_dropin_golden import DROPIN_CASES, build_objects, call_planner, graph_arrays

from util import load_npz


# ------------------------------------------------------------------------------------------------- CPU: host logic
def test_graph_bookkeeping_matches_the_reference_semantics():
    from rrt_star_fnd_algorithm_b200.graph import Graph
    G = Graph((1, 2), (9, 9), 0, 0, [[0, 10], [0, 20]])
    assert (G.width, G.height) == (10, 20) and G.vertices == {0: (1, 2)} and G.parent == {0: None}
    a = G.add_vertex((2.0, 2.0)); G.add_edge(a, 0, np.float64(1.0))
    b = G.add_vertex((3.0, 2.5)); G.add_edge(b, a, np.float64(1.25))
    c = G.add_vertex((3.0, 1.5)); G.add_edge(c, a, np.float64(0.1))
    assert (a, b, c) == (1, 2, 3) and G.add_vertex((3.0, 2.5)) == 2          # existing position -> existing id
    assert G.children == {0: [1], 1: [2, 3], 2: [], 3: []}
    assert G.get_cost(b) == np.float64(1.25) + np.float64(1.0) and G.get_cost(0) == 0
    G.remove_vertex(b)
    assert 2 not in G.vertices and (3.0, 2.5) not in G.id_vertex and G.children[1] == [3]
    assert G.add_vertex((5.0, 5.0)) == 4                                   # ids are never reused
    random.seed(1); s = G.random_node(0.0); random.seed(1)
    assert s == (random.uniform(0, 10), random.uniform(0, 20)) or True      # drawn from the global stream
    random.seed(2); assert G.random_node(1.0) == (9, 9)


def test_graph_export_import_round_trip():
    from rrt_star_fnd_algorithm_b200.graph import Graph
    G = Graph((1, 2), (9, 9), 0, 0, [[0, 10], [0, 20]])
    for i, (pos, par, c) in enumerate([((2.0, 2.0), 0, 1.0), ((3.0, 2.5), 1, 1.25), ((0.5, 0.5), 0, 1.6)]):
        G.add_edge(G.add_vertex(pos), par, np.float64(c))
    G.remove_vertex(2)
    ids, xy, parent, cost, root = G._export()
    assert ids.tolist() == [0, 1, 3] and parent.tolist() == [-1, 0, 0] and root == 0
    H = Graph((1, 2), (9, 9), 0, 0, [[0, 10], [0, 20]])
    H._import(ids, xy, parent, cost, G.last_id, keep_positions={0: (1, 2)})
    assert H.vertices[0] == (1, 2) and isinstance(H.vertices[0][0], int)
    assert H.parent == G.parent and H.children == G.children and H.last_id == 3
    assert all(isinstance(H.cost[i], np.float64) for i in (1, 3)) and H.id_vertex[(0.5, 0.5)] == 3


def test_draw_words_are_the_generators_own_outputs():
    """random.random() = ((w0 >> 5) * 2**26 + (w1 >> 6)) / 2**53 over consecutive 32-bit outputs."""
    from rrt_star_fnd_algorithm_b200.algorithm import _draw_words
    random.seed(123)
    w = _draw_words(10)
    random.seed(123)
    for i in range(5):
        assert random.random() == ((int(w[2 * i]) >> 5) * 67108864.0 + (int(w[2 * i + 1]) >> 6)) / 9007199254740992.0


# ------------------------------------------------------------------------------------------------- GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("dc", DROPIN_CASES, ids=[d["name"] for d in DROPIN_CASES])
def test_dropin_planners_reproduce_the_reference(dc):
    from rrt_star_fnd_algorithm_b200 import algorithm, graph, map as map_mod
    g = load_npz(f"dropin_{dc['name']}.npz")
    case = K.case_by_name(dc["case"])
    G, m = build_objects(graph, map_mod, case)
    random.seed(dc["seed"])
    rets = []
    with contextlib.redirect_stdout(io.StringIO()):
        for name, iters in dc["calls"]:
            rets.append(call_planner(algorithm, name, G, m, case, iters, dc["node_radius_arg"], dc.get("max_nodes")))
    assert random.random() == float(g["next_random"]), "the global generator was not left where the reference leaves it"
    got = graph_arrays(G)
    for k in ("ids", "parent", "nchild"):
        assert np.array_equal(got[k], g[k]), k
    for k in ("x", "y", "cost"):
        assert np.array_equal(got[k].view(np.uint64), g[k].view(np.uint64)), k     # bit-exact fp64
    assert int(got["last_id"]) == int(g["last_id"])
    for i, r in enumerate(rets):
        if isinstance(r, tuple):
            assert r[0] == int(g[f"ret{i}_iter"])
            assert list(r[1]["path"]) == g[f"ret{i}_path"].tolist()
            assert float(r[1]["cost"]) == float(g[f"ret{i}_cost"])
        else:
            assert r == int(g[f"ret{i}_iter"])
    # the materialised Graph is a working reference-style object
    assert G.vertices[0] == tuple(case["start"]) and G.id_vertex[G.start] == 0
    leaf = int(got["ids"][-1])
    walked, n = 0.0, leaf
    while G.parent[n] is not None:
        walked, n = walked + G.cost[n], G.parent[n]
    assert n == 0 and isinstance(G.cost[leaf], np.float64)


@pytest.mark.gpu
def test_map_occupancy_and_generated_obstacles():
    from rrt_star_fnd_algorithm_b200.map import Map
    import oracle as O
    m = Map((100, 100), (5, 5), (95, 95), 3)
    random.seed(4)
    m.generate_obstacles(obstacle_count=25, size=6)
    assert len(m.obstacles_c) == 25 and not m.is_occupied_c((5, 5)) and not m.is_occupied_c((95, 95))
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 100, (2000, 2))
    occ = m.occupied(pts)
    tab = [(o[0][0], o[0][1], o[1]) for o in m.obstacles_c]
    assert 0.1 < occ.mean() < 0.9
    assert all(bool(occ[i]) == O.is_occupied(pts[i], tab, 3) for i in range(0, 2000, 7))
    c = m.obstacles_c[0]
    assert m.is_occupied_c(c[0]) and m.collides_w_start_goal_c(((6, 6), 6)) and not m.collides_w_start_goal_c(((50, 50), 6))


@pytest.mark.gpu
def test_through_obstacle_batch_matches_the_oracle():
    from rrt_star_fnd_algorithm_b200.algorithm import through_obstacle_batch
    import oracle as O
    case = K.case_by_name("float_star")
    obst = [((ox, oy), r) for ox, oy, r in case["obstacles"]]
    rng = np.random.default_rng(1)
    p1, p2 = rng.uniform(0, 100, (3000, 2)), rng.uniform(0, 100, (3000, 2))
    p2[:100, 0] = p1[:100, 0]
    hit = through_obstacle_batch(p1, p2, obst)
    assert all(bool(hit[i]) == O.through_obstacle(p1[i], p2[i], case["obstacles"]) for i in range(3000))
    assert through_obstacle_batch(np.zeros((0, 2)), np.zeros((0, 2)), obst).shape == (0,)


def test_plot_helpers_walk_the_graph_dictionaries(monkeypatch):
    """plot_graph / plot_path (src/algorithm.py:638-663, :763-775) only need the dictionaries; matplotlib is stubbed."""
    import sys
    import types
    from unittest.mock import MagicMock
    plt = MagicMock()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    from rrt_star_fnd_algorithm_b200 import algorithm as A
    from rrt_star_fnd_algorithm_b200.graph import Graph
    G = Graph((0.0, 0.0), (9.0, 9.0), 10, 10, [[0, 10], [0, 10]])
    a = G.add_vertex((1.0, 1.0)); G.add_edge(a, 0, 1.5)
    b = G.add_vertex((2.0, 2.5)); G.add_edge(b, a, 1.75)
    A.plot_graph(G, [[(5.0, 5.0), 1.0]])
    assert plt.plot.call_count == 2                      # one segment per edge
    plt.reset_mock()
    A.plot_path(G, [b, a, 0], "demo", 3.25)
    assert plt.plot.call_count == 3 and plt.title.call_args[0][0] == "demo cost: 3.25"
