"""The reference's per-iteration helpers on Graph objects (new_node, nearest_node[_kdtree], steer, choose_parent[_kdtree],
rewire[_kdtree], forced_removal, intersection_circle with calc_delta / delta_solutions / is_between, Line, remove_children,
reconnect_ancestors, delete_ancestors, get_near_nodes, check_for_tree_associativity).

tests/golden/helpers_*.npz hold what the UNMODIFIED reference returns for one fixed call sequence
(oracle/make_helper_golden.py `drive`); the same sequence runs here on the product's modules and every return value, side
effect on the Graph and the following random.random() must be equal bit for bit. GPU test: the device operators do the
collision tests, the visibility search and the get_cost walks. CPU test: the same host logic with those three operators
answered by the CPU oracle (checker only), so that the bookkeeping is covered where there is no GPU."""
import contextlib
import io

import numpy as np
import pytest

import oracle as O
from oracle import cases as K
from oracle.make_helper_golden import HELPER_CASES, drive
from oracle.refharness import AscendingKDTree

from util import load_npz


def run_and_compare(name):
    from rrt_star_fnd_algorithm_b200 import algorithm, graph, line, map as map_mod
    g = load_npz(f"helpers_{name}.npz")
    with contextlib.redirect_stdout(io.StringIO()):
        got = drive(algorithm, graph, map_mod, line, AscendingKDTree, K.case_by_name(name))
    assert sorted(got) == sorted(g)
    for k in sorted(g):
        a, b = np.asarray(got[k]), np.asarray(g[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if a.dtype.kind == "f":
            assert np.array_equal(a.astype(np.float64).view(np.uint64), b.astype(np.float64).view(np.uint64)), k   # bit-exact fp64
        else:
            assert np.array_equal(a, b), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", HELPER_CASES)
def test_helpers_reproduce_the_reference_on_the_gpu(name):
    run_and_compare(name)


@pytest.mark.parametrize("name", HELPER_CASES)
def test_helper_host_logic_with_oracle_backed_operators(name, monkeypatch):
    """No GPU here: the three device operators the helpers call are answered by the CPU oracle."""
    from rrt_star_fnd_algorithm_b200 import algorithm as A, map as map_mod

    def tab(obstacles):
        return [(float(o[0][0]), float(o[0][1]), float(o[1])) if len(o) == 2 else tuple(o) for o in obstacles]

    def tob(p1, p2, obstacles):
        a, b = np.asarray(p1, np.float64).reshape(-1, 2), np.asarray(p2, np.float64).reshape(-1, 2)
        if len(a) == 0 or len(obstacles) == 0:
            return np.zeros(len(a), bool)
        return O.through_obstacle_batch(a, b, tab(obstacles))

    def oracle_tree(G):
        ids, xy, parent, cost, root = G._export()
        t = O.OracleTree(G.vertices[int(ids[0])], int(ids.max()) + 2)
        t.load(ids, xy[:, 0], xy[:, 1], parent, cost, root, int(ids.max()))
        return t

    def chain(G, node_ids, obstacles=()):
        t = oracle_tree(G)
        return np.array([t.get_cost(int(i)) for i in node_ids], np.float64)

    def nearest_kd(G, vertex, obstacles, separate_tree_nodes=(), kdtree=None):
        try:
            return np.array(vertex), G.id_vertex[vertex]
        except KeyError:
            pass
        i, _ = oracle_tree(G).nearest_visible(vertex, tab(obstacles))
        return (np.array(vertex), None) if i < 0 else (np.array(G.vertices[i], dtype=np.float64), i)

    monkeypatch.setattr(A, "through_obstacle_batch", tob)
    monkeypatch.setattr(A, "_chain_costs", chain)
    monkeypatch.setattr(A, "nearest_node_kdtree", nearest_kd)
    monkeypatch.setattr(map_mod.Map, "is_occupied_c", lambda self, p: O.is_occupied(p, tab(self.obstacles_c), self.node_radius))
    run_and_compare(name)


def test_line_object_matches_the_reference_attributes():
    from rrt_star_fnd_algorithm_b200.line import Line
    l = Line((1.0, 2.0), (4.0, 8.0))
    assert l.dir == 2.0 and l.const_term == 0.0 and l.calculate_y(3.0) == 6.0 and l.calculate_x(6.0) == 3.0
    assert l.norm == np.linalg.norm(np.array((4.0, 8.0)) - np.array((1.0, 2.0)))
    v = Line((3, 1), (3, 9))
    assert v.dir == float("inf") and v.x_pos == 3 and v.const_term is None
    k = Line((2.0, 5.0), (0.0, 1.0))                 # the intercept uses p1 (src/line.py:16)
    assert k.const_term == 5.0 - 2.0 * 2.0
