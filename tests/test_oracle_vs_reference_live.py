"""Build container only: the C oracle against the unmodified Python reference run live (fresh random cases)."""
import numpy as np
import pytest

import oracle as O
from oracle import cases as K
from oracle import refharness as H
from oracle.philox import philox_words
from rrt_star_fnd_algorithm_b200 import _structs as S

from util import assert_tree_equal

pytestmark = pytest.mark.reference


def _live(case, ascending_ball=True):
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    ref = H.run_reference(case["mode"], start=case["start"], goal=case["goal"], xy_range=case["xy_range"],
                          obstacles=case["obstacles"], words=words, iter_num=case["iter_num"],
                          step_length=case["step_length"], radius=case["radius"], node_radius=case["node_radius"],
                          bias=case["bias"], max_nodes=case["max_nodes"], ascending_ball=ascending_ball)
    prm = K.params_for(case, n_words=len(words))
    tree = O.OracleTree(case["start"], case["iter_num"] + 2)
    st, tr = tree.run(prm, case["obstacles"], case["goal"], words=words, trace=True)
    return ref, tree, st, tr


@pytest.mark.parametrize("seed", [101, 102, 103])
@pytest.mark.parametrize("mode", ["star", "fn", "rrt"])
def test_fresh_random_cases(seed, mode):
    rng = np.random.default_rng(seed)
    start, goal = (float(rng.uniform(5, 20)), float(rng.uniform(5, 95))), (float(rng.uniform(80, 95)), float(rng.uniform(5, 95)))
    obst = K.float_circles(seed, 50, 2, 98, 0.5, 3.5, start, goal, 1.5)
    case = dict(name="live", mode=mode, seed=seed, stream_id=3, start=start, goal=goal, xy_range=[[0, 100], [0, 100]],
                obstacles=obst, iter_num=300, step_length=4, radius=8 if mode != "rrt" else None, node_radius=1.5,
                bias=0.03, max_nodes=80)
    ref, tree, st, tr = _live(case)
    assert st == S.ST_DONE
    assert_tree_equal(tree.export(), ref, f"{mode}/{seed}")
    pl = tree.planner()
    assert (pl.iter, pl.cursor, pl.attempts) == (ref["iter"], ref["words_used"], ref["rec"]["attempts"])
    assert np.array_equal(tree.best_path(), ref["best_path"]) and pl.best_cost == ref["best_cost"]
    n = pl.iter
    assert np.array_equal(tr["removed_id"][:n], ref["rec"]["removed"][:n])
    if mode != "rrt":
        assert np.array_equal(tr["cp_parent"][:n], ref["rec"]["cp_parent"][:n])


def test_tree_does_not_depend_on_ball_order():
    """With scipy's native (tree-internal) ball order the reference grows the same tree (rewire is order free)."""
    case = K.case_by_name("float_star")
    ref, tree, st, _ = _live(case, ascending_ball=False)
    assert_tree_equal(tree.export(), ref, "native ball order")
