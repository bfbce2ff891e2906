"""Golden vectors for the FND tree surgery (select_branch, valid_path, reconnect; src/algorithm.py:294-434) from the
UNMODIFIED reference — TEST INFRASTRUCTURE, build container only.

    python -m oracle.make_fnd_golden          # writes tests/golden/fnd_*.npz

The sequence is the one of the reference's only usage example, test_select_branch (src/algorithm.py:437-479): grow an
RRT*-FN tree, re-root at a node of the best path, drop obstacles onto later path nodes, validate the path, reconnect.
Two harness shims, reference files untouched: Graph gets its xy_range argument, and
get_near_nodes' call of get_distance_dict(G, node) gets a default for the unused third parameter.
"""
import contextlib
import io
import os

import numpy as np

from . import cases as K
from . import refharness as H
from .philox import philox_words

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# (name, base case, index of the re-root node counted from the root, path indices (goal side first, in the NEW path)
#  that receive an obstacle, obstacle radius, reconnect radius)
FND_CASES = [
    dict(name="main_a", case="main_fn", reroot=2, block=[3], block_r=1.0, radius=125.0),
    dict(name="float_c", case="float_fn", reroot=1, block=[20, 30], block_r=0.5, radius=8.0),
    dict(name="float_a", case="float_fn", reroot=4, block=[6, 7], block_r=0.8, radius=15.0),
    dict(name="float_b", case="float_fn", reroot=3, block=[5], block_r=0.8, radius=2.0),      # too small: no reconnect
    dict(name="star_a", case="float_star", reroot=5, block=[8], block_r=0.7, radius=30.0),
    dict(name="star_b", case="float_star", reroot=2, block=[12, 13, 25], block_r=0.4, radius=12.0),
    dict(name="int_a", case="int_star", reroot=1, block=[2], block_r=2.0, radius=150.0),
    dict(name="leaf_blocked", case="float_fn", reroot=2, block=[0], block_r=0.8, radius=15.0),  # IndexError at :358
    dict(name="none_blocked", case="main_fn", reroot=3, block=[], block_r=1.0, radius=80.0),
    dict(name="cascade", case="main_fn", reroot=1, block=[4], block_r=7.0, radius=200.0),     # inflated circle covers several path nodes
]


def run_case(fc):
    case = K.case_by_name(fc["case"])
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    ref = H.run_reference(case["mode"], start=case["start"], goal=case["goal"], xy_range=case["xy_range"],
                          obstacles=case["obstacles"], words=words, iter_num=case["iter_num"],
                          step_length=case["step_length"], radius=case["radius"], node_radius=case["node_radius"],
                          bias=case["bias"], max_nodes=case["max_nodes"], record=False)
    G, m = ref["graph"], ref["map"]
    alg = H.load_reference().algorithm
    path = list(ref["best_path"])
    assert len(path) > max(fc["block"] + [0]) + fc["reroot"] + 2, (fc["name"], len(path))
    out = {"path0": np.array(path, np.int32)}
    last_node = path[0]
    current = list(reversed(path))[fc["reroot"]]
    orig_gdd = alg.get_distance_dict
    alg.get_distance_dict = lambda G_, node, idx=None: orig_gdd(G_, node, idx)      # shim B12
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            new_path = alg.select_branch(G, current, path, m)
            out["current"] = np.int32(current)
            out["path1"] = np.array(new_path, np.int32)
            for k, v in H.tree_arrays(G).items():
                out[f"sb_{k}"] = v
            out["sb_root"] = np.int32(G.id_vertex[G.start])
            extra = [[G.vertices[new_path[i]], fc["block_r"]] for i in fc["block"]]
            m.add_obstacles(extra)
            out["obstacles2"] = np.array([[float(o[0][0]), float(o[0][1]), float(o[1])] for o in m.obstacles_c])
            err = 0
            try:
                sep = alg.valid_path(G, new_path, m, current)
            except IndexError:
                sep, err = -1, 1
            out["vp_sep"], out["vp_index_error"] = np.int32(sep), np.int32(err)
            # the Graph may now hold several roots; children lists of removed parents are gone, which tree_arrays tolerates
            for k, v in H.tree_arrays(G).items():
                out[f"vp_{k}"] = v
            if not err:
                root = G.id_vertex[G.start]
                ret = alg.reconnect(G, new_path, m, fc["radius"], root, sep, last_node)
                out["rc_linked"] = np.int32(-1 if ret is None else G.parent[sep])
                if ret is not None:
                    out["rcret_path"], out["rcret_cost"] = np.array(ret[0], np.int32), np.float64(ret[1])
                for k, v in H.tree_arrays(G).items():
                    out[f"rc_{k}"] = v
                if ret is None:
                    # regrow (src/algorithm.py:482-528) on its own word stream, as test_select_branch calls it (:477-479)
                    from .philox import WordStreamRandom
                    words = philox_words(777, FND_CASES.index(fc), 80000)
                    rng = WordStreamRandom(words)
                    with H.patched_reference(rng, None):
                        alg.get_distance_dict = lambda G_, node, idx=None: orig_gdd(G_, node, idx)
                        alg.regrow(G=G, map=m, step_length=case["step_length"], radius=case["step_length"], id_root=root,
                                   id_separate_root=sep, last_node=last_node, bias=0.02)
                    out["rg_words_used"] = np.int64(rng.cursor)
                    # regrow returns nothing; it has reconnected iff an old node now hangs under a node created by regrow
                    old_last = int(out["rc_ids"].max())
                    out["rg_reconnected"] = np.int32(any(p is not None and p > old_last for i, p in G.parent.items() if i <= old_last))
                    for k, v in H.tree_arrays(G).items():
                        out[f"rg_{k}"] = v
    finally:
        alg.get_distance_dict = orig_gdd
    return out


def main():
    for fc in FND_CASES:
        out = run_case(fc)
        np.savez_compressed(os.path.join(OUT, f"fnd_{fc['name']}.npz"), **out)
        print(fc["name"], "path", len(out["path0"]), "->", len(out["path1"]), "nodes", len(out["sb_ids"]), "->",
              len(out["vp_ids"]), "sep", int(out["vp_sep"]), "err", int(out["vp_index_error"]),
              "linked", int(out.get("rc_linked", -9)), "regrow", int(out.get("rg_reconnected", -9)), int(out.get("rg_words_used", 0)),
              len(out.get("rg_ids", [])))


if __name__ == "__main__":
    main()
