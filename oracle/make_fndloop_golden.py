"""Golden vectors for the RRT*-FND replanning LOOP from the UNMODIFIED reference functions — TEST INFRASTRUCTURE, build
container only.

    python -m oracle.make_fndloop_golden          # writes tests/golden/fndloop_*.npz

The reference has no loop (RRT_STAR_FND stops at the init_movement() stub, src/algorithm.py:283-291); what this script
pins is that the COMPOSITION the product runs (rrt_star_fnd_algorithm_b200/fnd_loop.py, restated in oracle/fnd_loop.py)
is the same as calling the reference's own select_branch / valid_path / reconnect / regrow (:294-528) in that order on
the reference's own Graph and Map, tick after tick, on trees that earlier ticks have already pruned, split and mended
(several roots, orphan trees). Harness shims as in make_fnd_golden.py (B12). A planner's record ends with the tick in
which the reference first runs regrow: reconnect_ancestors leaves the reference's children lists inconsistent (SURVEY
App. B14), and its later select_branch / valid_path walk those lists, so from then on the reference contradicts its own
parent pointers and is no longer a specification.
"""
import contextlib
import io
import os

import numpy as np

from . import cases as K
from . import refharness as H
from .fnd_loop import MAX_ROUNDS
from .philox import WordStreamRandom, philox_words

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# (name, base planner case, seed of the obstacle motion, ticks, moving circles, reconnect radius)
LOOP_CASES = [
    dict(name="fn_a", case="float_fn", motion_seed=101, ticks=14, n_moving=30, radius=15.0),     # 14 ticks, 5 reconnects
    dict(name="fn_b", case="float_fn", motion_seed=2, ticks=14, n_moving=30, radius=15.0),
    dict(name="fn_nan", case="float_fn", motion_seed=3, ticks=14, n_moving=30, radius=15.0),     # get_distance_dict cancels to NaN (:947)
    dict(name="fn_late_regrow", case="float_fn", motion_seed=7, ticks=14, n_moving=30, radius=15.0),   # reconnects, then regrow + link in tick 4
    dict(name="fn_regrow_link", case="float_fn", motion_seed=202, ticks=14, n_moving=45, radius=6.0),  # regrow, then a link, in one tick
    dict(name="fn_two_regrows", case="float_fn", motion_seed=2, ticks=14, n_moving=45, radius=10.0),   # two regrows in one tick
    dict(name="star_a", case="float_star", motion_seed=2, ticks=14, n_moving=40, radius=15.0),
    dict(name="star_b", case="float_star", motion_seed=3, ticks=14, n_moving=40, radius=15.0),
    dict(name="star_three_regrows", case="float_star", motion_seed=7, ticks=14, n_moving=40, radius=15.0),
    dict(name="int_a", case="int_star", motion_seed=1, ticks=14, n_moving=25, radius=60.0),
    dict(name="int_lost", case="int_star", motion_seed=4, ticks=14, n_moving=25, radius=60.0),   # regrow gives up: planner lost
    dict(name="main_lost", case="main_fn", motion_seed=404, ticks=8, n_moving=9, radius=120.0),  # the robot's own node is run over
]
REGROW_ITERS = 500          # hard-coded in the reference (src/algorithm.py:485)


def motion(lc, case):
    """Obstacle tables per tick: the case's own circles stay, `n_moving` extra circles drift through the map. Tracks
    keep clear of the goal region (a covered goal-side leaf is the IndexError quirk, covered by make_fnd_golden)."""
    rng = np.random.default_rng(lc["motion_seed"])
    (x0, x1), (y0, y1) = case["xy_range"]
    span = max(x1 - x0, y1 - y0)
    nr, goal, start = case["node_radius"], np.asarray(case["goal"], float), np.asarray(case["start"], float)
    base, vel = [], []
    while len(base) < lc["n_moving"]:
        c = np.array([rng.uniform(x0, x1), rng.uniform(y0, y1)])
        r = rng.uniform(0.006, 0.02) * span
        v = rng.uniform(-0.006, 0.006, 2) * span
        k = np.arange(lc["ticks"] + 1)[:, None]
        track = c[None, :] + k * v[None, :]
        if (np.hypot(track[:, 0] - goal[0], track[:, 1] - goal[1]) < 3 * nr + r + 0.005 * span).any():
            continue
        if np.hypot(*(c - start)) < 2 * nr + r:
            continue
        base.append((float(c[0]), float(c[1]), float(r))); vel.append((float(v[0]), float(v[1])))
    from rrt_star_fnd_algorithm_b200.workloads import moved_obstacles
    return [list(case["obstacles"]) + moved_obstacles(base, vel, k) for k in range(lc["ticks"] + 1)]


def run_case(lc):
    case = K.case_by_name(lc["case"])
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case) + 64 * REGROW_ITERS)
    ref = H.run_reference(case["mode"], start=case["start"], goal=case["goal"], xy_range=case["xy_range"],
                          obstacles=case["obstacles"], words=words, iter_num=case["iter_num"],
                          step_length=case["step_length"], radius=case["radius"], node_radius=case["node_radius"],
                          bias=case["bias"], max_nodes=case["max_nodes"], record=False)
    G, m = ref["graph"], ref["map"]
    alg = H.load_reference().algorithm
    path = [int(i) for i in ref["best_path"]]
    assert len(path) >= 8, (lc["name"], len(path))
    tables = motion(lc, case)
    rng = WordStreamRandom(words)
    rng.cursor = int(ref["words_used"])                   # regrow continues the planner's own stream
    out = {"path0": np.array(path, np.int32), "cursor0": np.int64(rng.cursor), "n_ticks": np.int32(0)}
    for k in range(lc["ticks"] + 1):
        out[f"obst_{k}"] = np.array(tables[k], np.float64)
    orig_gdd = alg.get_distance_dict
    last_node = path[0]
    stop = False
    try:
        for k in range(1, lc["ticks"] + 1):
            if len(path) < 2 or stop:
                break
            rec = dict(current=path[-2], status=0, separate_root=-1, index_error=0, linked=[-1] * MAX_ROUNDS,
                       regrow=[0] * MAX_ROUNDS, lost=0)
            with contextlib.redirect_stdout(io.StringIO()), H.patched_reference(rng, None):
                alg.get_distance_dict = lambda G_, node, idx=None: orig_gdd(G_, node, idx)      # shim B12
                cur = path[-2]
                path = alg.select_branch(G, cur, path, m)
                sb_path = list(path)          # reconnect() reads the root of the robot's tree from path[-1] (:384)
                m.obstacles_c = [((ox, oy), r) for ox, oy, r in tables[k]]                     # the obstacles have moved
                try:
                    sep = alg.valid_path(G, path, m, cur)
                except IndexError:
                    sep, rec["index_error"], rec["lost"] = -1, 1, 1
                rec["separate_root"] = sep
                root_alive = cur in G.vertices
                if not rec["lost"] and not root_alive:
                    rec["lost"] = 1
                todo = sep if (not rec["lost"] and sep != cur) else -1
                broken = False
                for rnd in range(MAX_ROUNDS):
                    if todo < 0:
                        break
                    try:
                        ret = alg.reconnect(G, sb_path, m, lc["radius"], cur, todo, last_node)
                        if ret is not None:
                            rec["linked"][rnd] = int(G.parent[todo])
                        else:
                            old_last = G.last_id
                            alg.regrow(G=G, map=m, step_length=case["step_length"], radius=case["step_length"], id_root=cur,
                                       id_separate_root=todo, last_node=last_node, bias=case["bias"])
                            ok = any(p is not None and p > old_last for i, p in G.parent.items() if i <= old_last)
                            rec["regrow"][rnd] = int(ok)
                            stop = True                                                        # B14: no further ticks
                            if not ok:
                                rec["lost"] = 1
                                todo = -1
                                break
                            ret = alg.find_path(G, last_node, cur)
                    except (ValueError, KeyError, IndexError):
                        broken = True                                                          # the reference tripped over its own children lists (B14)
                        break
                    path = [i for i in ret[0] if i is not None]                                # a broken chain ends in None (:789-792)
                    rec["path_cost"] = float(ret[1])
                    todo = path[-1] if path[-1] != cur else -1
                if broken:
                    break
                if todo >= 0:
                    rec["lost"] = 1
            alg.get_distance_dict = orig_gdd
            for key, v in rec.items():
                out[f"t{k}_{key}"] = np.asarray(v)
            out[f"t{k}_path"] = np.array(path, np.int32)
            out[f"t{k}_cursor"] = np.int64(rng.cursor)
            for key, v in H.tree_arrays(G).items():
                out[f"t{k}_{key}"] = v
            out["n_ticks"] = np.int32(k)
            if rec["lost"]:
                break
    finally:
        alg.get_distance_dict = orig_gdd
    return out


def main():
    for lc in LOOP_CASES:
        out = run_case(lc)
        np.savez_compressed(os.path.join(OUT, f"fndloop_{lc['name']}.npz"), **out)
        n = int(out["n_ticks"])
        print(lc["name"], "ticks", n, "splits", sum(int(out[f"t{k}_separate_root"]) not in (-1, int(out[f"t{k}_current"])) for k in range(1, n + 1)),
              "linked", sum(int((out[f"t{k}_linked"] >= 0).sum()) for k in range(1, n + 1)),
              "regrown", sum(int(out[f"t{k}_regrow"].sum()) for k in range(1, n + 1)),
              "lost", int(out[f"t{n}_lost"]), "nodes", len(out[f"t{n}_ids"]))


if __name__ == "__main__":
    main()
