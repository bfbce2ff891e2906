"""Stress run of the replanning loop: the oracle's composition against the UNMODIFIED reference's own select_branch /
valid_path / reconnect / regrow composed the same way, on many random obstacle motions — TEST INFRASTRUCTURE, build
container only.

    python -m oracle.stress_fndloop_vs_reference [scenarios=200] [processes=8]

Each scenario: one of the planner cases, a random number / size / velocity of drifting circles, a random reconnect
radius; the reference is driven by oracle/make_fndloop_golden.run_case and the oracle loop must reproduce every tick:
tree, path, path cost, word cursor, which separate trees were linked or regrown, and whether the planner was lost.
"""
import multiprocessing as mp
import sys
import time

import numpy as np


def one(idx):
    import oracle as O
    from oracle import cases as K
    from oracle import make_fndloop_golden as M
    from oracle.fnd_loop import LOST, OracleReplanner
    from rrt_star_fnd_algorithm_b200 import _structs as S
    rng = np.random.default_rng([77, idx])
    case_name = ("float_fn", "float_star", "int_star", "dense_fn", "main_star")[idx % 5]
    case = K.case_by_name(case_name)
    span = case["xy_range"][0][1] - case["xy_range"][0][0]
    lc = dict(name=f"s{idx}", case=case_name, motion_seed=int(rng.integers(1 << 30)), ticks=int(rng.integers(6, 16)),
              n_moving=int(rng.integers(8, 50)), radius=float(rng.uniform(0.04, 0.3) * span))
    try:
        g = M.run_case(lc)
    except AssertionError:
        return idx, case_name, 0, "no path"
    n = int(g["n_ticks"])
    prm = K.params_for(case, stream_mode=S.STREAM_PHILOX, flags=0)
    t = O.OracleTree(case["start"], case["iter_num"] + 2 + (lc["ticks"] + 1) * M.REGROW_ITERS)
    t.run(prm, case["obstacles"], case["goal"], seed=case["seed"], stream_id=case["stream_id"])
    rep = OracleReplanner(t, prm, case["goal"], case["seed"], case["stream_id"], case["node_radius"], lc["radius"], M.REGROW_ITERS)
    for k in range(1, n + 1):
        got = rep.tick(g[f"obst_{k}"])
        bad = []
        if got["current"] != int(g[f"t{k}_current"]): bad.append("current")
        if (got["phase"] == LOST) != bool(int(g[f"t{k}_lost"])): bad.append("lost")
        if list(got["linked"]) != [int(v) for v in g[f"t{k}_linked"]]: bad.append("linked")
        if [int(v == 1) for v in got["regrow"]] != [int(v) for v in g[f"t{k}_regrow"]]: bad.append("regrow")
        e = t.export()
        for key in ("ids", "x", "y", "parent", "cost"):
            a, b = np.asarray(e[key]), np.asarray(g[f"t{k}_{key}"])
            if a.shape != b.shape or not np.array_equal(a.view(np.uint8), b.astype(a.dtype).view(np.uint8)): bad.append(key)
        if t.planner().cursor != int(g[f"t{k}_cursor"]): bad.append("cursor")
        if not int(g[f"t{k}_lost"]) and [int(i) for i in t.best_path()] != [int(i) for i in g[f"t{k}_path"]]: bad.append("path")
        if bad:
            return idx, case_name, k, ",".join(bad)
    return idx, case_name, n, ""


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    t0 = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(one, range(n), chunksize=1)
    bad = [r for r in res if r[3] not in ("", "no path")]
    print(f"{n} scenarios, {sum(r[2] for r in res)} ticks compared, {sum(r[3] == 'no path' for r in res)} without an initial path, "
          f"{len(bad)} differ, {time.time() - t0:.0f} s")
    for r in bad[:10]:
        print("  DIFFERS:", r)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
