"""Stress run: the C oracle against the UNMODIFIED Python reference on many fresh random planners — TEST
INFRASTRUCTURE, build container only (the reference is ~100 extensions/s per core).

    python -m oracle.stress_vs_reference [planners=240] [iterations=400] [processes=8] [libm]

Every planner gets its own random map (float circles, so nothing is exactly representable), start / goal, mode
(RRT* / RRT*-FN / RRT) and Philox stream; the run fails on the first tree, best path, cost, word cursor or
attempt count that is not bit-equal. What it measures is the residual risk of SURVEY fact 8: the reference's `**2` /
`**0.5` go through libm `pow`, the restatement (and the CUDA path) use `x*x` / `sqrt`, which differ by an ulp in
~0.1 % of evaluations — a flipped predicate would show up here as a different tree.
"""
import multiprocessing as mp
import sys
import time

import numpy as np


def one(idx, iters, libm_pow=False):
    import oracle as O
    from oracle import cases as K
    from oracle import refharness as H
    from oracle.philox import philox_words
    from rrt_star_fnd_algorithm_b200 import _structs as S
    rng = np.random.default_rng([2024, idx])
    mode = ("star", "fn", "star", "rrt")[idx % 4]
    size = float(rng.choice([60.0, 100.0, 200.0]))
    nr = float(rng.uniform(0.5, 2.0))
    start = (float(rng.uniform(0.05, 0.3) * size), float(rng.uniform(0.05, 0.95) * size))
    goal = (float(rng.uniform(0.7, 0.95) * size), float(rng.uniform(0.05, 0.95) * size))
    obst = K.float_circles(int(rng.integers(1 << 30)), int(rng.integers(20, 120)), 0.02 * size, 0.98 * size, 0.005 * size,
                           0.04 * size, start, goal, nr)
    step = float(rng.uniform(0.02, 0.06) * size)
    case = dict(name=f"stress{idx}", mode=mode, seed=int(rng.integers(1 << 30)), stream_id=idx, start=start, goal=goal,
                xy_range=[[0, size], [0, size]], obstacles=obst, iter_num=iters, step_length=step,
                radius=None if mode == "rrt" else float(rng.uniform(1.5, 3.0) * step), node_radius=nr,
                bias=float(rng.choice([0.0, 0.02, 0.1])), max_nodes=int(rng.integers(40, 200)))
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    ref = H.run_reference(case["mode"], start=start, goal=goal, xy_range=case["xy_range"], obstacles=obst, words=words,
                          iter_num=iters, step_length=step, radius=case["radius"], node_radius=nr, bias=case["bias"],
                          max_nodes=case["max_nodes"], record=True)
    prm = K.params_for(case, n_words=len(words), flags=0)
    O.set_libm_pow(libm_pow)       # the reference's `**` as libm pow() (by construction) instead of x*x / sqrt
    t = O.OracleTree(start, iters + 2)
    st, _ = t.run(prm, obst, goal, words=words)
    got, pl = t.export(), t.planner()
    bad = [k for k in ("ids", "x", "y", "parent", "cost", "nchild")
           if not np.array_equal(np.asarray(got[k]).view(np.uint8 if got[k].dtype.kind == "f" else got[k].dtype),
                                 np.asarray(ref[k]).astype(got[k].dtype).view(np.uint8 if got[k].dtype.kind == "f" else got[k].dtype))]
    if (pl.iter, pl.cursor, pl.attempts) != (ref["iter"], ref["words_used"], ref["rec"]["attempts"]):
        bad.append("counters")
    if not np.array_equal(t.best_path(), ref["best_path"]) or pl.best_cost != ref["best_cost"]:
        bad.append("best_path")
    return idx, mode, int(ref["iter"]), int(ref["rec"]["edge_calls"]) if "edge_calls" in ref["rec"] else 0, bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    libm = len(sys.argv) > 4 and sys.argv[4] == "libm"
    t0 = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.starmap(one, [(i, iters, libm) for i in range(n)], chunksize=1)
    bad = [r for r in res if r[4]]
    ext = sum(r[2] for r in res)
    print(f"[oracle `**` = {'libm pow emulation' if libm else 'x*x / sqrt'}] {n} planners x {iters} iterations: {ext} extensions, {sum(r[3] for r in res)} through_obstacle calls, "
          f"{len(bad)} planners differ, {time.time() - t0:.0f} s")
    for r in bad[:10]:
        print("  DIFFERS:", r)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
