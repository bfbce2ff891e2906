"""How often would libm pow() at the reference's `**` sites change a predicate? — TEST INFRASTRUCTURE.

    make -C oracle audit && python -m oracle.pow_audit [workload=c3] [planners=64] [iterations=2000] [threads]

The reference writes `x ** 2` and `delta ** 0.5` (src/algorithm.py:602-603, :617, :619-620; src/map.py:44); on Python
floats and numpy scalars that is libm pow(), which is not correctly rounded: it differs from x*x in ~0.08 % and from
sqrt in ~0.06 % of inputs, by one ulp. The oracle and the CUDA path use x*x / sqrt. The audit build of the oracle
evaluates every circle test and every occupancy test both ways and counts how often an intermediate value and how often
the resulting BOOLEAN differ while it grows real trees on the bench workloads.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def run(name="c3", P=64, iters=2000, threads=None):
    """Grow P planners of workload `name` with the audit build; returns (extensions, bad planners, the six counters)."""
    import subprocess
    subprocess.check_call(["make", "-C", HERE, "-s", "audit"])
    sys.path.insert(0, os.path.join(HERE, ".."))
    import bench
    from rrt_star_fnd_algorithm_b200 import _structs as S, workloads as W
    import oracle as O
    threads = threads or len(os.sched_getaffinity(0))
    L = C.CDLL(os.path.join(HERE, "_build", "librrt_oracle_audit.so"))
    vp = C.c_void_p
    L.orc_run_batch.argtypes = [C.POINTER(S.Params), C.c_int, vp, vp, vp, vp, C.c_int]
    L.orc_pow_audit.argtypes = [vp, C.c_int]
    w = dict(W.WORKLOADS[name](), iter_num=iters)
    s, g, seeds, sids = W.planner_problems(w, 0, P)
    prm = bench.oracle_params(w, S.FLAG_EVAL_CHOOSE_PARENT)
    sg = np.ascontiguousarray(np.concatenate([s, g], 1))
    sd = np.ascontiguousarray(np.stack([seeds, sids], 1).astype(np.uint64))
    ob = O.obst_table3(w["obstacles"])
    out = np.zeros(P, dtype=S.PLANNER_DTYPE)
    p = lambda a: a.ctypes.data_as(vp)
    cnt = np.zeros(6, np.int64)
    L.orc_pow_audit(p(cnt), 1)
    bad = L.orc_run_batch(C.byref(prm), P, p(sg), p(sd), p(ob), p(out), threads)
    L.orc_pow_audit(p(cnt), 1)
    return int(out["iter"].sum()), int(bad), cnt


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    threads = int(sys.argv[4]) if len(sys.argv) > 4 else len(os.sched_getaffinity(0))
    t0 = time.time()
    ext, bad, cnt = run(name, P, iters, threads)
    print(f"{name}: {P} planners x {iters} iterations on {threads} threads, {ext} extensions, {time.time() - t0:.0f} s, "
          f"bad planners {bad}")
    print(f"  circle tests   {cnt[0]:>14,d}   an intermediate differs {cnt[1]:>12,d} ({100 * cnt[1] / max(cnt[0], 1):.3f} %)   boolean differs {cnt[2]}")
    print(f"  occupancy tests{cnt[3]:>14,d}   the distance differs    {cnt[4]:>12,d} ({100 * cnt[4] / max(cnt[3], 1):.3f} %)   boolean differs {cnt[5]}")


if __name__ == "__main__":
    main()
