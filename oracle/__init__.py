"""CPU oracle for the RRT / RRT* / RRT*-FN / FND hot path — TEST INFRASTRUCTURE ONLY.

`oracle/rrt_oracle.c` restates the reference's algorithm (brute force, one fp64 rounding per
operation). This module builds it with gcc and wraps it with ctypes. It is imported only by
`tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / --impl reference); the product
package never imports it and has no CPU fallback.

Parity status: PINNED. `oracle/make_golden.py` runs the unmodified Python reference in the build
container (`oracle/refharness.py`) and `tests/test_oracle_golden.py` requires bit-equal results.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rrt_star_fnd_algorithm_b200._structs import (  # struct mirrors only; no CUDA library involved
    Params, Planner, Trace, PLANNER_DTYPE, TRACE_DTYPE,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librrt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "rrt_oracle.c"), os.path.join(_HERE, "cull_probe.c"),
            os.path.join(_HERE, "..", "include", "rrtfnd.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    subprocess.check_call(["make", "-C", _HERE, "-s", "powcheck"])      # csrc/libm_pow.h compiled for the host (checker)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.orc_create.restype = vp
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.orc_destroy.argtypes = [vp]
    L.orc_load.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int]
    L.orc_export.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.orc_get_planner.argtypes = [vp, C.POINTER(Planner)]
    L.orc_get_best_path.argtypes = [vp, vp]
    L.orc_get_finish.argtypes = [vp, vp]
    L.orc_through_obstacle.argtypes = [C.c_double] * 4 + [vp, C.c_int]
    L.orc_through_obstacle_batch.argtypes = [vp, C.c_int64, vp, C.c_int, vp]
    L.orc_is_occupied.argtypes = [C.c_double, C.c_double, vp, C.c_int, C.c_double]
    L.orc_get_cost.restype = C.c_double
    L.orc_get_cost.argtypes = [vp, C.c_int]
    L.orc_dist_fma.restype = C.c_double
    L.orc_dist_fma.argtypes = [C.c_double] * 4
    L.orc_dist_plain.restype = C.c_double
    L.orc_dist_plain.argtypes = [C.c_double] * 4
    L.orc_nearest_visible.argtypes = [vp, C.c_double, C.c_double, vp, C.c_int, ip]
    L.orc_near_set.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, vp]
    L.orc_run.argtypes = [vp, C.POINTER(Params), vp, C.c_double, C.c_double, vp, C.c_int64, C.c_uint64,
                          C.c_uint64, vp]
    L.orc_fnd_select_branch.argtypes = [vp, C.c_int]
    L.orc_fnd_valid_path.argtypes = [vp, vp, C.c_int, C.c_double, C.c_int, ip]
    L.orc_fnd_reconnect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_int]
    L.orc_fnd_cut_blocked_edges.argtypes = [vp, vp, C.c_int, C.c_int, ip]
    L.orc_fnd_regrow.argtypes = [vp, C.POINTER(Params), vp, C.c_double, C.c_double, C.c_int, vp, C.c_int64, C.c_uint64, C.c_uint64,
                                 C.c_int, ip]
    L.orc_set_cursor.argtypes = [vp, C.c_uint64]
    L.orc_set_best_path.argtypes = [vp, vp, C.c_int]
    L.orc_get_root.argtypes = [vp]
    L.orc_cull_probe.restype = C.c_int64
    L.orc_cull_probe.argtypes = [C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int64)]
    L.orc_run_batch.argtypes = [C.POINTER(Params), C.c_int, vp, vp, vp, vp, C.c_int]
    L.orc_set_libm_pow.argtypes = [C.c_int]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def obst_table3(obstacles) -> np.ndarray:
    """[(ox, oy, r), ...] (or rows that already carry K) -> contiguous float64 [M][4]: ox, oy, r, K.

    K = -r ** 2 + ox ** 2 + oy ** 2 is evaluated here in Python on the caller's number types with the reference's own
    expression (src/algorithm.py:619): `**` on Python numbers is libm pow (exact integers for int obstacles), which is
    not always fl(r * r)."""
    rows = []
    for ob in obstacles:
        ob = tuple(ob)
        if len(ob) == 4:
            rows.append(tuple(float(v) for v in ob))
            continue
        x0, y0, r = ob
        if isinstance(x0, np.generic): x0 = x0.item()
        if isinstance(y0, np.generic): y0 = y0.item()
        if isinstance(r, np.generic): r = r.item()
        rows.append((float(x0), float(y0), float(r), float(-r ** 2 + x0 ** 2 + y0 ** 2)))
    return np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1, 4))


class OracleTree:
    """One planner's tree + bookkeeping inside the C oracle."""

    def __init__(self, start, id_capacity, finish_cap=4096, path_cap=4096):
        self.L = lib()
        self.id_capacity = int(id_capacity)
        self.h = self.L.orc_create(int(id_capacity), int(finish_cap), int(path_cap), float(start[0]), float(start[1]))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, prm: Params, obstacles, goal, words=None, seed=0, stream_id=0, trace=False):
        ob = obst_table3(obstacles)
        w = None if words is None else np.ascontiguousarray(words, dtype=np.uint32)
        tr = np.zeros(max(int(prm.trace_cap), 1), dtype=TRACE_DTYPE) if trace else None
        st = self.L.orc_run(self.h, C.byref(prm), _ptr(ob), float(goal[0]), float(goal[1]), _ptr(w),
                            0 if w is None else len(w), int(seed), int(stream_id), _ptr(tr))
        return st, tr

    def planner(self) -> Planner:
        p = Planner()
        self.L.orc_get_planner(self.h, C.byref(p))
        return p

    def export(self):
        n = self.id_capacity
        ids = np.zeros(n, np.int32); x = np.zeros(n); y = np.zeros(n)
        parent = np.zeros(n, np.int32); cost = np.zeros(n); nchild = np.zeros(n, np.int32)
        k = self.L.orc_export(self.h, _ptr(ids), _ptr(x), _ptr(y), _ptr(parent), _ptr(cost), _ptr(nchild))
        return dict(ids=ids[:k].copy(), x=x[:k].copy(), y=y[:k].copy(), parent=parent[:k].copy(),
                    cost=cost[:k].copy(), nchild=nchild[:k].copy())

    def load(self, ids, x, y, parent, cost, root_id, last_id):
        a = [np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(x, np.float64),
             np.ascontiguousarray(y, np.float64), np.ascontiguousarray(parent, np.int32),
             np.ascontiguousarray(cost, np.float64)]
        rc = self.L.orc_load(self.h, len(a[0]), *[_ptr(v) for v in a], int(root_id), int(last_id))
        if rc:
            raise ValueError("orc_load: id out of capacity")

    def best_path(self):
        out = np.zeros(65536, np.int32)
        n = self.L.orc_get_best_path(self.h, _ptr(out))
        return out[:n].copy()

    def finish(self):
        out = np.zeros(65536, np.int32)
        n = self.L.orc_get_finish(self.h, _ptr(out))
        return out[:n].copy()

    # ---- FND tree surgery (src/algorithm.py:294-404) on the stored best path
    def set_best_path(self, ids):
        a = np.ascontiguousarray(ids, np.int32)
        if self.L.orc_set_best_path(self.h, _ptr(a), len(a)):
            raise ValueError("path longer than path_cap")

    def root(self):
        return self.L.orc_get_root(self.h)

    def select_branch(self, current_id):
        return self.L.orc_fnd_select_branch(self.h, int(current_id))

    def valid_path(self, obstacles, node_radius, previous_root):
        ob = obst_table3(obstacles)
        err = C.c_int(0)
        sep = self.L.orc_fnd_valid_path(self.h, _ptr(ob), len(ob), float(node_radius), int(previous_root), C.byref(err))
        return sep, bool(err.value)

    def cut_blocked_edges(self, obstacles, separate_root):
        """Edge invalidation (extension, see rrt_oracle.c): returns (separate root after the cuts, edges cut)."""
        ob = obst_table3(obstacles)
        n = C.c_int(0)
        sep = self.L.orc_fnd_cut_blocked_edges(self.h, _ptr(ob), len(ob), int(separate_root), C.byref(n))
        return sep, n.value

    def reconnect(self, obstacles, separate_root, radius, last_node):
        ob = obst_table3(obstacles)
        return self.L.orc_fnd_reconnect(self.h, _ptr(ob), len(ob), int(separate_root), float(radius), int(last_node))

    def regrow(self, prm: Params, obstacles, goal, separate_root, words=None, seed=0, stream_id=0, max_iter=500, cursor=None):
        """regrow (src/algorithm.py:482-528): returns (reconnected or negative status, extensions made)."""
        ob = obst_table3(obstacles)
        prm.n_obst = len(ob)
        w = None if words is None else np.ascontiguousarray(words, dtype=np.uint32)
        if cursor is not None:
            self.L.orc_set_cursor(self.h, int(cursor))
        it = C.c_int(0)
        r = self.L.orc_fnd_regrow(self.h, C.byref(prm), _ptr(ob), float(goal[0]), float(goal[1]), int(separate_root), _ptr(w),
                                  0 if w is None else len(w), int(seed), int(stream_id), int(max_iter), C.byref(it))
        return r, it.value

    def get_cost(self, id_):
        return self.L.orc_get_cost(self.h, int(id_))

    def nearest_visible(self, q, obstacles):
        ob = obst_table3(obstacles)
        rounds = C.c_int(0)
        i = self.L.orc_nearest_visible(self.h, float(q[0]), float(q[1]), _ptr(ob), len(ob), C.byref(rounds))
        return i, rounds.value

    def near_set(self, q, radius, skip_id=-1):
        out = np.zeros(self.id_capacity, np.int32)
        n = self.L.orc_near_set(self.h, float(q[0]), float(q[1]), float(radius), int(skip_id), _ptr(out))
        return out[:n].copy()


def through_obstacle(p1, p2, obstacles) -> bool:
    ob = obst_table3(obstacles)
    return bool(lib().orc_through_obstacle(float(p1[0]), float(p1[1]), float(p2[0]), float(p2[1]), _ptr(ob), len(ob)))


def through_obstacle_batch(p1, p2, obstacles) -> np.ndarray:
    seg = np.ascontiguousarray(np.concatenate([np.asarray(p1, np.float64).reshape(-1, 2), np.asarray(p2, np.float64).reshape(-1, 2)], 1))
    ob = obst_table3(obstacles)
    out = np.zeros(len(seg), np.uint8)
    lib().orc_through_obstacle_batch(_ptr(seg), len(seg), _ptr(ob), len(ob), _ptr(out))
    return out.astype(bool)


def is_occupied(p, obstacles, node_radius) -> bool:
    ob = obst_table3(obstacles)
    return bool(lib().orc_is_occupied(float(p[0]), float(p[1]), _ptr(ob), len(ob), float(node_radius)))


def run_batch(prm: Params, starts_goals, seeds, obstacles, n_threads=1):
    """CPU baseline driver: P planners (Philox streams) on n_threads host threads. Returns planner records."""
    sg = np.ascontiguousarray(starts_goals, np.float64).reshape(-1, 4)
    sd = np.ascontiguousarray(seeds, np.uint64).reshape(-1, 2)
    ob = obst_table3(obstacles)
    out = np.zeros(len(sg), dtype=PLANNER_DTYPE)
    bad = lib().orc_run_batch(C.byref(prm), len(sg), _ptr(sg), _ptr(sd), _ptr(ob), _ptr(out), int(n_threads))
    return out, bad


def set_libm_pow(on: bool):
    """Evaluate the reference's `** 2` / `** 0.5` as the libm pow() they call (csrc/libm_pow.h) instead of x*x / sqrt —
    process-wide switch of the oracle; the CUDA planners compute the x*x / sqrt form."""
    lib().orc_set_libm_pow(int(bool(on)))


def cull_probe(seed, n_trials, coord_max, slope_max, margin, scale=1.0):
    """Adversarial check of the grid-culling rule (oracle/cull_probe.c): (hits, valid pairs)."""
    nv = C.c_int64(0)
    h = lib().orc_cull_probe(int(seed), int(n_trials), float(coord_max), float(slope_max), float(margin), float(scale),
                             C.byref(nv))
    return int(h), int(nv.value)
