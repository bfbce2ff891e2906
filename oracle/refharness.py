"""Run the UNMODIFIED Python reference headless on an explicit random-word stream (TEST INFRASTRUCTURE).

Only usable where `/root/reference` exists (the build container). Used by
`oracle/make_golden.py` to write the fixtures under `tests/golden/` and by the CPU tests that
pin `oracle/rrt_oracle.c` against the reference itself. Nothing here is imported by the product.

What is patched, and why (reference files are never edited):
  * `matplotlib` / `matplotlib.pyplot` are absent in this image and imported at module level
    (src/algorithm.py:4, src/map.py:3) -> inert stub modules in `sys.modules`.
  * `graph.random`, `algorithm.random`, `map.random` (the stdlib module object, used at
    src/graph.py:96-98 and src/algorithm.py:813-816) -> a `WordStreamRandom` instance, so the
    reference draws from the same 32-bit word stream as the oracle and the CUDA path.
  * `algorithm.tqdm` -> silent no-op bar; `algorithm.plot_path/plot_graph` -> no-ops.
  * optional recorders around `new_node`, `choose_parent_kdtree`, `rewire_kdtree`,
    `forced_removal`, `through_obstacle` (per-iteration golden records).
  * optional `algorithm.cKDTree` -> `AscendingKDTree`, a wrapper over the real cKDTree that returns
    ball-query indices in ascending order. It only matters for the *dropped* return value of
    choose_parent_kdtree (src/algorithm.py:845-852 is order dependent); trees are identical either way.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("RRTFND_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "algorithm.py"))


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_rrtfnd_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    mpl._rrtfnd_stub = True
    plt = types.ModuleType("matplotlib.pyplot")

    class _Anything:
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, name):
            return self

    def _getattr(name):
        return _Anything()

    plt.__getattr__ = _getattr
    mpl.__getattr__ = _getattr
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_REF = None


def load_reference():
    """Import the reference's modules (graph, map, line, algorithm) by path; cached."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    _install_matplotlib_stub()
    # the reference's modules are top-level names (`graph`, `map`, `line`); keep them out of the
    # way of anything else by importing with the path pushed only for the duration of the import
    sys.path.insert(0, REFERENCE_SRC)
    try:
        with contextlib.redirect_stdout(io.StringIO()):  # src/graph.py:4-5 prints at import
            import importlib
            mods = {n: importlib.import_module(n) for n in ("graph", "map", "line", "algorithm")}
    finally:
        sys.path.remove(REFERENCE_SRC)
    _REF = types.SimpleNamespace(**mods)
    return _REF


class _NoBar:
    def __init__(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def close(self):
        pass


class AscendingKDTree:
    """Real cKDTree underneath; `query_ball_point` results sorted ascending (see module doc)."""

    def __init__(self, data, *a, **k):
        from scipy.spatial import cKDTree
        self._t = cKDTree(data, *a, **k)
        self.data = self._t.data

    def query(self, *a, **k):
        return self._t.query(*a, **k)

    def query_ball_point(self, *a, **k):
        return sorted(self._t.query_ball_point(*a, **k))


@contextlib.contextmanager
def patched_reference(rng, record=None, ascending_ball=True):
    """Context in which the reference modules draw from `rng` and (optionally) record events."""
    ref = load_reference()
    alg = ref.algorithm
    saved = {
        (ref.graph, "random"): ref.graph.random,
        (ref.map, "random"): ref.map.random,
        (alg, "random"): alg.random,
        (alg, "tqdm"): alg.tqdm,
        (alg, "plot_path"): alg.plot_path,
        (alg, "plot_graph"): alg.plot_graph,
        (alg, "cKDTree"): alg.cKDTree,
    }
    for name in ("new_node", "choose_parent_kdtree", "rewire_kdtree", "forced_removal", "through_obstacle"):
        saved[(alg, name)] = getattr(alg, name)
    try:
        ref.graph.random = rng
        ref.map.random = rng
        alg.random = rng
        alg.tqdm = _NoBar
        alg.plot_path = lambda *a, **k: None
        alg.plot_graph = lambda *a, **k: None
        if ascending_ball:
            alg.cKDTree = AscendingKDTree
        if record is not None:
            _install_recorders(alg, saved, record)
        yield ref
    finally:
        for (mod, name), val in saved.items():
            setattr(mod, name, val)


def _install_recorders(alg, saved, rec):
    rec.setdefault("attempts", 0)
    rec.setdefault("id_new", [])
    rec.setdefault("id_near", [])
    rec.setdefault("n_near", [])
    rec.setdefault("cp_parent", [])
    rec.setdefault("cp_cost", [])
    rec.setdefault("rewired", [])
    rec.setdefault("removed", [])
    rec.setdefault("edges", [])       # (p1x,p1y,p2x,p2y,hit) of every through_obstacle call
    rec.setdefault("edge_calls", 0)
    keep_edges = rec.get("keep_edges", 0)

    o_new_node = saved[(alg, "new_node")]
    o_cp = saved[(alg, "choose_parent_kdtree")]
    o_rw = saved[(alg, "rewire_kdtree")]
    o_fr = saved[(alg, "forced_removal")]
    o_to = saved[(alg, "through_obstacle")]

    def new_node(G, map, obstacles, step_length, bias):
        rec["attempts"] += 1
        out = o_new_node(G, map, obstacles, step_length, bias)
        if out[0] is not None:
            rec["id_new"].append(int(out[1]))
            rec["id_near"].append(int(out[2]))
            rec["removed"].append(-1)
        return out

    def choose_parent_kdtree(G, q_new, id_new, best_edge, radius, obstacles, *a, **k):
        kd = k.get("kdtree")
        ball = kd.query_ball_point(q_new, r=radius)
        rec["n_near"].append(len(ball) - 1)          # the tree already holds q_new (src/algorithm.py:145)
        out = o_cp(G, q_new, id_new, best_edge, radius, obstacles, *a, **k)
        rec["cp_parent"].append(int(out[1]))
        rec["cp_cost"].append(float(out[2]))
        return out

    def rewire_kdtree(G, q_new, id_new, radius, obstacles, *a, **k):
        before = dict(G.parent)
        out = o_rw(G, q_new, id_new, radius, obstacles, *a, **k)
        rec["rewired"].append(sorted(n for n, p in G.parent.items() if before.get(n) != p))
        return out

    def forced_removal(G, id_new, path):
        out = o_fr(G, id_new, path)
        rec["removed"][-1] = int(out)
        return out

    def through_obstacle(line, obstacles):
        hit = o_to(line, obstacles)
        rec["edge_calls"] += 1
        if len(rec["edges"]) < keep_edges:
            rec["edges"].append((float(line.p1[0]), float(line.p1[1]), float(line.p2[0]), float(line.p2[1]), bool(hit)))
        return hit

    alg.new_node = new_node
    alg.choose_parent_kdtree = choose_parent_kdtree
    alg.rewire_kdtree = rewire_kdtree
    alg.forced_removal = forced_removal
    alg.through_obstacle = through_obstacle


def make_map_and_graph(ref, start, goal, xy_range, node_radius, obstacles):
    """Reference `Map` + `Graph` for a list of (ox, oy, r) circles."""
    m = ref.map.Map((xy_range[0][1] - xy_range[0][0], xy_range[1][1] - xy_range[1][0]), start, goal, node_radius)
    m.add_obstacles([((ox, oy), r) for ox, oy, r in obstacles])
    g = ref.graph.Graph(start, goal, 0, 0, xy_range)
    return m, g


def tree_arrays(G):
    """Canonical arrays of a reference Graph: ids ascending, parent id (-1 = none), np.float64 edge cost."""
    ids = np.array(sorted(G.vertices), dtype=np.int32)
    x = np.array([float(G.vertices[i][0]) for i in ids], dtype=np.float64)
    y = np.array([float(G.vertices[i][1]) for i in ids], dtype=np.float64)
    parent = np.array([-1 if G.parent[i] is None else G.parent[i] for i in ids], dtype=np.int32)
    cost = np.array([float(G.cost[i]) for i in ids], dtype=np.float64)
    nchild = np.array([len(G.children[i]) for i in ids], dtype=np.int32)
    return dict(ids=ids, x=x, y=y, parent=parent, cost=cost, nchild=nchild)


def run_reference(mode, *, start, goal, xy_range, obstacles, words, iter_num, step_length, radius=None,
                  node_radius=1.0, bias=0.0, max_nodes=200, record=True, keep_edges=0, ascending_ball=True):
    """Run RRT / RRT_star / RRT_star_FN from the reference on the given word stream.

    Returns a dict with the final tree (`tree_arrays`), `iter`, `best_path` ids, `best_cost`,
    `words_used` and the per-iteration records.
    """
    from .philox import WordStreamRandom
    rng = WordStreamRandom(words)
    rec = {"keep_edges": keep_edges} if record else None
    with patched_reference(rng, rec, ascending_ball=ascending_ball) as ref, contextlib.redirect_stdout(io.StringIO()):
        m, g = make_map_and_graph(ref, start, goal, xy_range, node_radius, obstacles)
        alg = ref.algorithm
        best = {"path": [], "cost": float("inf")}
        if mode == "rrt":
            it = alg.RRT(g, iter_num, m, step_length, node_radius, bias)
            # RRT returns only `iter`; on success the path is plotted, not returned (src/algorithm.py:78-81)
            last = g.last_id
            if alg.check_solution(g, g.vertices[last], node_radius):
                p, c = alg.find_path(g, last, g.id_vertex[g.start])
                best = {"path": p, "cost": c}
        elif mode == "star":
            it, best = alg.RRT_star(g, iter_num, m, step_length, radius, node_radius, bias)
        elif mode == "fn":
            it, best = alg.RRT_star_FN(g, iter_num, m, step_length, radius, node_radius, max_nodes, bias)
        else:
            raise ValueError(mode)
    out = tree_arrays(g)
    out.update(iter=int(it), best_path=np.array(best["path"], dtype=np.int32), best_cost=float(best["cost"]),
               words_used=int(rng.cursor), last_id=int(g.last_id))
    if rec is not None:
        out["rec"] = rec
    out["graph"] = g
    out["map"] = m
    return out
