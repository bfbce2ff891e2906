"""Drop-in planners: RRT, RRT_star, RRT_star_FN, RRT_STAR_FND with the reference's signatures
(src/algorithm.py:55, :98, :190-191, :268-269), running on the GPU.

A call uploads the caller's `Graph` (whatever tree it already holds) to the device, grows it with the fused sm_100a
planner kernel through the C ABI (`rrtfnd_plan_batch`), writes the result back into the same `Graph` dictionaries and
returns what the reference returns. Randomness is the caller's: the kernel consumes the 32-bit words Python's global
`random` would have produced (MT19937), interpreted with CPython's own formulas for `random()`, `uniform()` and
`choice()`, and the global generator is left exactly where the reference would have left it. So

    random.seed(7); RRT_star(G, 2000, my_map, 25, 50, 20, bias=0.01)

gives the same tree, the same best path and the same subsequent `random.random()` as the reference.

There is no CPU fallback: without the CUDA library / a GPU these functions raise.
"""
import random
from functools import wraps
from time import perf_counter

import numpy as np

from .graph import Graph
from .map import Map
from ._structs import ST_DONE, ST_NEED_WORDS, ST_REGROW_NEED_WORDS, STATUS_NAMES


def time_function(func):
    """Print the wall time of a planner call like the reference's decorator does (src/algorithm.py:15-33)."""
    @wraps(func)
    def wrapper(*args, **kwargs):
        t0 = perf_counter()
        out = func(*args, **kwargs)
        print(f"Execution time of {func.__name__}: {round(perf_counter() - t0, 2)}")
        return out
    return wrapper


class PlannerError(RuntimeError):
    """The device planner stopped in a state the reference never returns from (it would loop or raise)."""


def _draw_words(n: int) -> np.ndarray:
    """The next n 32-bit outputs of the global Mersenne Twister, in generation order."""
    big = random.getrandbits(32 * n)
    return np.frombuffer(big.to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def _grow(mode: str, G: Graph, iter_num: int, map: Map, step_length: float, radius: float, node_radius: float,
          max_nodes: int, bias: float):
    """Run one planner on the GPU against the global `random` stream; returns the finished PlannerBatch."""
    import torch
    from .batch import PlannerBatch, fn_capacity
    if iter_num <= 0:
        return None
    ids, xy, parent, cost, root_id = G._export()
    n0 = len(ids)
    if mode == "fn":
        # RRT_star_FN counts only this call's nodes (n_of_nodes starts at 1, src/algorithm.py:209), so removals begin after
        # max_nodes insertions whatever the Graph already held: live nodes peak at n0 + min(iter_num, max_nodes)
        capacity = max(fn_capacity(max_nodes), n0 + min(int(iter_num), int(max_nodes)) + 2 + max(256, int(max_nodes) // 4))
    else:
        capacity = n0 + int(iter_num) + 2
    state = random.getstate()
    n_words = 256 + 16 * int(iter_num)
    words = _draw_words(n_words)
    b = PlannerBatch(mode, starts=[(float(G.start[0]), float(G.start[1]))], goals=[(float(G.goal[0]), float(G.goal[1]))],
                     obstacles=map.obstacles_c, rectangles=getattr(map, "obstacles_r", ()), xy_range=G.xy_range, iter_num=iter_num, step_length=step_length,
                     radius=radius, node_radius=map.node_radius, goal_node_radius=node_radius, bias=bias,
                     max_nodes=max_nodes, words=words[None, :], capacity=capacity, eval_choose_parent=False,
                     path_cap=max(1024, capacity), finish_cap=max(1024, capacity))
    b.reset()
    if n0 > 1 or int(ids[0]) != 0 or G.last_id != 0:
        b.load_tree(0, ids, xy, parent, cost, root_id, G.last_id)
        b.rebuild_grid()                                 # the caller's vertices may lie outside the box of xy_range
    while True:
        b.run()
        pl = b.planners()[0]
        st = int(pl["status"])
        if st != ST_NEED_WORDS:
            break
        more = _draw_words(n_words)                      # the stream continues where the last draw stopped
        words = np.concatenate([words, more])
        n_words *= 2
        b.words = torch.from_numpy(words[None, :].view(np.int32)).to(b.device)
        b.prm.words_per_planner = len(words)
    # leave the global generator where the reference would have left it: `cursor` words consumed
    random.setstate(state)
    used = int(pl["cursor"])
    if used:
        random.getrandbits(32 * used)
    if st != ST_DONE:
        raise PlannerError(f"{mode} planner stopped with status {st} ({STATUS_NAMES.get(st, '?')}) after "
                           f"{int(pl['iter'])} iterations")
    t = b.tree(0)
    keep = {i: G.vertices[i] for i in G.vertices}
    G._import(t["ids"], np.column_stack([t["x"], t["y"]]), t["parent"], t["cost"], int(pl["last_id"]), keep)
    return b


def _best_path(b) -> dict:
    pl = b.planners()[0]
    if int(pl["best_len"]) == 0:
        return {"path": [], "cost": float("inf")}
    return {"path": [int(i) for i in b.best_path(0)], "cost": np.float64(pl["best_cost"])}


@time_function
def RRT(G: Graph, iter_num: int, map: Map, step_length: float, node_radius: int, bias: float = .0,
        live_update: bool = False):
    """RRT (src/algorithm.py:54-94): grows until the goal is reached or iter_num extensions were made; returns the
    number of completed iterations (the pass that reaches the goal is not counted, :78-81)."""
    b = _grow("rrt", G, iter_num, map, step_length, 0.0, node_radius, 0, bias)
    return 0 if b is None else int(b.planners()[0]["iter"])


@time_function
def RRT_star(G, iter_num, map, step_length, radius, node_radius: int, bias=.0, live_update=False) -> tuple:
    """RRT* (src/algorithm.py:97-186): returns (iterations, {"path": ids goal-side first, "cost": cost})."""
    b = _grow("star", G, iter_num, map, step_length, radius, node_radius, 0, bias)
    if b is None:
        return 0, {"path": [], "cost": float("inf")}
    return int(b.planners()[0]["iter"]), _best_path(b)


@time_function
def RRT_star_FN(G, iter_num, map, step_length, radius, node_radius: int, max_nodes=200, bias=.0,
                live_update=False):
    """RRT*-FN (src/algorithm.py:189-265): RRT* with at most max_nodes live nodes (forced removal of childless nodes)."""
    b = _grow("fn", G, iter_num, map, step_length, radius, node_radius, max_nodes, bias)
    if b is None:
        return 0, {"path": [], "cost": float("inf")}
    return int(b.planners()[0]["iter"]), _best_path(b)


def init_movement():
    """Stub in the reference (src/algorithm.py:290-291)."""


def RRT_STAR_FND(G: Graph, iter_num: int, map: Map, step_length: int, radius: float, node_radius: int,
                 max_nodes: int = 200, bias: float = .0, live_update: bool = False) -> None:
    """RRT*-FND entry point as shipped (src/algorithm.py:268-287): an RRT*-FN run followed by the movement stub."""
    RRT_star_FN(G=G, iter_num=iter_num, map=map, step_length=step_length, radius=radius, node_radius=node_radius,
                max_nodes=max_nodes, bias=bias, live_update=live_update)
    init_movement()


# ---- FND tree surgery (src/algorithm.py:294-434) on Graph objects --------------------------------------------------

def _surgery_batch(G: Graph, map: Map, path):
    """Upload G (possibly a forest) with `path` as the stored path; returns the PlannerBatch and the kept positions."""
    from .batch import PlannerBatch
    ids, xy, parent, cost, root_id = G._export()
    cap = len(ids) + 2
    b = PlannerBatch("star", starts=[(float(G.start[0]), float(G.start[1]))], goals=[(float(G.goal[0]), float(G.goal[1]))],
                     obstacles=map.obstacles_c, rectangles=getattr(map, "obstacles_r", ()), xy_range=G.xy_range, iter_num=1, step_length=1.0, radius=1.0,
                     node_radius=map.node_radius, capacity=cap, path_cap=max(64, len(path) + 2), finish_cap=8, grid=False)
    b.reset()
    b.load_tree(0, ids, xy, parent, cost, root_id, G.last_id)
    b.set_best_path(0, path)
    return b, {i: G.vertices[i] for i in G.vertices}


def _download(G: Graph, b, keep):
    t = b.tree(0)
    pl = b.planners()[0]
    G._import(t["ids"], np.column_stack([t["x"], t["y"]]), t["parent"], t["cost"], G.last_id, keep)
    root_id = b.slot_id(0, int(pl["root_slot"]))
    G.start = G.vertices[root_id]


def select_branch(G: Graph, current_node: int, path: list, map: Map) -> list:
    """Re-root the tree at `current_node` (a node of `path`), deleting everything outside its subtree; returns the
    remaining path (src/algorithm.py:294-325)."""
    b, keep = _surgery_batch(G, map, path)
    st = int(b.select_branch([current_node])[0])
    if st == -1:
        raise ValueError(f"{current_node} is not in list")          # path.index(current_node) in the reference
    if st == -2:
        raise KeyError(None)                                        # G.children[None] in the reference
    _download(G, b, keep)
    return [int(i) for i in b.best_path(0)]


def valid_path(G: Graph, path: list, map: Map, previous_root: int) -> int:
    """Delete path nodes that collide with obstacles, and what hangs off them; returns the root of the separated
    tree (src/algorithm.py:344-367)."""
    b, keep = _surgery_batch(G, map, path)
    sep, err = b.valid_path([previous_root])
    _download(G, b, keep)
    if int(err[0]):
        raise IndexError("list index out of range")                 # G.children[id_node][0] at :358
    return int(sep[0])


def reconnect(G: Graph, path: list, map: Map, step_size: float, id_root: int, id_separate_root: int,
              last_node: int):
    """Link the separated tree back to the first visible root-tree node within `step_size`; returns (path, cost) from
    `last_node` to the root, or None (src/algorithm.py:370-404)."""
    b, keep = _surgery_batch(G, map, [last_node] + [i for i in path if i != last_node])
    linked = int(b.reconnect([id_separate_root], step_size)[0])
    if linked < 0:
        return None
    _download(G, b, keep)
    return [int(i) for i in b.best_path(0)], np.float64(b.planners()[0]["best_cost"])


def regrow(G: Graph, map: Map, step_length: float, radius: float, id_root: int, id_separate_root: int, last_node: int,
           bias: float, _n_words: int = 16 * 500 + 256):
    """Grow the root tree (at most 500 extensions, drawing from the global `random` like the planners) until a new node
    can adopt the separated tree; returns None like the reference (src/algorithm.py:482-528). `radius` only feeds the
    reference's choose_parent call whose result it drops. `_n_words` (not in the reference) sizes the chunks in which
    words of the global generator are handed to the device; the result does not depend on it."""
    import torch
    from .batch import PlannerBatch
    ids, xy, parent, cost, root_id = G._export()
    cap = len(ids) + 502
    state = random.getstate()
    n_words = max(128, int(_n_words))
    words = _draw_words(n_words)
    b = PlannerBatch("star", starts=[(float(G.start[0]), float(G.start[1]))], goals=[(float(G.goal[0]), float(G.goal[1]))],
                     obstacles=map.obstacles_c, rectangles=getattr(map, "obstacles_r", ()), xy_range=G.xy_range, iter_num=1, step_length=step_length, radius=radius,
                     node_radius=map.node_radius, bias=bias, capacity=cap, path_cap=max(64, cap), finish_cap=8,
                     words=words[None, :])
    b.reset()
    b.load_tree(0, ids, xy, parent, cost, root_id, G.last_id)
    b.rebuild_grid()
    b.set_best_path(0, [last_node])
    left = 500                                           # the reference's own cap on extensions (:486)
    while True:
        r, its = b.regrow([id_separate_root], max_iter=left)
        if int(r[0]) != ST_REGROW_NEED_WORDS:
            break
        left -= int(its[0])                              # rejected samples ate the buffer: the stream continues
        words = np.concatenate([words, _draw_words(n_words)])
        b.words = torch.from_numpy(words[None, :].view(np.int32)).to(b.device)
        b.prm.words_per_planner = len(words)
    pl = b.planners()[0]
    random.setstate(state)
    if int(pl["cursor"]):
        random.getrandbits(32 * int(pl["cursor"]))
    if int(r[0]) < 0:
        raise PlannerError(f"regrow stopped with status {int(r[0])}")
    t = b.tree(0)
    G._import(t["ids"], np.column_stack([t["x"], t["y"]]), t["parent"], t["cost"], int(pl["last_id"]), {i: G.vertices[i] for i in G.vertices})
    return None


# ---- plotting (src/algorithm.py:638-663, :763-775): matplotlib is imported on use, like every visual part ------------

def plot_graph(G: Graph, obstacles: list, xy_range: tuple = None):
    """Draw the tree, start / goal and the circular obstacles (src/algorithm.py:638-663)."""
    import matplotlib.pyplot as plt
    xr, yr = (G.xy_range if xy_range is None else xy_range)
    for i, q in G.parent.items():
        if q is not None:
            (x0, y0), (x1, y1) = G.vertices[i], G.vertices[q]
            plt.plot((x0, x1), (y0, y1), color="orange", marker="o", markersize=2)
    ax = plt.gca()
    for centre, r in obstacles:
        ax.add_patch(plt.Circle((float(centre[0]), float(centre[1])), float(r), color="black"))
    ax.add_patch(plt.Circle(G.start, 2, color="blue"))
    ax.add_patch(plt.Circle(G.goal, 2, color="green"))
    ax.set_aspect("equal", adjustable="box")
    plt.xlim(xr[0], xr[1])
    plt.ylim(yr[0], yr[1])


def plot_path(G: Graph, path: list, title: str = "", cost: float = float("inf")):
    """Draw `path` (ids, goal side first) starting from the goal position, and put the cost in the title
    (src/algorithm.py:763-775)."""
    import matplotlib.pyplot as plt
    prev = G.goal
    for node in path:
        pos = G.vertices[node]
        plt.plot((prev[0], pos[0]), (prev[1], pos[1]), c="#057af7", linewidth=2)
        prev = pos
    plt.title(title + f" cost: {round(cost, 2)}")


# ---- helpers user code calls directly --------------------------------------------------------------------------------

def calc_distance(p1: tuple, p2: tuple) -> float:
    """Euclidean distance through the same numpy call as the reference (src/algorithm.py:954-961), so a host-side
    value carries the same rounding; the planners use the device form sqrt(fma(dy, dy, dx*dx))."""
    return np.linalg.norm(np.array(p1, dtype=np.float64) - np.array(p2, dtype=np.float64))


def check_solution(G: Graph, q_new: tuple, node_radius: int) -> bool:
    """src/algorithm.py:749-760"""
    return bool(calc_distance(q_new, G.goal) < 2 * node_radius)


def find_path(G: Graph, from_node: int, root_node: int) -> tuple:
    """Ids from `from_node` to the root and the cost summed in that order (src/algorithm.py:778-797); a broken
    chain yields the partial path, like the reference's swallowed exception."""
    path, cost, node = [], 0, from_node
    while node != root_node:
        if node not in G.parent or G.parent[node] is None:
            return path + ([node] if node in G.parent else []), cost
        path.append(node)
        cost = cost + G.cost[node]
        node = G.parent[node]
    path.append(root_node)
    return path, cost


def through_obstacle_batch(p1, p2, obstacles) -> np.ndarray:
    """through_obstacle(Line(p1[i], p2[i]), obstacles) for many segments at once (src/algorithm.py:581-591) on the GPU.
    p1, p2: [n][2]; obstacles: the reference's list of [(x, y), r]."""
    import ctypes as C
    import torch
    from . import _abi
    from .batch import obstacle_table
    a = np.ascontiguousarray(p1, dtype=np.float64).reshape(-1, 2)
    c = np.ascontiguousarray(p2, dtype=np.float64).reshape(-1, 2)
    n = len(a)
    if n == 0 or len(obstacles) == 0:
        return np.zeros(n, dtype=bool)
    dev = torch.device("cuda", torch.cuda.current_device())
    cols = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (a[:, 0], a[:, 1], c[:, 0], c[:, 1])]
    tab = torch.from_numpy(obstacle_table(obstacles)).to(dev)
    out = torch.zeros(n, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _abi.check(_abi.lib.rrtfnd_edges_vs_circles(*[C.c_void_p(t.data_ptr()) for t in cols], n, C.c_void_p(tab.data_ptr()),
                                                len(obstacles), C.c_void_p(out.data_ptr()), st), "rrtfnd_edges_vs_circles")
    return out.cpu().numpy().astype(bool)


def through_obstacle(line, obstacles: list) -> bool:
    """through_obstacle for one `Line`-like object with .p1 / .p2 (src/algorithm.py:581-591)."""
    return bool(through_obstacle_batch([line.p1], [line.p2], obstacles)[0])


# ---- the reference's per-iteration helpers on Graph objects ------------------------------------------------------------
# Users who drive the loop themselves (src/algorithm.py:139-149) call these directly. Each keeps the reference's
# signature, return value and side effects on `G`; the work that dominates the reference's profile - every
# through_obstacle call (74 % of its time), the visibility search of nearest_node_kdtree, the Graph.get_cost walks of the
# candidates - runs on the GPU through the C-ABI operators (rrtfnd_edges_vs_circles, rrtfnd_nearest_visible,
# rrtfnd_chain_cost); what stays on the host is the bookkeeping on the caller's own dictionaries and a handful of scalar
# operations per call, written with the reference's expressions on the same number types.

def _device_graph(G: Graph, obstacles: list):
    """G uploaded as a one-planner batch (rows in ascending id): returns (batch, ids, slot_of)."""
    from .batch import PlannerBatch
    ids, xy, parent, cost, root_id = G._export()
    b = PlannerBatch("star", starts=[(float(G.start[0]), float(G.start[1]))], goals=[(float(G.goal[0]), float(G.goal[1]))],
                     obstacles=obstacles, xy_range=G.xy_range, iter_num=1, step_length=1.0, radius=1.0, node_radius=0.0,
                     capacity=len(ids) + 2, path_cap=8, finish_cap=8, grid=False)
    b.reset()
    b.load_tree(0, ids, xy, parent, cost, root_id, G.last_id)
    return b, ids, {int(i): k for k, i in enumerate(ids.tolist())}


def _chain_costs(G: Graph, node_ids, obstacles=()) -> np.ndarray:
    """Graph.get_cost for many nodes at once (src/graph.py:34-46): leaf -> root fp64 sums on the GPU, np.float64 each."""
    import ctypes as C
    import torch
    from . import _abi
    node_ids = [int(i) for i in node_ids]
    if not node_ids:
        return np.zeros(0, dtype=np.float64)
    b, _, slot_of = _device_graph(G, list(obstacles))
    sl = torch.tensor([slot_of[i] for i in node_ids], dtype=torch.int32, device=b.device)
    pi = torch.zeros(len(node_ids), dtype=torch.int32, device=b.device)
    out = torch.zeros(len(node_ids), dtype=torch.float64, device=b.device)
    bt = b._batch()
    _abi.check(_abi.lib.rrtfnd_chain_cost(C.byref(bt), b.prm.capacity, pi.data_ptr(), sl.data_ptr(), len(node_ids),
                                          out.data_ptr(), b._stream()), "rrtfnd_chain_cost")
    return out.cpu().numpy()


def intersection_circle(line, circle: list) -> bool:
    """Does the segment `line` (a Line-like object with .p1 / .p2) hit the circle `[(x, y), r]`? The reference's predicate
    with all its quirks - infinite vertical lines, x-only strict between test (src/algorithm.py:555-578)."""
    return bool(through_obstacle_batch([line.p1], [line.p2], [circle])[0])


def calc_delta(line, circle: list) -> tuple:
    """(discriminant, a, b) of the quadratic in x that intersects the infinite line y = dir * x + const_term with the circle
    `[(x0, y0), r]` (src/algorithm.py:607-621). Host scalars in the reference's operation order, Python `**` included; the
    kernels evaluate the same expressions per circle with the constant part of `c` precomputed (arith.cuh)."""
    (x0, y0), r = (circle[0][0], circle[0][1]), circle[1]
    m, n = line.dir, line.const_term
    a = (1 + m ** 2)
    b = 2 * (-x0 + m * n - m * y0)
    c = -r ** 2 + x0 ** 2 + y0 ** 2 - 2 * y0 * n + n ** 2
    return b ** 2 - 4 * a * c, a, b


def delta_solutions(delta: float, a: float, b: float) -> tuple:
    """The two roots' x as `[x, None]` lists, larger root first for a > 0 (src/algorithm.py:594-604); the caller fills in y."""
    root = delta ** 0.5
    return [(-b + root) / (2 * a), None], [(-b - root) / (2 * a), None]


def is_between(pb, p1, p2) -> bool:
    """Is pb's x STRICTLY inside the x-interval of (p1, p2)? y is not looked at (src/algorithm.py:624-635)."""
    return pb[0] > min(p1[0], p2[0]) and pb[0] < max(p1[0], p2[0])


def steer(to_vertex: tuple, from_vertex: tuple, max_length: float) -> tuple:
    """Point at most `max_length` from `from_vertex` towards `to_vertex`; `to_vertex` itself, untouched, when it is closer
    (src/algorithm.py:731-746). Six scalar operations in the reference's order (the planner kernel does the same on the
    device)."""
    distance = calc_distance(to_vertex, from_vertex)
    ux, uy = (to_vertex[0] - from_vertex[0]) / distance, (to_vertex[1] - from_vertex[1]) / distance
    moved = (from_vertex[0] + ux * max_length, from_vertex[1] + uy * max_length)
    return moved if distance > max_length else to_vertex


def nearest_node_kdtree(G: Graph, vertex: tuple, obstacles: list, separate_tree_nodes: list = (), kdtree=None):
    """Nearest node of G that sees `vertex`: (position as ndarray, id); a vertex that already exists is returned as is;
    (np.array(vertex), None) when no node sees it (src/algorithm.py:666-699). The j-th-nearest-until-visible search runs on
    the GPU over all of G's nodes (the reference builds `kdtree` from exactly those, :41-43; `separate_tree_nodes` is
    ignored there as well)."""
    import ctypes as C
    import torch
    from . import _abi
    try:
        return np.array(vertex), G.id_vertex[vertex]
    except KeyError:
        pass
    b, ids, _ = _device_graph(G, obstacles)
    q = torch.tensor([[float(vertex[0]), float(vertex[1])]], dtype=torch.float64, device=b.device)
    slot = torch.zeros(1, dtype=torch.int32, device=b.device)
    bt = b._batch()
    _abi.check(_abi.lib.rrtfnd_nearest_visible(C.byref(bt), b.prm.capacity, q.data_ptr(), b.prm.n_obst, slot.data_ptr(), None,
                                               b._stream()), "rrtfnd_nearest_visible")
    k = int(slot.item())
    if k < 0:
        return np.array(vertex), None
    node = int(ids[k])
    return np.array(G.vertices[node], dtype=np.float64), node


def nearest_node(G: Graph, vertex: tuple, obstacles: list, separate_tree_nodes: list = ()):
    """Brute-force twin (src/algorithm.py:702-728): among the nodes outside `separate_tree_nodes` whose
    Line(node, vertex) is free, the closest by calc_distance, first in dictionary order on ties; (None, None) if there
    is none. All the segment tests are one GPU call."""
    try:
        return vertex, G.id_vertex[vertex]
    except KeyError:
        pass
    cand = [(i, v) for i, v in G.vertices.items() if i not in separate_tree_nodes]
    if not cand:
        return None, None
    blocked = through_obstacle_batch([v for _, v in cand], [vertex] * len(cand), obstacles)
    best, best_id, best_pos = float("inf"), None, None
    for (i, v), hit in zip(cand, blocked):
        if hit:
            continue
        d = calc_distance(v, vertex)
        if d < best:
            best, best_id, best_pos = d, i, v
    return best_pos, best_id


def new_node(G: Graph, map: Map, obstacles: list, step_length: float, bias: float):
    """One attempt of the planners (src/algorithm.py:36-51): sample, reject if occupied, nearest visible node, steer,
    add the vertex. Returns (q_new, id_new, id_near, distance) or four Nones."""
    q_rand = G.random_node(bias=bias)
    if map.is_occupied_c(q_rand):
        return None, None, None, None
    q_near, id_near = nearest_node_kdtree(G, q_rand, obstacles)
    if q_near is None or (q_rand == q_near).all():
        return None, None, None, None
    q_new = steer(q_rand, q_near, step_length)
    id_new = G.add_vertex(q_new)
    return q_new, id_new, id_near, calc_distance(q_new, q_near)


def _ball_candidates(G: Graph, q_new, radius, kdtree):
    """(ids, positions, np.linalg.norm distances) of the caller's ball query, in the caller's k-d tree's own order - the
    order choose_parent_kdtree's result depends on (src/algorithm.py:837-842)."""
    i = kdtree.query_ball_point(q_new, r=radius)
    pos = kdtree.data[i]
    ids = [G.id_vertex[p[0], p[1]] for p in pos]
    return ids, pos, np.linalg.norm(pos - q_new, axis=1)


def choose_parent_kdtree(G: Graph, q_new: tuple, id_new: int, best_edge: tuple, radius: float, obstacles: list,
                         separate_tree_nodes: list = (), kdtree=None) -> tuple:
    """The edge (id_new, parent, cost) the reference returns (src/algorithm.py:822-852), including its side effect: each
    accepted candidate overwrites G.cost[id_new] while G.parent[id_new] stays, so later comparisons mix the new edge
    with the old chain. Segment tests and the candidates' get_cost walks run on the GPU."""
    ids, pos, costs = _ball_candidates(G, q_new, radius, kdtree)
    if not ids:
        return best_edge
    blocked = through_obstacle_batch(pos, [q_new] * len(ids), obstacles)
    walk_on_host = len(G.children.get(id_new, ())) > 0          # id_new above a candidate: its chain changes as we go
    chain = None if walk_on_host else _chain_costs(G, ids)
    for k, (id_ver, cost) in enumerate(zip(ids, costs)):
        if blocked[k] or id_ver == id_new:               # the ball holds q_new itself: x > x + 0 is never true (:848)
            continue
        via = (G.get_cost(id_ver) if walk_on_host else chain[k]) + cost
        if G.get_cost(id_new) > via:
            G.cost[id_new] = cost
            best_edge = (id_new, id_ver, cost)
    return best_edge


def choose_parent(G: Graph, q_new: tuple, id_new: int, best_edge: tuple, radius: float, obstacles: list,
                  separate_tree_nodes: list = ()) -> tuple:
    """Brute-force twin (src/algorithm.py:855-879): every vertex in dictionary order, calc_distance rounded to three
    decimals against the radius."""
    cand = [(i, v, calc_distance(q_new, v)) for i, v in G.vertices.items() if i != id_new]
    cand = [c for c in cand if not round(c[2], 3) > radius]
    if not cand:
        return best_edge
    blocked = through_obstacle_batch([v for _, v, _ in cand], [q_new] * len(cand), obstacles)
    walk_on_host = len(G.children.get(id_new, ())) > 0
    chain = None if walk_on_host else _chain_costs(G, [i for i, _, _ in cand])
    for k, (id_ver, _, d) in enumerate(cand):
        if blocked[k]:
            continue
        via = (G.get_cost(id_ver) if walk_on_host else chain[k]) + d
        if G.get_cost(id_new) > via:
            G.cost[id_new] = d
            best_edge = (id_new, id_ver, d)
    return best_edge


def _rewire_apply(G: Graph, id_new: int, cand, blocked):
    """Re-parent under id_new every free candidate whose cost-to-come improves (src/algorithm.py:905-910, :926-935).
    Decisions are taken against the tree as it is before the call: a candidate's chain only changes when one of its
    ancestors is re-parented, and that only ever lowers its cost further below the threshold that was not met."""
    free = [k for k in range(len(cand)) if not blocked[k]]
    if not free:
        return
    costs = _chain_costs(G, [cand[k][0] for k in free] + [id_new])
    c_new = costs[-1]
    for k, c_ver in zip(free, costs[:-1]):
        id_ver, d = cand[k][0], cand[k][1]
        if c_ver > c_new + d:
            parent = G.parent[id_ver]
            G.children[parent].remove(id_ver)
            G.parent[id_ver] = id_new
            G.children[id_new].append(id_ver)
            G.cost[id_ver] = d


def rewire_kdtree(G: Graph, q_new: tuple, id_new: int, radius: float, obstacles: list, kdtree=None):
    """Rewire step of RRT* over the caller's ball query (src/algorithm.py:882-910)."""
    ids, pos, costs = _ball_candidates(G, q_new, radius, kdtree)
    if not ids:
        return
    blocked = through_obstacle_batch(pos, [q_new] * len(ids), obstacles)
    blocked = [bool(h) or i == id_new for h, i in zip(blocked, ids)]       # the new node itself never passes the strict test
    _rewire_apply(G, id_new, list(zip(ids, costs)), blocked)


def rewire(G: Graph, q_new: tuple, id_new: int, radius: float, obstacles: list):
    """Brute-force twin (src/algorithm.py:913-935): every vertex but the start node and the new one, calc_distance
    against the radius."""
    root = G.id_vertex[G.start]
    cand = [(i, calc_distance(q_new, v), v) for i, v in G.vertices.items() if i != root and i != id_new]
    cand = [c for c in cand if not c[1] > radius]
    if not cand:
        return
    blocked = through_obstacle_batch([v for _, _, v in cand], [q_new] * len(cand), obstacles)
    _rewire_apply(G, id_new, cand, blocked)


def forced_removal(G: Graph, id_new: int, path: list) -> int:
    """Delete a random childless node other than id_new and the goal-side end of `path`; returns its id
    (src/algorithm.py:800-819). Draws from the global `random` exactly like the reference (random.choice per draw)."""
    keep = path[0] if path else -1
    childless = [node for node, children in G.children.items() if len(children) == 0]
    victim = random.choice(childless)
    while victim == id_new or victim == keep:
        victim = random.choice(childless)
    G.remove_vertex(victim)
    return victim


def remove_children(G: Graph, id_node: int, path: list) -> None:
    """Delete every descendant of id_node that hangs off a child outside `path` (src/algorithm.py:328-341); children on
    the path, and what hangs below them, stay."""
    stack = [c for c in G.children[id_node] if c not in path]
    order = []
    while stack:                                    # subtree of each off-path child, parents before children
        n = stack.pop()
        order.append(n)
        stack.extend(c for c in G.children[n] if c not in path)
    for n in reversed(order):                       # leaves first, as the recursion deletes them; an off-path node goes
        G.remove_vertex(n)                          # even when an on-path child of it stays behind (as in the reference)


def reconnect_ancestors(G: Graph, id_node: int) -> None:
    """Make id_node the root of its tree by turning every edge on its way up around (src/algorithm.py:531-543; what
    reconnect and regrow do to the separate tree once a link is found, fnd.cu `reroot`). As in the reference, the last flip is
    half done - the old root gets its new parent but is not entered in that parent's children list, nor is the parent taken out
    of the old root's -, costs are not touched, and id_node keeps its old parent pointer (the callers overwrite it with the
    link they found)."""
    child, up = id_node, G.parent[id_node]
    if up is None:
        return
    while G.parent[up] is not None:
        above = G.parent[up]
        G.parent[up] = child
        G.children[child].append(up)
        G.children[up].remove(child)
        child, up = up, above
    G.parent[up] = child


def delete_ancestors(G: Graph, id_node: int) -> None:
    """Walk from id_node to its root and strip every ancestor of ALL its subtrees (remove_children with an empty path):
    id_node goes with the first of them, and in the end only the root is left. Unused in the reference (its one call is
    commented out, src/algorithm.py:514, :546-551); kept for name parity."""
    node = id_node
    up = G.parent[node]
    while up is not None:
        node = up
        remove_children(G, node, [])
        up = G.parent[node]


def check_for_tree_associativity(G: Graph, root_node: int, node_to_check: int) -> bool:
    """Does node_to_check hang (through any number of parents) from root_node? (src/algorithm.py:407-420)"""
    node = node_to_check
    while G.parent[node] is not None:
        node = G.parent[node]
    return node == root_node


def get_distance_dict(G: Graph, node_to_check: int, indeces_to_check: list = None) -> dict:
    """{id: distance to node_to_check} for every vertex with the reference's expansion sqrt(|x|^2 - 2 x.y + |y|^2)
    (src/algorithm.py:938-951; cancellation can make an entry NaN, as there). The third parameter is unused in the
    reference as well; it gets a default here because get_near_nodes calls the function without it (:431)."""
    pts = np.array([v for v in G.vertices.values()])
    q = np.array(G.vertices[node_to_check]).reshape(-1, 2)
    x2 = np.sum(pts ** 2, axis=1).reshape(-1, 1)
    y2 = np.sum(q ** 2, axis=1).reshape(-1, 1)
    d = np.sqrt(x2 - 2 * np.matmul(pts, q.T) + y2.T)
    return {i: row[0] for i, row in zip(G.vertices, d)}


def get_near_nodes(G: Graph, node_to_check: int, radius: float, root_node: int) -> list:
    """Ids within `radius` of node_to_check that belong to the tree rooted at root_node (src/algorithm.py:423-434)."""
    return [i for i, d in get_distance_dict(G, node_to_check).items()
            if d <= radius and check_for_tree_associativity(G, root_node, i)]
