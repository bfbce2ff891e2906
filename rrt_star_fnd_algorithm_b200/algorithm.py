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
                     obstacles=map.obstacles_c, xy_range=G.xy_range, iter_num=iter_num, step_length=step_length,
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
                     obstacles=map.obstacles_c, xy_range=G.xy_range, iter_num=1, step_length=1.0, radius=1.0,
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
                     obstacles=map.obstacles_c, xy_range=G.xy_range, iter_num=1, step_length=step_length, radius=radius,
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
