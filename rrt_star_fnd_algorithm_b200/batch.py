"""PlannerBatch: P independent planners resident on one GPU as structure-of-arrays (torch tensors own the
memory, the CUDA library does the work). This is the batch API behind the drop-in entry points in
planners.py and the object bench.py drives."""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._structs import (
    Batch, NodeGrid, ObstGrid, Params, FLAG_NODE_GRID, PLANNER_BYTES, PLANNER_DTYPE, TRACE_BYTES, TRACE_DTYPE,
    MODE_RRT, MODE_STAR, MODE_FN, STREAM_WORDS, STREAM_PHILOX, FLAG_EVAL_CHOOSE_PARENT, FLAG_TRACE, FLAG_RECTANGLES,
    ST_DONE, ST_RUNNING, ST_NEED_WORDS, STATUS_NAMES,
)

_MODES = {"rrt": MODE_RRT, "star": MODE_STAR, "fn": MODE_FN, MODE_RRT: MODE_RRT, MODE_STAR: MODE_STAR, MODE_FN: MODE_FN}


def obstacle_table(obstacles, rectangles=()) -> np.ndarray:
    """Rows of the device obstacle table from the reference's circle list `[(pos, r), ...]` (or (ox, oy, r) triples) and,
    as an EXTENSION the reference only declares (Map.obstacles_r, src/map.py:18), axis-aligned rectangles
    `[(cx, cy), (hx, hy)]` (centre, half extents; or (cx, cy, hx, hy)).

    Circle rows are (ox, oy, r, K): K is the circle-only part of calc_delta's `c` (src/algorithm.py:619), evaluated with
    the reference's own expression on the caller's own number types: `-r ** 2 + x0 ** 2 + y0 ** 2`. Rectangle rows are
    (cx, cy, -hx, hy): the negative third column marks them (include/rrtfnd.h).
    """
    rows = []
    for ob in obstacles:
        if len(ob) == 4:                      # a row that already carries K (or a rectangle row)
            rows.append(tuple(float(v) for v in ob))
            continue
        if len(ob) == 2:
            (x0, y0), r = (ob[0][0], ob[0][1]), ob[1]
        else:
            x0, y0, r = ob
        K = -r ** 2 + x0 ** 2 + y0 ** 2
        rows.append((float(x0), float(y0), float(r), float(K)))
    for rc in rectangles:
        (cx, cy), (hx, hy) = ((rc[0][0], rc[0][1]), (rc[1][0], rc[1][1])) if len(rc) == 2 else ((rc[0], rc[1]), (rc[2], rc[3]))
        if not (hx > 0 and hy >= 0):
            raise ValueError("a rectangle needs half extents hx > 0, hy >= 0")
        rows.append((float(cx), float(cy), -float(hx), float(hy)))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 4)


def build_obstacle_grid(table4: np.ndarray, node_radius: float, bounds, cell: float = 0.0):
    """Host-side culling grid (rrtfnd_obst_grid_build_host4) over the rows of the device table [M][4]: returns (ObstGrid
    without pointers, 4 int32 arrays).

    `bounds` = (lo_x, hi_x, lo_y, hi_y) must contain every coordinate the planners can produce."""
    ob = np.asarray(table4, dtype=np.float64)
    if ob.ndim == 2 and ob.shape[1] == 3:                     # plain (ox, oy, r) circles
        ob = np.column_stack([ob, np.zeros(len(ob))])
    ob = np.ascontiguousarray(ob).reshape(-1, 4)
    g = ObstGrid()
    cnt = (C.c_int64 * 2)()
    args = (ob.ctypes.data, len(ob), float(node_radius), float(bounds[0]), float(bounds[1]), float(bounds[2]),
            float(bounds[3]), float(cell), C.byref(g))
    _abi.check(_abi.lib.rrtfnd_obst_grid_build_host4(*args, None, None, None, None, cnt), "rrtfnd_obst_grid_build_host4")
    if g.nx == 0:
        return g, None
    nc = g.nx * g.ny
    arrs = [np.zeros(nc + 1, np.int32), np.zeros(max(int(cnt[0]), 1), np.int32),
            np.zeros(nc + 1, np.int32), np.zeros(max(int(cnt[1]), 1), np.int32)]
    _abi.check(_abi.lib.rrtfnd_obst_grid_build_host4(*args, *[a.ctypes.data for a in arrs], cnt),
               "rrtfnd_obst_grid_build_host4")
    return g, arrs


def fn_capacity(max_nodes: int) -> int:
    """Slots for an RRT*-FN planner: the cap, the transient extra node, and tombstone slack between compactions."""
    return int(max_nodes) + 2 + max(256, int(max_nodes) // 4)


class PlannerBatch:
    def __init__(self, mode, *, starts, goals, obstacles, xy_range, iter_num, step_length, radius=0.0, rectangles=(),
                 node_radius=1.0, goal_node_radius=None, bias=0.0, max_nodes=200, seeds=None, stream_ids=None, words=None,
                 capacity=None, near_cap=64, finish_cap=1024, path_cap=1024, trace=False,
                 eval_choose_parent=True, commit_choose_parent=False, max_attempts=0, device=None, launch=0,
                 grid=True, grid_cell=0.0, node_grid=False, node_grid_cell=0.0, node_grid_min_nodes=4096,
                 node_grid_cand_cap=8192):
        if not torch.cuda.is_available():
            raise RuntimeError("PlannerBatch needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        starts = np.asarray(starts, dtype=np.float64).reshape(-1, 2)
        goals = np.asarray(goals, dtype=np.float64).reshape(-1, 2)
        P = len(starts)
        assert len(goals) == P
        self.P = P
        m = _MODES[mode]
        if capacity is None:
            capacity = fn_capacity(max_nodes) if m == MODE_FN else int(iter_num) + 2
        prm = Params()
        prm.mode, prm.iter_num, prm.max_nodes, prm.capacity = m, int(iter_num), int(max_nodes), int(capacity)
        self.obst_host = obstacle_table(obstacles, rectangles)
        prm.n_obst = len(self.obst_host)
        prm.near_cap, prm.finish_cap, prm.path_cap = int(near_cap), int(finish_cap), int(path_cap)
        prm.flags = (FLAG_EVAL_CHOOSE_PARENT if eval_choose_parent else 0) | (2 if commit_choose_parent else 0) | \
            (FLAG_TRACE if trace else 0)
        prm.trace_cap = int(iter_num) + 1 if trace else 0
        prm.max_attempts = int(max_attempts)
        prm.reserved = int(launch)   # planners per CTA (1, 2, 4) + 16 * xy-ring stages (2, 4) + 256 * resident warps per SM (16, 24, 32); 0 fields = automatic
        prm.step_length, prm.radius, prm.node_radius, prm.bias = float(step_length), float(radius or 0.0), \
            float(node_radius), float(bias)
        prm.goal_node_radius = float(node_radius if goal_node_radius is None else goal_node_radius)
        (prm.x_lo, prm.x_hi), (prm.y_lo, prm.y_hi) = [(float(a), float(b)) for a, b in xy_range]
        self.prm = prm
        self._flag_rectangles()

        dev = self.device
        pl = np.zeros(P, dtype=PLANNER_DTYPE)
        pl["start_x"], pl["start_y"] = starts[:, 0], starts[:, 1]
        pl["goal_x"], pl["goal_y"] = goals[:, 0], goals[:, 1]
        if words is not None:
            w = np.ascontiguousarray(words, dtype=np.uint32).reshape(P, -1)
            prm.stream_mode, prm.words_per_planner = STREAM_WORDS, w.shape[1]
            self.words = torch.from_numpy(w.view(np.int32)).to(dev)
        else:
            prm.stream_mode, prm.words_per_planner = STREAM_PHILOX, 0
            self.words = None
            pl["seed"] = np.asarray(seeds if seeds is not None else np.arange(P), dtype=np.uint64)
            pl["stream_id"] = np.asarray(stream_ids if stream_ids is not None else np.zeros(P), dtype=np.uint64)
        self.planners_t = torch.from_numpy(pl.view(np.uint8).reshape(P, PLANNER_BYTES).copy()).to(dev)
        self._problems_t = self.planners_t.clone()      # start / goal / stream keys as given (select_branch moves `start`)
        Cn = prm.capacity
        self.xy = torch.empty((P, Cn, 2), dtype=torch.float64, device=dev)
        self.ecost = torch.empty((P, Cn), dtype=torch.float64, device=dev)
        self.parent = torch.empty((P, Cn), dtype=torch.int32, device=dev)
        self.nchild = torch.empty((P, Cn), dtype=torch.int32, device=dev)
        self.id = torch.empty((P, Cn), dtype=torch.int32, device=dev)
        self.nflags = torch.zeros((P, Cn), dtype=torch.uint8, device=dev)
        # rrtfnd_walk_t scratch: the node-grid kernel of large single trees walks parent chains over 16-byte records
        self.walk = torch.empty((P, Cn, 2), dtype=torch.float64, device=dev) if node_grid else None
        self.finish = torch.empty((P, prm.finish_cap), dtype=torch.int32, device=dev)
        self.best_path_t = torch.empty((P, prm.path_cap), dtype=torch.int32, device=dev)
        self.obst = torch.from_numpy(self.obst_host if len(self.obst_host) else np.zeros((1, 4))).to(dev)
        # culling grid over the obstacles; its box bounds the sampling range and every start / goal
        self.grid = ObstGrid()
        self._grid_t = None
        if grid and len(self.obst_host):
            pts = np.concatenate([starts, goals, np.asarray(xy_range, dtype=np.float64).T])
            bounds = (pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max())
            self._bounds = bounds
            g, arrs = build_obstacle_grid(self.obst_host, prm.node_radius, bounds, grid_cell)
            if arrs is not None:
                self._grid_t = [torch.from_numpy(a).to(dev) for a in arrs]
                g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in self._grid_t]
                self.grid = g
        # grid over the nodes themselves for large trees (config 4): same results as the scans, far fewer nodes touched
        self.node_grid = NodeGrid()
        if node_grid:
            if m == MODE_FN:
                raise ValueError("the node grid does not support forced removal (RRT*-FN)")
            pts = np.concatenate([starts, goals, np.asarray(xy_range, dtype=np.float64).T])
            x0, x1, y0, y1 = pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max()
            cell = float(node_grid_cell) if node_grid_cell > 0 else max(float(radius or step_length) / 2.0, max(x1 - x0, y1 - y0) / 1024.0)
            g = self.node_grid
            g.x0, g.y0, g.cell, g.inv_cell = x0, y0, cell, 1.0 / cell
            g.nx, g.ny = max(1, int(np.ceil((x1 - x0) / cell))), max(1, int(np.ceil((y1 - y0) / cell)))
            g.min_nodes, g.cand_cap = int(node_grid_min_nodes), int(node_grid_cand_cap)
            self._ng_t = [torch.empty((P, g.nx * g.ny), dtype=torch.int32, device=dev),
                          torch.empty((P, Cn), dtype=torch.int32, device=dev),
                          torch.empty((P, 2 * g.cand_cap), dtype=torch.int32, device=dev)]
            g.head, g.next, g.cand = [t.data_ptr() for t in self._ng_t]
            prm.flags |= FLAG_NODE_GRID
        self.trace_t = torch.zeros((P, max(prm.trace_cap, 1) * TRACE_BYTES), dtype=torch.uint8, device=dev) if trace else None
        self._initialised = False
        self._loaded = False     # load_tree() may bring coordinates from outside the box of xy_range / starts / goals
        smem = _abi.lib.rrtfnd_plan_smem_bytes(C.byref(prm))
        if smem < 0:
            raise _abi.LibraryError(f"planner does not fit in shared memory: {_abi.status_string(smem)}")
        self.smem_bytes = int(smem)

    def _flag_rectangles(self):
        """RRTFND_FLAG_RECTANGLES follows the table: the batch kernels of circle-only maps are compiled without the dispatch."""
        has = self.obst_host is not None and len(self.obst_host) > 0 and bool((self.obst_host[:, 2] < 0).any())
        self.prm.flags = (self.prm.flags | FLAG_RECTANGLES) if has else (self.prm.flags & ~FLAG_RECTANGLES)

    # ------------------------------------------------------------------ C-ABI plumbing
    def _batch(self) -> Batch:
        b = Batch()
        b.n_planners = self.P
        b.planners = self.planners_t.data_ptr()
        b.xy, b.ecost = self.xy.data_ptr(), self.ecost.data_ptr()
        b.parent, b.nchild, b.id = self.parent.data_ptr(), self.nchild.data_ptr(), self.id.data_ptr()
        b.nflags = self.nflags.data_ptr()
        b.grid = self.grid
        b.node_grid = self.node_grid
        b.finish, b.best_path = self.finish.data_ptr(), self.best_path_t.data_ptr()
        b.obst = self.obst.data_ptr()
        b.words = self.words.data_ptr() if self.words is not None else None
        b.trace = self.trace_t.data_ptr() if self.trace_t is not None else None
        b.walk = self.walk.data_ptr() if self.walk is not None else None
        return b

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        """Graph.__init__ for every planner (src/graph.py:25-32): a single root node at `start`."""
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_init_batch(C.byref(self.prm), C.byref(b), self._stream()), "rrtfnd_init_batch")
        self._initialised = True

    def load_tree(self, p: int, ids, xy, parent_ids, cost, root_id: int, last_id: int):
        """Replace planner p's tree by an existing one (a Graph that already holds nodes): rows in ascending id,
        parents given as ids (-1 = none). Planner-local bookkeeping (finish list, best path) starts empty, as it does
        at the top of the reference's planner functions (src/algorithm.py:113-116)."""
        if self.prm.flags & FLAG_NODE_GRID:
            raise ValueError("load_tree is not available with the node grid (it is filled by inserts only)")
        if not self._initialised:
            self.reset()
        self._loaded = True
        ids = np.asarray(ids, dtype=np.int32)
        n = len(ids)
        if n > self.prm.capacity - 1:
            raise ValueError(f"tree of {n} nodes does not fit the capacity {self.prm.capacity}")
        assert np.all(np.diff(ids) > 0), "ids must be strictly ascending"
        slot_of = {int(i): s for s, i in enumerate(ids.tolist())}
        par = np.asarray(parent_ids, dtype=np.int64)
        pslot = np.array([slot_of[int(q)] if q >= 0 else -1 for q in par], dtype=np.int32)
        nchild = np.bincount(pslot[pslot >= 0], minlength=n).astype(np.int32)
        dev = self.device
        self.xy[p, :n] = torch.from_numpy(np.ascontiguousarray(xy, dtype=np.float64).reshape(n, 2)).to(dev)
        self.ecost[p, :n] = torch.from_numpy(np.ascontiguousarray(cost, dtype=np.float64)).to(dev)
        self.parent[p, :n] = torch.from_numpy(pslot).to(dev)
        self.nchild[p, :n] = torch.from_numpy(nchild).to(dev)
        self.id[p, :n] = torch.from_numpy(ids).to(dev)
        self.nflags[p, :n] = 0
        pl = self.planners_t[p].cpu().numpy().view(PLANNER_DTYPE).copy()
        pl["n_slots"], pl["n_live"], pl["last_id"], pl["root_slot"] = n, n, int(last_id), slot_of[int(root_id)]
        self.planners_t[p] = torch.from_numpy(pl.view(np.uint8)).to(dev)

    def grow_capacity(self, capacity: int):
        """Re-allocate the per-node arrays with `capacity` slots per planner, keeping every tree (slot numbers are
        planner-local, so nothing else changes). The replanning loop uses it to make room for regrow() after an RRT*-FN
        run whose own capacity is kept tight (its scans cover the slots in use, tombstones included)."""
        old, capacity = int(self.prm.capacity), int(capacity)
        if capacity <= old:
            return
        if self.prm.flags & FLAG_NODE_GRID:
            raise ValueError("grow_capacity is not available with the node grid")
        for name in ("xy", "ecost", "parent", "nchild", "id", "nflags"):
            t = getattr(self, name)
            new = torch.zeros((self.P, capacity) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
            new[:, :old] = t
            setattr(self, name, new)
        self.prm.capacity = capacity

    def restore_problems(self):
        """Planner records back to the problems given at construction (select_branch re-roots a planner at the robot's
        position, i.e. overwrites its `start`); reset() follows."""
        self.planners_t.copy_(self._problems_t)
        self._initialised = False

    def reallocate(self, capacity: int):
        """Fresh per-node arrays with `capacity` slots per planner; every tree is dropped (reset() follows)."""
        if self.prm.flags & FLAG_NODE_GRID:
            raise ValueError("reallocate is not available with the node grid")
        for name in ("xy", "ecost", "parent", "nchild", "id", "nflags"):
            t = getattr(self, name)
            setattr(self, name, torch.zeros((self.P, int(capacity)) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device))
        self.prm.capacity = int(capacity)
        self._initialised = False

    # ------------------------------------------------------------------ FND tree surgery (src/algorithm.py:294-434)
    def _ids_tensor(self, ids):
        a = np.asarray(ids, dtype=np.int32).reshape(-1)
        assert len(a) == self.P, "one node id per planner"
        return torch.from_numpy(a.copy()).to(self.device)

    def set_obstacles(self, obstacles, grid=True, grid_cell=0.0, _table=None, rectangles=()):
        """Replace the obstacle table (a moved / extended map) and rebuild the culling grid.

        The grid's box must bound every node: trees grown by the planners stay inside the box of xy_range, starts and
        goals (a steered point lies between a node and a sample), so that box is reused; only after load_tree() are the
        node coordinates themselves read back to bound them."""
        self.obst_host = obstacle_table(obstacles, rectangles) if _table is None else _table     # _table: rows that already carry K
        self.prm.n_obst = len(self.obst_host)
        self._flag_rectangles()
        self.obst = torch.from_numpy(self.obst_host if len(self.obst_host) else np.zeros((1, 4))).to(self.device)
        self.grid, self._grid_t = ObstGrid(), None
        if grid and len(self.obst_host):
            bounds = getattr(self, "_bounds", None)
            if bounds is None or self._loaded:
                pl = self.planners()
                pts = np.concatenate([np.column_stack([pl["start_x"], pl["start_y"]]), np.column_stack([pl["goal_x"], pl["goal_y"]]),
                                      np.array([[self.prm.x_lo, self.prm.y_lo], [self.prm.x_hi, self.prm.y_hi]])])
                n = pl["n_slots"]
                xy = self.xy.cpu().numpy()
                live = [xy[p, :int(n[p])] for p in range(self.P)]
                allxy = np.concatenate(live + [pts])
                allxy = allxy[~np.isnan(allxy).any(axis=1)]
                bounds = (allxy[:, 0].min(), allxy[:, 0].max(), allxy[:, 1].min(), allxy[:, 1].max())
            g, arrs = build_obstacle_grid(self.obst_host, self.prm.node_radius, bounds, grid_cell)
            if arrs is not None:
                self._grid_t = [torch.from_numpy(a).to(self.device) for a in arrs]
                g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in self._grid_t]
                self.grid = g

    def set_obstacle_table_async(self, table4_host: np.ndarray, grid_cell=0.0):
        """set_obstacles for the replanning loop: the rows (ox, oy, r, K) go up with one asynchronous copy and the culling grid
        is rebuilt by a kernel; the host neither builds lists nor waits."""
        t = np.ascontiguousarray(table4_host, dtype=np.float64).reshape(-1, 4)
        b = self._bounds
        coord_max = float(max(np.abs(t[:, :3]).max(), np.abs(t[t[:, 2] < 0, 3]).max() if (t[:, 2] < 0).any() else 0.0,
                              abs(b[0]), abs(b[1]), abs(b[2]), abs(b[3])))
        self.set_obstacles_device(torch.from_numpy(t).to(self.device, non_blocking=True), coord_max, grid_cell)
        self.obst_host = t
        self._flag_rectangles()

    def set_obstacles_device(self, table4_dev, coord_max, grid_cell=0.0):
        """Replace the obstacle table by `table4_dev` (a float64 CUDA tensor [M][4] holding ox, oy, r, K rows, same M as
        now) and rebuild the culling grid with a kernel (rrtfnd_obst_grid_build_device): nothing crosses the host, which
        is what the replanning loop wants when the obstacles move every tick. `coord_max` bounds every |coordinate| and
        radius the table can hold (it sets the grid's exactness margin). Only for trees grown by the planners (the box of
        xy_range, starts and goals bounds them)."""
        assert table4_dev.is_cuda and table4_dev.dtype == torch.float64 and tuple(table4_dev.shape) == (self.prm.n_obst, 4)
        assert not self._loaded, "after load_tree() use set_obstacles(), which bounds the loaded nodes"
        self.obst = table4_dev.contiguous()
        self.obst_host = None                                   # stale from now on
        M = self.prm.n_obst
        if getattr(self, "_dev_grid", None) is None:
            b = self._bounds
            cells_cap, items_cap = 512 * 512 + 1, max(4096, 64 * M)
            self._dev_grid = [torch.zeros(cells_cap, dtype=torch.int32, device=self.device), torch.zeros(items_cap, dtype=torch.int32, device=self.device),
                              torch.zeros(cells_cap, dtype=torch.int32, device=self.device), torch.zeros(items_cap, dtype=torch.int32, device=self.device)]
            self._grid_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        g = ObstGrid()
        g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in self._dev_grid]
        b = self._bounds
        with torch.cuda.device(self.device):
            _abi.check(_abi.lib.rrtfnd_obst_grid_build_device(self.obst.data_ptr(), M, float(self.prm.node_radius), float(b[0]), float(b[1]),
                                                              float(b[2]), float(b[3]), float(grid_cell), float(coord_max), C.byref(g),
                                                              self._dev_grid[1].numel(), self._dev_grid[3].numel(), self._dev_grid[0].numel(),
                                                              self._grid_overflow.data_ptr(), self._stream()),
                       "rrtfnd_obst_grid_build_device")
        self.grid, self._grid_t = g, self._dev_grid

    def rebuild_grid(self, grid_cell=0.0):
        """Rebuild the culling grid over a box that bounds the nodes the planners hold NOW (after load_tree() a caller's
        vertices may lie outside xy_range; the grid's exactness argument needs every segment endpoint inside its box)."""
        self.set_obstacles(None, grid=True, grid_cell=grid_cell, _table=self.obst_host)

    def select_branch(self, current_ids):
        """select_branch for every planner: re-root at node current_ids[p] of its best path. Returns statuses."""
        cur = self._ids_tensor(current_ids)
        out = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_select_branch(C.byref(self.prm), C.byref(b), C.c_void_p(cur.data_ptr()),
                                                         C.c_void_p(out.data_ptr()), self._stream()), "rrtfnd_fnd_select_branch")
        return out.cpu().numpy()

    def valid_path(self, previous_root_ids):
        """valid_path for every planner against the current obstacle table. Returns (separate root ids, index_error flags)."""
        prev = self._ids_tensor(previous_root_ids)
        sep = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        err = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_valid_path(C.byref(self.prm), C.byref(b), C.c_void_p(prev.data_ptr()),
                                                      C.c_void_p(sep.data_ptr()), C.c_void_p(err.data_ptr()), self._stream()),
                       "rrtfnd_fnd_valid_path")
        return sep.cpu().numpy(), err.cpu().numpy()

    def cut_blocked_edges(self, separate_root_ids):
        """EXTENSION (edge invalidation, include/rrtfnd.h): after valid_path, cut the path edges that cross an obstacle.
        Returns (separate root ids after the cuts, edges cut)."""
        sep = self._ids_tensor(separate_root_ids)
        out = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        cuts = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_cut_blocked_edges(C.byref(self.prm), C.byref(b), C.c_void_p(sep.data_ptr()),
                                                             C.c_void_p(out.data_ptr()), C.c_void_p(cuts.data_ptr()), self._stream()),
                       "rrtfnd_fnd_cut_blocked_edges")
        return out.cpu().numpy(), cuts.cpu().numpy()

    def reconnect(self, separate_root_ids, radius):
        """reconnect for every planner. Returns the id each separate root was linked to (-1 = none)."""
        sep = self._ids_tensor(separate_root_ids)
        out = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_reconnect(C.byref(self.prm), C.byref(b), C.c_void_p(sep.data_ptr()), float(radius),
                                                     C.c_void_p(out.data_ptr()), self._stream()), "rrtfnd_fnd_reconnect")
        return out.cpu().numpy()

    def regrow(self, separate_root_ids, max_iter=500):
        """regrow for every planner (draws from each planner's word stream at its cursor). Returns (reconnected or negative
        status, extensions made)."""
        sep = self._ids_tensor(separate_root_ids)
        out = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        its = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_regrow(C.byref(self.prm), C.byref(b), C.c_void_p(sep.data_ptr()), int(max_iter),
                                                  C.c_void_p(out.data_ptr()), C.c_void_p(its.data_ptr()), self._stream()),
                       "rrtfnd_fnd_regrow")
        return out.cpu().numpy(), its.cpu().numpy()

    def set_stream_words(self, words, cursor=0):
        """Switch the planners to an explicit word buffer ([P][n] uint32) read from `cursor` on."""
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(self.P, -1)
        self.words = torch.from_numpy(w.view(np.int32)).to(self.device)
        self.prm.stream_mode, self.prm.words_per_planner = STREAM_WORDS, w.shape[1]
        pl = self.planners_t.cpu().numpy().reshape(-1).view(PLANNER_DTYPE).copy()
        pl["cursor"] = cursor
        self.planners_t = torch.from_numpy(pl.view(np.uint8).reshape(self.P, PLANNER_BYTES)).to(self.device)

    def slot_id(self, p: int, slot: int) -> int:
        return int(self.id[p, slot].item())

    def set_best_path(self, p: int, ids):
        """Store `ids` (goal side first) as planner p's path: the `path` argument of the FND functions."""
        a = np.asarray(ids, dtype=np.int32)
        assert len(a) <= self.prm.path_cap
        self.best_path_t[p, :len(a)] = torch.from_numpy(a).to(self.device)
        pl = self.planners_t[p].cpu().numpy().view(PLANNER_DTYPE).copy()
        pl["best_len"] = len(a)
        self.planners_t[p] = torch.from_numpy(pl.view(np.uint8)).to(self.device)

    def run(self):
        """Grow every tree until iter == iter_num (asynchronous on the current stream)."""
        if not self._initialised:
            self.reset()
        with torch.cuda.device(self.device):
            b = self._batch()
            _abi.check(_abi.lib.rrtfnd_plan_batch(C.byref(self.prm), C.byref(b), self._stream()), "rrtfnd_plan_batch")
        return self

    # ------------------------------------------------------------------ results
    def planners(self) -> np.ndarray:
        return self.planners_t.cpu().numpy().reshape(-1).view(PLANNER_DTYPE).copy()

    def check_status(self):
        st = self.planners()["status"]
        bad = np.flatnonzero((st != ST_DONE))
        if bad.size:
            p = int(bad[0])
            raise RuntimeError(f"planner {p}: status {int(st[p])} ({STATUS_NAMES.get(int(st[p]), '?')}); {bad.size} planners not done")

    def tree(self, p: int) -> dict:
        """Canonical tree of planner p: live nodes in ascending id, parent as id (-1 = none)."""
        n = int(self.planners()[p]["n_slots"])
        ids = self.id[p, :n].cpu().numpy()
        live = ids >= 0
        par_slot = self.parent[p, :n].cpu().numpy()
        parent = np.where(par_slot >= 0, ids[np.clip(par_slot, 0, None)], -1).astype(np.int32)
        xy = self.xy[p, :n].cpu().numpy()
        return dict(ids=ids[live].copy(), x=xy[live, 0].copy(), y=xy[live, 1].copy(),
                    parent=parent[live], cost=self.ecost[p, :n].cpu().numpy()[live],
                    nchild=self.nchild[p, :n].cpu().numpy()[live])

    def graph(self, p: int, xy_range=None):
        """Planner p as a `Graph` (src/graph.py:8-98), built on demand from the device arrays: ids, positions, parent /
        children / cost dictionaries, `last_id`, `start` = the current root. Batches stay on the device; only the
        planners a caller looks at are turned into dictionaries."""
        from .graph import Graph
        pl = self.planners()[p]
        rng = xy_range if xy_range is not None else [[self.prm.x_lo, self.prm.x_hi], [self.prm.y_lo, self.prm.y_hi]]
        G = Graph((np.float64(pl["start_x"]), np.float64(pl["start_y"])), (np.float64(pl["goal_x"]), np.float64(pl["goal_y"])),
                  0, 0, rng)
        t = self.tree(p)
        G._import(t["ids"], np.column_stack([t["x"], t["y"]]), t["parent"], t["cost"], int(pl["last_id"]))
        G.start = G.vertices[self.slot_id(p, int(pl["root_slot"]))] if self.slot_id(p, int(pl["root_slot"])) in G.vertices else G.start
        return G

    def best_path(self, p: int) -> np.ndarray:
        n = int(self.planners()[p]["best_len"])
        return self.best_path_t[p, :n].cpu().numpy().copy()

    def trace(self, p: int) -> np.ndarray:
        return self.trace_t[p].cpu().numpy().view(TRACE_DTYPE).copy()
