"""RRT*-FND replanning loop for a batch of planners (BASELINE.json config 5): moving obstacles, path invalidation,
reconnect-or-regrow.

The reference stops at a stub (`RRT_STAR_FND` calls `init_movement()`, src/algorithm.py:283-291); what it ships is the
tree surgery (`select_branch`, `valid_path`, `reconnect`, `regrow`, :294-552) and one usage example of it
(`test_select_branch`, :437-479). This module is the loop that example implies — Algorithm 5 of the RRT*FND paper with
the reference's own functions as its steps — so its *steps* are pinned on reference goldens (tests/test_fnd.py) while
their *composition* is defined here and checked against the same composition of the CPU oracle's steps
(tests/test_fnd_loop.py). One tick, for every planner that is still moving:

    1. the robot advances to the next node of its path: current = path[-2]; select_branch(G, current, path, map) (:294)
    2. the obstacles move (the caller hands in the new table); id_separate_root = valid_path(G, path, map, current) (:344)
    3. if a part of the tree was split off: reconnect(G, path, map, reconnect_radius, root, id_separate_root, last) (:370)
    4. if that found no link: regrow(...) (:482) with at most `regrow_iters` extensions from the planner's own stream

Planner phases: MOVING; ARRIVED (the path has a single node left: the robot stands on the goal-side leaf); LOST (no
initial path, the goal-side leaf became occupied — the reference raises IndexError at :358 —, the robot's own node
became occupied, or regrow gave up). LOST planners are left alone; a caller would plan them afresh.

All device work goes through the C ABI (rrtfnd_fnd_*); the host logic here is integer bookkeeping on P-sized arrays.
"""
import numpy as np
import torch

from .workloads import moved_obstacles  # noqa: F401  (re-exported: the loop's callers take it from here)


MOVING, ARRIVED, LOST = 0, 1, 2


MAX_ROUNDS = 4   # separate trees mended per tick (one per occupied stretch of the path), goal side first


class FNDReplanner:
    """Drives the replanning ticks of a PlannerBatch whose planners hold a solution (run() it first)."""

    def __init__(self, batch, reconnect_radius, regrow_iters=500, edge_invalidation=False):
        self.b = batch
        self.edge_invalidation = bool(edge_invalidation)    # extension: also cut path EDGES that cross an obstacle
        self.reconnect_radius = float(reconnect_radius)
        self.regrow_iters = int(regrow_iters)
        pl = batch.planners()
        self.phase = np.where(pl["best_len"] >= 1, MOVING, LOST).astype(np.int32)
        self.phase[(self.phase == MOVING) & (pl["best_len"] == 1)] = ARRIVED
        self.ticks = 0
        self.stats = dict(moves=0, splits=0, reconnected=0, regrown=0, regrow_extensions=0, lost=0, arrived=0)

    def _path_ends(self):
        """(best_len, id of the last stored path node, id of the root node) per planner: two gathers on the device."""
        b = self.b
        pl = b.planners()
        best_len = pl["best_len"].astype(np.int64)
        col = torch.from_numpy(np.clip(best_len - 1, 0, None)).to(b.device)
        tail = b.best_path_t.gather(1, col[:, None]).squeeze(1).cpu().numpy()
        root = torch.from_numpy(pl["root_slot"].astype(np.int64)).to(b.device)
        root_id = b.id.gather(1, root[:, None]).squeeze(1).cpu().numpy()
        return best_len, np.where(best_len >= 1, tail, -1), root_id

    def tick(self, obstacles):
        """One replanning tick against the obstacle table `obstacles` [(ox, oy, r), ...]. Returns per-planner arrays:
        current, status, separate_root, index_error, phase [P] and linked, regrow, regrow_iters [P][MAX_ROUNDS]."""
        b, P = self.b, self.b.P
        best_len = b.planners()["best_len"].astype(np.int64)
        # path[-2] of every planner: one gather on the device, P ints to the host
        col = torch.from_numpy(np.clip(best_len - 2, 0, None)).to(b.device)
        nxt = b.best_path_t.gather(1, col[:, None]).squeeze(1).cpu().numpy()
        cur = np.where((self.phase == MOVING) & (best_len >= 2), nxt, -1).astype(np.int32)
        moving = cur >= 0
        st = b.select_branch(cur)
        b.set_obstacles(obstacles)
        sep, err = b.valid_path(np.where(moving & (st == 0), cur, -1).astype(np.int32))
        if self.edge_invalidation:
            sep, cuts = b.cut_blocked_edges(np.where(err != 0, -1, sep).astype(np.int32))
            self.stats["edge_cuts"] = self.stats.get("edge_cuts", 0) + int(cuts.sum())
        # is the robot's own node still there? (valid_path deletes an occupied root like any other path node)
        _, _, root_id = self._path_ends()
        bad = moving & ((st != 0) | (err != 0) | (root_id < 0))
        self.phase[bad] = LOST
        linked = np.full((P, MAX_ROUNDS), -1, np.int32)
        grown = np.zeros((P, MAX_ROUNDS), np.int32)
        grown_its = np.zeros((P, MAX_ROUNDS), np.int32)
        s = self.stats
        # several occupied stretches leave several separate trees; valid_path reports the goal-side one, and mending it
        # leaves a path that ends at the root of the next one: mend them one after the other, goal side first
        todo = np.where(moving & ~bad & (sep != cur), sep, -1).astype(np.int32)
        for rnd in range(MAX_ROUNDS):
            split = todo >= 0
            if not split.any():
                break
            linked[:, rnd] = b.reconnect(todo, self.reconnect_radius)
            need = split & (linked[:, rnd] < 0)
            if need.any():
                res, its = b.regrow(np.where(need, todo, -1).astype(np.int32), self.regrow_iters)
                grown[:, rnd], grown_its[:, rnd] = np.where(need, res, 0), np.where(need, its, 0)
            failed = need & (grown[:, rnd] != 1)
            self.phase[failed] = LOST
            s["splits"] += int(split.sum()); s["reconnected"] += int((split & ~need).sum())
            s["regrown"] += int((need & ~failed).sum()); s["regrow_extensions"] += int(grown_its[need, rnd].sum())
            _, tail, root_id = self._path_ends()
            todo = np.where(split & ~failed & (tail != root_id), tail, -1).astype(np.int32)
        self.phase[todo >= 0] = LOST                      # still broken after MAX_ROUNDS
        # a planner whose path shrank to its last node has arrived
        self.phase[(self.phase == MOVING) & moving & (b.planners()["best_len"] <= 1)] = ARRIVED
        self.ticks += 1
        s["moves"] += int(moving.sum())
        s["lost"], s["arrived"] = int((self.phase == LOST).sum()), int((self.phase == ARRIVED).sum())
        return dict(current=cur, status=st, separate_root=sep, index_error=err, linked=linked, regrow=grown,
                    regrow_iters=grown_its, phase=self.phase.copy())


class DeviceFNDReplanner:
    """The same loop with the bookkeeping on the device (rrtfnd_fnd_tick): per-planner phases and the tick's outcome live in
    device arrays, a tick is one C-ABI call that only launches kernels, the moved obstacle table goes up with one asynchronous
    copy and its culling grid is rebuilt by a kernel. Nothing is copied back unless the caller asks (`fetch=True`, what the
    parity tests do; `stats` downloads 16 counters and the phases). Results are identical to FNDReplanner's."""

    FIELDS = ("phase", "current", "status", "prev_root", "separate", "index_error", "edge_cuts", "todo", "regrow_arg",
              "link_tmp", "regrow_tmp", "iters_tmp")

    def __init__(self, batch, reconnect_radius, regrow_iters=500, edge_invalidation=False):
        import ctypes as C
        from ._structs import FndLoop
        self.b = batch
        self.reconnect_radius, self.regrow_iters = float(reconnect_radius), int(regrow_iters)
        self.edge_invalidation = bool(edge_invalidation)
        P, dev = batch.P, batch.device
        pl = batch.planners()                               # the one download: the phases at the start
        phase = np.where(pl["best_len"] >= 1, MOVING, LOST).astype(np.int32)
        phase[(phase == MOVING) & (pl["best_len"] == 1)] = ARRIVED
        self.t = {k: torch.zeros(P, dtype=torch.int32, device=dev) for k in self.FIELDS}
        self.t["phase"] = torch.from_numpy(phase).to(dev)
        for k in ("linked", "regrow", "regrow_iters"):
            self.t[k] = torch.zeros((P, MAX_ROUNDS), dtype=torch.int32, device=dev)
        self.stats_t = torch.zeros(16, dtype=torch.int64, device=dev)
        lp = FndLoop()
        lp.max_rounds = MAX_ROUNDS
        for k in self.FIELDS + ("linked", "regrow", "regrow_iters"):
            setattr(lp, k, self.t[k].data_ptr())
        lp.stats = self.stats_t.data_ptr()
        self._lp, self._C = lp, C
        self.ticks = 0

    def tick(self, obstacles, fetch=True):
        """One replanning tick against `obstacles` [(ox, oy, r), ...] (or table rows (ox, oy, r, K)). With fetch=True returns the
        same per-planner arrays as FNDReplanner.tick (one download); with fetch=False nothing leaves the device."""
        from . import _abi
        from .batch import obstacle_table
        b, C = self.b, self._C
        b.set_obstacle_table_async(obstacle_table(obstacles))
        with torch.cuda.device(b.device):
            bt = b._batch()
            _abi.check(_abi.lib.rrtfnd_fnd_tick(C.byref(b.prm), C.byref(bt), C.byref(self._lp), self.reconnect_radius,
                                                self.regrow_iters, int(self.edge_invalidation), b._stream()), "rrtfnd_fnd_tick")
        self.ticks += 1
        if not fetch:
            return None
        g = {k: self.t[k].cpu().numpy() for k in ("current", "status", "separate", "index_error", "linked", "regrow", "regrow_iters",
                                                  "phase", "edge_cuts")}
        return dict(current=g["current"], status=g["status"], separate_root=g["separate"], index_error=g["index_error"],
                    linked=g["linked"], regrow=g["regrow"], regrow_iters=g["regrow_iters"], phase=g["phase"], edge_cuts=g["edge_cuts"])

    @property
    def phase(self):
        return self.t["phase"].cpu().numpy()

    @property
    def stats(self):
        s = self.stats_t.cpu().numpy()
        if getattr(self.b, "_grid_overflow", None) is not None and int(self.b._grid_overflow.item()):
            raise RuntimeError("rrtfnd_obst_grid_build_device: a culling list did not fit its buffer during the ticks")
        ph = self.phase
        return dict(moves=int(s[0]), splits=int(s[1]), reconnected=int(s[2]), regrown=int(s[3]), regrow_extensions=int(s[4]),
                    lost=int((ph == LOST).sum()), arrived=int((ph == ARRIVED).sum()), edge_cuts=int(s[8]))
