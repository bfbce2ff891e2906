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

    def __init__(self, batch, reconnect_radius, regrow_iters=500):
        self.b = batch
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
