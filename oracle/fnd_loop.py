"""CPU restatement of the RRT*-FND replanning loop (rrt_star_fnd_algorithm_b200/fnd_loop.py) — TEST INFRASTRUCTURE ONLY.

The steps are the oracle's restatements of the reference's select_branch / valid_path / reconnect / regrow
(src/algorithm.py:294-552), each pinned on goldens from the unmodified reference; their composition into a tick has no
counterpart in the reference (RRT_STAR_FND stops at the init_movement() stub, :283-291), so for the loop itself parity
is "device composition == oracle composition", not "== reference"."""
import numpy as np

MOVING, ARRIVED, LOST = 0, 1, 2
MAX_ROUNDS = 4


class OracleReplanner:
    """One planner (an oracle.OracleTree that already ran its planner) stepped through replanning ticks."""

    def __init__(self, tree, prm, goal, seed, stream_id, node_radius, reconnect_radius, regrow_iters=500, edge_invalidation=False):
        self.edge_invalidation = bool(edge_invalidation)
        self.t, self.prm, self.goal, self.seed, self.stream_id = tree, prm, goal, int(seed), int(stream_id)
        self.node_radius, self.reconnect_radius, self.regrow_iters = float(node_radius), float(reconnect_radius), int(regrow_iters)
        n = len(tree.best_path())
        self.phase = LOST if n == 0 else (ARRIVED if n == 1 else MOVING)

    def tick(self, obstacles):
        """Returns dict(current, status, separate_root, index_error, phase, linked[R], regrow[R], regrow_iters[R])."""
        t = self.t
        out = dict(current=-1, status=-1, separate_root=-1, index_error=0, linked=[-1] * MAX_ROUNDS, regrow=[0] * MAX_ROUNDS,
                   regrow_iters=[0] * MAX_ROUNDS, edge_cuts=0)
        path = t.best_path()
        if self.phase == MOVING and len(path) >= 2:
            cur = int(path[-2])                                   # the robot steps to the next node of its path
            out["current"] = cur
            st = t.select_branch(cur)
            out["status"] = st
            bad, sep = st != 0, cur
            if not bad:
                sep, err = t.valid_path(obstacles, self.node_radius, cur)
                if self.edge_invalidation and not err:            # extension: also cut path edges that now cross an obstacle
                    sep, out["edge_cuts"] = t.cut_blocked_edges(obstacles, sep)
                out["separate_root"], out["index_error"] = (-1 if err else sep), int(err)
                bad = err or not self._alive(t.root())
            if bad:
                self.phase = LOST
            todo = sep if (not bad and sep != cur) else -1
            for rnd in range(MAX_ROUNDS):                         # one separate tree per round, goal side first
                if todo < 0:
                    break
                out["linked"][rnd] = t.reconnect(obstacles, todo, self.reconnect_radius, int(t.best_path()[0]))
                if out["linked"][rnd] < 0:
                    res, its = t.regrow(self.prm, obstacles, self.goal, todo, seed=self.seed, stream_id=self.stream_id,
                                        max_iter=self.regrow_iters)
                    out["regrow"][rnd], out["regrow_iters"][rnd] = res, its
                    if res != 1:
                        self.phase = LOST
                        todo = -1
                        break
                tail = int(t.best_path()[-1])
                todo = tail if tail != t.root() else -1
            if todo >= 0:
                self.phase = LOST                                 # still broken after MAX_ROUNDS
            if self.phase == MOVING and len(t.best_path()) <= 1:
                self.phase = ARRIVED
        out["phase"] = self.phase
        return out

    def _alive(self, node_id):
        return int(node_id) in set(self.t.export()["ids"].tolist())
