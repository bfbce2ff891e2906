"""Graph — the tree object of the reference (src/graph.py:8-98), kept as the API surface of the GPU planners.

The planners in `algorithm.py` upload a Graph to the device (structure-of-arrays, one slot per node in ascending id
order), grow it there and write the result back into these same dictionaries, so code written against the
reference — `G.vertices`, `G.parent`, `G.children`, `G.cost`, `G.id_vertex`, `G.last_id`, `G.get_cost(...)`,
plotting helpers that walk the dictionaries — keeps working unchanged.
"""
import random

import numpy as np


class OutOfBoundsException(Exception):
    """Declared by the reference (src/graph.py:4) but never raised there: its bounds checks are commented out."""


class Graph:
    def __init__(self, start: tuple, goal: tuple, width: int, height: int, xy_range: list):
        # like the reference, `width` / `height` arguments are ignored: the extent is that of xy_range (src/graph.py:22-23)
        self.start = start
        self.goal = goal
        self.width = xy_range[0][1] - xy_range[0][0]
        self.height = xy_range[1][1] - xy_range[1][0]
        self.xy_range = xy_range
        self.parent = {0: None}
        self.children = {0: []}
        self.vertices = {0: self.start}
        self.cost = {0: 0.0}            # edge cost to the node's parent
        self.id_vertex = {self.start: 0}
        self.last_id = 0

    # ---- bookkeeping with the reference's semantics (src/graph.py:34-87) -------------------------------------
    def get_cost(self, from_node: int) -> float:
        """Cost to the root: edge costs summed leaf -> root, one fp64 addition per edge (src/graph.py:41-46)."""
        total = np.float64(0.0)
        node = from_node
        first = True
        while self.parent[node] is not None:
            # sum() starts from int 0; 0 + x is exact, so starting from the first addend is the same value
            total = np.float64(self.cost[node]) if first else total + np.float64(self.cost[node])
            first = False
            node = self.parent[node]
        return total if not first else 0

    def add_vertex(self, pos: tuple) -> int:
        """Id of the vertex at `pos`, creating it if needed; ids grow monotonically and are never reused."""
        if pos in self.id_vertex:
            return self.id_vertex[pos]
        self.last_id += 1
        self.vertices[self.last_id] = pos
        self.id_vertex[pos] = self.last_id
        self.children[self.last_id] = []
        return self.last_id

    def remove_vertex(self, id: int):
        parent = self.parent[id]
        if parent is not None:
            self.children[parent].remove(id)
        pos = self.vertices[id]
        for table in (self.parent, self.children, self.vertices, self.cost):
            del table[id]
        del self.id_vertex[pos]

    def add_edge(self, id_node: int, id_parent: int, cost: float):
        self.parent[id_node] = id_parent
        self.children[id_parent].append(id_node)
        self.cost[id_node] = cost

    def random_node(self, bias: float = 0) -> tuple:
        """A sample drawn exactly as the reference draws it from the global `random` (src/graph.py:96-98)."""
        if random.random() < bias:
            return self.goal
        (x0, x1), (y0, y1) = self.xy_range
        return random.uniform(x0, x1), random.uniform(y0, y1)

    # ---- device round trip ---------------------------------------------------------------------------------------
    def _export(self):
        """Arrays for the device: ids ascending, positions, parent ids (-1 = none), edge costs, root id."""
        ids = np.array(sorted(self.vertices), dtype=np.int32)
        xy = np.array([[float(self.vertices[i][0]), float(self.vertices[i][1])] for i in ids], dtype=np.float64).reshape(-1, 2)
        parent = np.array([-1 if self.parent.get(int(i)) is None else self.parent[int(i)] for i in ids], dtype=np.int32)
        cost = np.array([float(self.cost.get(int(i), 0.0)) for i in ids], dtype=np.float64)
        return ids, xy, parent, cost, int(self.id_vertex[self.start])

    def _import(self, ids, xy, parent, cost, last_id, keep_positions=None):
        """Rebuild the dictionaries from device arrays (ids ascending; parent ids, -1 = none).

        Positions of vertices that existed before the run are kept as the caller's own objects (a start tuple of ints
        stays a tuple of ints, as in the reference); new ones are tuples of numpy float64 like steer() returns. Edge
        costs are numpy float64 — the reference's get_cost relies on that (sum() of np.float64 is a plain
        sequential sum, sum() of Python floats is compensated in CPython 3.12)."""
        keep = keep_positions or {}
        self.vertices, self.parent, self.children, self.cost, self.id_vertex = {}, {}, {}, {}, {}
        for i, (x, y), p, c in zip(ids.tolist(), xy, parent.tolist(), cost):
            pos = keep.get(i, (np.float64(x), np.float64(y)))
            self.vertices[i] = pos
            self.id_vertex[pos] = i
            self.parent[i] = None if p < 0 else p
            self.cost[i] = 0.0 if p < 0 and i == 0 else np.float64(c)
            self.children[i] = []
        # children in ascending id: the order the reference produces (creation order; nodes re-parented to a new
        # node were all created before it), up to cKDTree's internal order among nodes rewired in the same call
        for i, p in self.parent.items():
            if p is not None:
                self.children[p].append(i)
        self.last_id = int(last_id)
