"""Map — obstacle container of the reference (src/map.py:6-72) in front of the device obstacle table.

Circular obstacles are `[(x, y), r]` entries of `obstacles_c`, exactly as the reference stores them (tuples from
generate_obstacles, object-ndarray rows from the SDF reader). Point occupancy is evaluated by the CUDA operator
`rrtfnd_points_occupied` (there is no CPU fallback); the planners read `obstacles_c` and `node_radius` directly.
"""
import ctypes as C
import random

import numpy as np


class Map:
    def __init__(self, size: tuple, start: tuple, goal: tuple, node_radius: int):
        self.obstacles_c = []   # circles: [(x, y), r]
        # rectangles [(cx, cy), (hx, hy)] (centre, half extents): the list is declared by the reference and never filled or
        # tested there (src/map.py:18, :54); here it is an EXTENSION with builder-defined semantics (include/rrtfnd.h)
        self.obstacles_r = []
        self.width = size[0]
        self.height = size[1]
        self.start = start
        self.goal = goal
        self.node_radius = node_radius

    def add_obstacles(self, obstacles):
        for obstacle in obstacles:
            self.obstacles_c.append(obstacle)

    def add_rectangles(self, rectangles):
        """Extension: axis-aligned rectangles `[(cx, cy), (hx, hy)]`, tested like the circles (segments against the raw
        shape, points against the shape inflated by node_radius)."""
        for rect in rectangles:
            self.obstacles_r.append(rect)

    def obstacle_table(self) -> np.ndarray:
        """[M][3] float64 (x, y, r) view of obstacles_c."""
        return np.array([[float(o[0][0]), float(o[0][1]), float(o[1])] for o in self.obstacles_c], dtype=np.float64).reshape(-1, 3)

    # ---- Map.is_occupied_c (src/map.py:38-47): strict `<` against r + node_radius, on the GPU ---------------------
    def occupied(self, points) -> np.ndarray:
        """Occupancy of many points at once: bool array, one per row of `points` [n][2]."""
        import torch
        from . import _abi
        from .batch import obstacle_table
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        if len(pts) == 0 or not (self.obstacles_c or self.obstacles_r):
            return np.zeros(len(pts), dtype=bool)
        dev = torch.device("cuda", torch.cuda.current_device())
        tab = torch.from_numpy(obstacle_table(self.obstacles_c, self.obstacles_r)).to(dev)
        px, py = torch.from_numpy(pts[:, 0].copy()).to(dev), torch.from_numpy(pts[:, 1].copy()).to(dev)
        out = torch.zeros(len(pts), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _abi.check(_abi.lib.rrtfnd_points_occupied(C.c_void_p(px.data_ptr()), C.c_void_p(py.data_ptr()), len(pts),
                                                   C.c_void_p(tab.data_ptr()), len(tab),
                                                   float(self.node_radius), C.c_void_p(out.data_ptr()), st),
                   "rrtfnd_points_occupied")
        return out.cpu().numpy().astype(bool)

    def is_occupied_c(self, p: tuple) -> bool:
        return bool(self.occupied([p])[0])

    def collides_w_start_goal_c(self, obs) -> bool:
        """Would the circle `obs = ((x, y), r)` cover the start or the goal node? (src/map.py:27-36)"""
        probe = Map((self.width, self.height), self.start, self.goal, self.node_radius)
        probe.obstacles_c = [obs]
        return bool(probe.occupied([self.start, self.goal]).any())

    def generate_obstacles(self, obstacle_count: int = 10, size: int = 1):
        """Random integer-centred circles of radius `size` clear of start and goal, drawn from the global `random`
        in the reference's order (src/map.py:53-64): a rejected circle is redrawn before the next one."""
        made = 0
        while made < obstacle_count:
            centre = (random.randint(0 + size, self.width - size), random.randint(0 + size, self.height - size))
            if self.collides_w_start_goal_c((centre, size)):
                continue
            self.obstacles_c.append((centre, size))
            made += 1

    def show(self):
        """Plot the map (needs matplotlib, like the reference; visualisation is outside the GPU path)."""
        import matplotlib.pyplot as plt
        plt.gca().set_aspect('equal', adjustable='box')
        for obstacle in self.obstacles_c:
            plt.gca().add_patch(plt.Circle(obstacle[0], obstacle[1], color='black'))
        plt.show()
