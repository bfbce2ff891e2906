"""Line — the segment object the reference hands to its collision helpers (src/line.py:4-32), kept so that user code
calling `through_obstacle(Line(p1, p2), obstacles)` or `intersection_circle(Line(p1, p2), circle)` works unchanged.

Only the endpoints matter to the GPU path: the device evaluates slope / intercept itself with the reference's
roundings (csrc/arith.cuh `make_seg`; which endpoint is p1 decides the intercept, src/line.py:16). The slope-form
attributes are still provided, computed with the reference's expressions, for callers that read them.
"""
import numpy as np


class Line:
    __slots__ = ("p1", "p2", "norm", "dir", "x_pos", "const_term")

    def __init__(self, p1: tuple, p2: tuple):
        a, b = np.array(p1), np.array(p2)
        self.p1, self.p2 = a, b
        self.norm = np.linalg.norm(b - a)
        vertical = p2[0] == p1[0]                      # exact comparison of the caller's own numbers (src/line.py:9)
        self.x_pos = p1[0] if vertical else None
        self.dir = float("inf") if vertical else (p2[1] - p1[1]) / (p2[0] - p1[0])
        self.const_term = None if vertical else a[1] - self.dir * a[0]

    def calculate_y(self, x0: float) -> float:
        """y of the line at x0 (src/line.py:18-24)."""
        return self.dir * x0 + self.const_term

    def calculate_x(self, y0: float) -> float:
        """x of the line at y0 (src/line.py:26-32)."""
        return (y0 - self.const_term) / self.dir
