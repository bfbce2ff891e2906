"""Obstacle ingestion (src/sdf_reader.py:7-33): the drop-in SDFReader on a synthetic 9-cylinder world, against the
reference's own reader where /root/reference is present and against known values everywhere."""
import os
import sys

import numpy as np
import pytest

from rrt_star_fnd_algorithm_b200.sdf_reader import SDFReader
from util import GOLDEN

WORLD = os.path.join(GOLDEN, "world9.sdf")


def read(cls):
    r = cls()
    assert r.get_obstacles() is None
    r.parse(WORLD)
    return r.get_obstacles()


def test_reader_returns_the_reference_layout():
    obstacles, (xr, yr) = read(SDFReader)
    assert obstacles.dtype == object and obstacles.shape == (9, 2)
    centres = [(a, b) for a in (-1.1, 0.0, 1.1) for b in (-1.1, 0.0, 1.1)]
    for row, c in zip(obstacles, centres):
        assert isinstance(row[0], np.ndarray) and tuple(row[0]) == c and row[1] == 0.15
    assert xr == [-1.1 * 100 - 100, 1.1 * 100 + 100] and yr == xr
    scaled = obstacles * 100                                    # what src/main.py:92 hands to Map.add_obstacles
    assert tuple(scaled[0][0]) == (-1.1 * 100, -1.1 * 100) and scaled[0][1] == 15.0


def test_scaled_obstacles_feed_the_device_table():
    from rrt_star_fnd_algorithm_b200.batch import obstacle_table
    obstacles, _ = read(SDFReader)
    tab = obstacle_table(list(obstacles * 100))
    assert tab.shape == (9, 4)
    x0, y0, r = -1.1 * 100, -1.1 * 100, 0.15 * 100
    assert tab[0, 3] == -r ** 2 + x0 ** 2 + y0 ** 2              # K with the reference's expression (src/algorithm.py:619)


@pytest.mark.reference
def test_reader_equals_the_reference_reader():
    sys.path.insert(0, "/root/reference/src")
    try:
        from sdf_reader import SDFReader as Ref
    finally:
        sys.path.pop(0)
    got, (gx, gy) = read(SDFReader)
    want, (wx, wy) = read(Ref)
    assert got.shape == want.shape
    for a, b in zip(got, want):
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and type(a[1]) is type(b[1])
    assert list(gx) == list(wx) and list(gy) == list(wy)
