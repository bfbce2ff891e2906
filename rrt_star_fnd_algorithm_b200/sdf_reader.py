"""SDFReader — obstacle ingestion from a Gazebo SDF world (src/sdf_reader.py:7-33), host side of the obstacle table.

Same surface as the reference: `parse(filepath)` then `get_obstacles()` -> `(obstacles, (x_range, y_range))` where
`obstacles` is an object ndarray of rows `[ndarray([x, y]), radius]` in metres (callers scale it, `obstacles * 100` in
src/main.py:92) and the ranges are the cylinder centres in centimetres widened by 100 on each side (:28-31). Cylinders
are the `<visual>` elements whose `<geometry>` holds a `<cylinder>`; the centre is the first two numbers of the visual's
`<pose>`. Elements are looked up by tag here, not by child position as the reference does, so a well-formed file gives
identical output and a reordered one still parses.
"""
import xml.etree.ElementTree as ET

import numpy as np


class SDFReader:
    def __init__(self):
        self.tree = None
        self.root = None

    def parse(self, filepath: str) -> None:
        if filepath is None:
            return None
        self.tree = ET.parse(filepath)
        self.root = self.tree.getroot()

    def _cylinders(self):
        """(pose numbers, radius) per cylinder visual, in document order."""
        found = []
        for visual in self.root.iter("visual"):
            radius = visual.find("./geometry/cylinder/radius")
            if radius is None:
                # any direct child holding a <cylinder> counts, as in the reference's `visual/*[cylinder]` query
                holder = next((c for c in visual if c.find("cylinder") is not None), None)
                radius = None if holder is None else holder.find("cylinder/radius")
            pose = visual.find("pose")
            if radius is None or pose is None:
                continue
            found.append(([float(tok) for tok in pose.text.split(" ")], float(radius.text)))
        return found

    def get_obstacles(self):
        if self.tree is None:
            return None
        cyl = self._cylinders()
        obstacles = np.empty((len(cyl), 2), dtype=object)
        for i, (pose, radius) in enumerate(cyl):
            obstacles[i, 0] = np.array([pose[0], pose[1]])
            obstacles[i, 1] = radius
        xs = np.array([pose[0] for pose, _ in cyl], dtype=object) * 100
        ys = np.array([pose[1] for pose, _ in cyl], dtype=object) * 100
        x_range = [np.min(xs) - 100, np.max(xs) + 100]
        y_range = [np.min(ys) - 100, np.max(ys) + 100]
        return obstacles, (x_range, y_range)
