"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d): maps, per-planner start/goal pairs and stream keys.

Everything is a pure function of (workload name, global planner index), so a planner's problem — and
therefore its tree — does not depend on how planners are sharded over GPUs.
"""
import numpy as np


def _circles(seed, n, lo, hi, rmin, rmax, keep_clear=(), node_radius=0.0):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c = rng.uniform(lo, hi, 2)
        r = rng.uniform(rmin, rmax)
        if any(np.hypot(c[0] - p[0], c[1] - p[1]) < node_radius + r for p in keep_clear):
            continue
        out.append((float(c[0]), float(c[1]), float(r)))
    return out


def _occupied(p, obst, node_radius):
    ob = np.asarray(obst)
    return bool((np.hypot(p[0] - ob[:, 0], p[1] - ob[:, 1]) < node_radius + ob[:, 2] + 1e-9).any())


def c1():
    """Config 1: RRT* on the reconstructed src/main.py map (9 cylinders), 2,000 iterations, one planner."""
    obst = [(float(x), float(y), 15.0) for x in (-110, 0, 110) for y in (-110, 0, 110)]
    return dict(name="c1", mode="star", obstacles=obst, xy_range=[[-210.0, 210.0], [-210.0, 210.0]], iter_num=2000,
                step_length=25.0, radius=50.0, node_radius=20.0, bias=0.01, max_nodes=0,
                start=(144.65717764052022, 108.34084923492702), goal=(-33.35993605104511, -101.2549648769554),
                seed_base=1, planners=1)


def c2():
    """Config 2: RRT*-FN, 1,000-node cap, 100x100 map, 200 circles, 4,096 planners sharing start/goal."""
    start, goal = (5.0, 5.0), (95.0, 95.0)
    obst = _circles(2, 200, 2, 98, 0.5, 2.0, keep_clear=(start, goal), node_radius=1.0)
    return dict(name="c2", mode="fn", obstacles=obst, xy_range=[[0.0, 100.0], [0.0, 100.0]], iter_num=5000,
                step_length=3.0, radius=6.0, node_radius=1.0, bias=0.01, max_nodes=1000, start=start, goal=goal,
                seed_base=1000, planners=4096)


def c2r():
    """Config 2 as BASELINE.json words it ("200 rectangular obstacles"): the same centres as c2, axis-aligned rectangles with
    half extents U(0.5, 2.0)^2 instead of circles. Rectangles are an EXTENSION - the reference declares Map.obstacles_r and
    never implements it (src/map.py:18, :54) - so this variant is parity-unpinned by the reference (pinned on the oracle)."""
    w = c2()
    rng = np.random.default_rng(22)
    rects = []
    for (cx, cy, _) in w["obstacles"]:
        while True:
            hx, hy = rng.uniform(0.5, 2.0, 2)
            clear = all(max(abs(p[0] - cx) - hx, 0.0) ** 2 + max(abs(p[1] - cy) - hy, 0.0) ** 2 >= (w["node_radius"] + 1e-9) ** 2
                        for p in (w["start"], w["goal"]))
            if clear:
                break
        rects.append((float(cx), float(cy), -float(hx), float(hy)))       # table rows (cx, cy, -hx, hy)
    return dict(w, name="c2r", obstacles=rects, seed_base=2000)


def c3():
    """Config 3: RRT*, 200x200 map, 500 circles, 16,384 planners with random start/goal, 5,000 iterations."""
    obst = _circles(3, 500, 0, 200, 1.0, 4.0)
    return dict(name="c3", mode="star", obstacles=obst, xy_range=[[0.0, 200.0], [0.0, 200.0]], iter_num=5000,
                step_length=5.0, radius=10.0, node_radius=2.0, bias=0.02, max_nodes=0, start=None, goal=None,
                seed_base=3000, planners=16384)


def c4():
    """Config 4: one RRT* tree grown to 1M nodes on a 1000x1000 map with 500 circles; near-set / nearest queries go
    through the device node grid (large-tree stress). Replicas (other streams) fill the GPU; the tree does not shard."""
    obst = _circles(4, 500, 0, 1000, 2.0, 10.0)
    return dict(name="c4", mode="star", obstacles=obst, xy_range=[[0.0, 1000.0], [0.0, 1000.0]], iter_num=1000000,
                step_length=2.0, radius=4.0, node_radius=1.0, bias=0.0, max_nodes=0, start=None, goal=None,
                seed_base=4000, planners=1, node_grid=True)


def moved_obstacles(base, velocity, tick):
    """Obstacle list (ox, oy, r) after `tick` ticks of straight-line motion: centre + tick * velocity (fp64, one multiply
    and one add per coordinate, the same expression for every consumer)."""
    b = np.asarray(base, dtype=np.float64).reshape(-1, 3)
    v = np.asarray(velocity, dtype=np.float64).reshape(-1, 2)
    x = b[:, 0] + float(tick) * v[:, 0]
    y = b[:, 1] + float(tick) * v[:, 1]
    return [(float(x[i]), float(y[i]), float(b[i, 2])) for i in range(len(b))]


def c5():
    """Config 5: RRT*-FND dynamic replanning. The config-2 problem (RRT*-FN, 1,000-node cap, 4,096 planners) with 50
    circles r = U(1,3) that translate by a per-obstacle velocity U(-0.5,0.5)^2 every replanning tick; each tick the
    robot steps to the next path node (select_branch), the moved obstacles invalidate path nodes (valid_path) and the
    split-off tree is reconnected or regrown (fnd_loop.DeviceFNDReplanner). `edge_invalidation`: the path's EDGES are tested as
    well as its nodes ("edge invalidation" in the config's wording; an extension, the reference's valid_path tests nodes only)."""
    start, goal, ticks, nr = (5.0, 5.0), (95.0, 95.0), 24, 1.0
    rng = np.random.default_rng(5)
    obst, vel = [], []
    while len(obst) < 50:
        c, r, v = rng.uniform(2, 98, 2), rng.uniform(1.0, 3.0), rng.uniform(-0.5, 0.5, 2)
        k = np.arange(ticks + 1)[:, None]
        track = c[None, :] + k * v[None, :]
        # clear of the start now, and of the goal region (goal-reaching leaves lie within 2 * node_radius of it) at every
        # tick: a covered goal-side leaf ends a planner's run (the reference raises IndexError there, src/algorithm.py:358)
        if np.hypot(*(c - start)) < nr + r or (np.hypot(track[:, 0] - goal[0], track[:, 1] - goal[1]) < 3 * nr + r + 0.5).any():
            continue
        obst.append((float(c[0]), float(c[1]), float(r))); vel.append((float(v[0]), float(v[1])))
    vel = np.asarray(vel)
    return dict(name="c5", mode="fn", obstacles=obst, velocity=vel, xy_range=[[0.0, 100.0], [0.0, 100.0]], iter_num=5000,
                step_length=3.0, radius=6.0, node_radius=1.0, bias=0.01, max_nodes=1000, start=start, goal=goal,
                seed_base=5000, planners=4096, ticks=ticks, reconnect_radius=15.0, regrow_iters=100, edge_invalidation=True)


WORKLOADS = {"c1": c1, "c2": c2, "c2r": c2r, "c3": c3, "c4": c4, "c5": c5}


def planner_problems(w, first, count):
    """starts [count,2], goals [count,2], seeds [count], stream_ids [count] for global planners first..first+count-1."""
    starts = np.zeros((count, 2))
    goals = np.zeros((count, 2))
    for i in range(count):
        g = first + i
        if w["start"] is not None:
            starts[i], goals[i] = w["start"], w["goal"]
            continue
        rng = np.random.default_rng([w["seed_base"], g])
        (x0, x1), (y0, y1) = w["xy_range"]
        pts = []
        while len(pts) < 2:
            p = (rng.uniform(x0, x1), rng.uniform(y0, y1))
            if not _occupied(p, w["obstacles"], w["node_radius"]):
                pts.append(p)
        starts[i], goals[i] = pts
    seeds = np.full(count, w["seed_base"], dtype=np.uint64)
    stream_ids = np.arange(first, first + count, dtype=np.uint64)
    return starts, goals, seeds, stream_ids
