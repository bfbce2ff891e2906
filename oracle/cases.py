"""Parity cases shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE).

A case is everything needed to run one planner: mode, map, start/goal, parameters and the Philox
stream (seed, stream_id) that supplies the random words. Maps are generated deterministically here.
"""
import numpy as np

XY_MAIN = [[-210, 210], [-210, 210]]
# reconstructed src/main.py map: 9 cylinders r=15 on a 110-spaced lattice, range [-210, 210]^2
OBST_MAIN = [(float(x), float(y), 15.0) for x in (-110, 0, 110) for y in (-110, 0, 110)]
# start / goal drawn as src/main.py:66-67 does after random.seed(0)
START_MAIN = (144.65717764052022, 108.34084923492702)
GOAL_MAIN = (-33.35993605104511, -101.2549648769554)


def float_circles(seed, n, lo, hi, rmin, rmax, start, goal, node_radius):
    """n circles with float centres U(lo,hi)^2 and radius U(rmin,rmax), clear of start/goal (cf. src/map.py:61)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c = rng.uniform(lo, hi, 2)
        r = rng.uniform(rmin, rmax)
        if np.hypot(c[0] - start[0], c[1] - start[1]) < node_radius + r:
            continue
        if np.hypot(c[0] - goal[0], c[1] - goal[1]) < node_radius + r:
            continue
        out.append((float(c[0]), float(c[1]), float(r)))
    return out


def int_circles(seed, n, lo, hi, r, start, goal, node_radius):
    """Integer-valued circles like Map.generate_obstacles (src/map.py:53-64)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c = (int(rng.integers(lo, hi)), int(rng.integers(lo, hi)))
        if np.hypot(c[0] - start[0], c[1] - start[1]) < node_radius + r:
            continue
        if np.hypot(c[0] - goal[0], c[1] - goal[1]) < node_radius + r:
            continue
        out.append((c[0], c[1], r))
    return out


def _case(name, mode, seed, start, goal, xy, obst, iter_num, step, radius, node_radius, bias, max_nodes=200):
    return dict(name=name, mode=mode, seed=seed, stream_id=0, start=start, goal=goal, xy_range=xy, obstacles=obst,
                iter_num=iter_num, step_length=step, radius=radius, node_radius=node_radius, bias=bias,
                max_nodes=max_nodes)


def all_cases():
    xy100 = [[0, 100], [0, 100]]
    xy200 = [[0, 200], [0, 200]]
    c2 = float_circles(2, 60, 2, 98, 0.5, 3.0, (5, 5), (95, 95), 1)
    c2d = float_circles(22, 200, 2, 98, 0.5, 2.0, (5, 5), (95, 95), 1)
    ci = int_circles(3, 40, 7, 193, 7, (50, 50), (150, 150), 5)
    return [
        _case("main_star", "star", 7, START_MAIN, GOAL_MAIN, XY_MAIN, OBST_MAIN, 400, 25, 50, 20, 0.01),
        _case("main_rrt", "rrt", 8, START_MAIN, GOAL_MAIN, XY_MAIN, OBST_MAIN, 400, 25, None, 20, 0.05),
        _case("main_rrt_nobias", "rrt", 18, START_MAIN, GOAL_MAIN, XY_MAIN, OBST_MAIN, 150, 25, None, 2, 0.0),
        _case("main_fn", "fn", 9, START_MAIN, GOAL_MAIN, XY_MAIN, OBST_MAIN, 600, 25, 50, 20, 0.01, 100),
        _case("float_fn", "fn", 10, (5.0, 5.0), (95.0, 95.0), xy100, c2, 800, 3, 6, 1, 0.02, 150),
        _case("float_star", "star", 11, (5, 5), (95, 95), xy100, c2, 600, 3, 6, 1, 0.05),
        _case("dense_fn", "fn", 14, (5.0, 5.0), (95.0, 95.0), xy100, c2d, 700, 3, 6, 1, 0.01, 300),
        _case("int_star", "star", 12, (50, 50), (150, 150), xy200, ci, 500, 15, 30, 5, 0.02),
        _case("int_fn_tiny", "fn", 13, (50, 50), (150, 150), xy200, ci, 500, 15, 15, 5, 0.0, 30),
        _case("main_star_long", "star", 15, START_MAIN, GOAL_MAIN, XY_MAIN, OBST_MAIN, 1500, 25, 50, 20, 0.01),
    ]


def case_by_name(name):
    for c in all_cases():
        if c["name"] == name:
            return c
    raise KeyError(name)


def words_needed(case):
    return 400 * case["iter_num"] + 4096


def params_for(case, *, flags=None, capacity=None, stream_mode=0, n_words=0, near_cap=1024, finish_cap=4096,
               path_cap=4096):
    """rrtfnd_params_t for a case (defaults: evaluate choose-parent, trace on, word-buffer stream)."""
    from rrt_star_fnd_algorithm_b200._structs import Params, FLAG_EVAL_CHOOSE_PARENT, FLAG_TRACE
    p = Params()
    p.mode = {"rrt": 0, "star": 1, "fn": 2}[case["mode"]]
    p.iter_num = case["iter_num"]
    p.max_nodes = case["max_nodes"]
    p.capacity = capacity if capacity is not None else case["iter_num"] + 2
    p.n_obst = len(case["obstacles"])
    p.near_cap, p.finish_cap, p.path_cap = near_cap, finish_cap, path_cap
    p.stream_mode = stream_mode
    p.flags = (FLAG_EVAL_CHOOSE_PARENT | FLAG_TRACE) if flags is None else flags
    p.trace_cap = case["iter_num"] + 1
    p.words_per_planner = n_words
    p.max_attempts = 0
    p.step_length = case["step_length"]
    p.radius = case["radius"] or 0.0
    p.node_radius = case["node_radius"]
    p.goal_node_radius = case.get("goal_node_radius", case["node_radius"])
    p.bias = case["bias"]
    (p.x_lo, p.x_hi), (p.y_lo, p.y_hi) = case["xy_range"]
    return p
