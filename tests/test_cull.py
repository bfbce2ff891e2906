"""Obstacle-grid culling must not change a single boolean.

CPU: adversarial probe of the culling rule itself (oracle/cull_probe.c) on the reference's predicate.
GPU: the culled operators (what the fused planner runs) against the un-culled operators and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle as O


def seg_margin(coord_max, slope_max=256.0):
    return 128.0 * (1.0 + slope_max) * coord_max / 2.0 ** 26      # grid_seg_margin in csrc/grid.cuh


@pytest.mark.parametrize("coord_max", [16.0, 256.0, 1024.0, 8192.0])
def test_no_culled_circle_ever_hits(coord_max):
    hits, valid = O.cull_probe(11, 3_000_000, coord_max, 256.0, seg_margin(coord_max), scale=1.0)
    assert valid > 2_000_000 and hits == 0
    # a quarter of the margin is the bound derived in DESIGN.md (the shipped margin carries a 4x safety factor)
    hits, _ = O.cull_probe(12, 3_000_000, coord_max, 256.0, seg_margin(coord_max), scale=0.25)
    assert hits == 0


@pytest.mark.parametrize("slope_max", [1024.0, 4096.0, 65536.0])      # 1024 is what the builders ship (grid_slope_max, grid.cuh)
def test_claim_holds_for_larger_slope_bounds(slope_max):
    """The margin follows the slope bound (it is linear in 1 + slope_max), so the claim is the same claim at any bound."""
    for coord_max in (256.0, 8192.0):
        for scale in (1.0, 0.25):
            hits, valid = O.cull_probe(31, 2_000_000, coord_max, slope_max, seg_margin(coord_max, slope_max), scale=scale)
            assert valid > 1_000_000 and hits == 0


def test_builders_ship_slope_bound_1024_and_follow_the_override():
    """grid_slope_max (grid.cuh): 1024 by default, RRTFND_SLOPE_MAX for A/B runs; the margin follows the bound either way (the host
    builder runs on the CPU, so this needs no GPU)."""
    import subprocess, sys, os
    code = ("import numpy as np\n"
            "from rrt_star_fnd_algorithm_b200.batch import build_obstacle_grid\n"
            "t = np.array([[10.0, 10.0, 2.0, 196.0], [30.0, 40.0, 3.0, 2491.0]])\n"
            "g, _ = build_obstacle_grid(t, 1.0, (0.0, 64.0, 0.0, 64.0), 0.0)\n"
            "print(g.slope_max, g.margin, g.coord_max)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_value, want in ((None, 1024.0), ("4096", 4096.0), ("0.5", 1024.0), ("junk", 1024.0)):
        env = dict(os.environ)
        env.pop("RRTFND_SLOPE_MAX", None)
        if env_value is not None:
            env["RRTFND_SLOPE_MAX"] = env_value
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=env, timeout=120)
        assert out.returncode == 0, out.stderr
        slope, margin, coord = (float(v) for v in out.stdout.split())
        assert slope == want and margin == seg_margin(coord, want)


def test_probe_is_adversarial():
    """With (almost) no margin the same placements do produce hits, i.e. the probe looks in the right place."""
    hits, valid = O.cull_probe(13, 3_000_000, 256.0, 256.0, seg_margin(256.0), scale=1e-6)
    assert hits > 100


# ----------------------------------------------------------------------------------------------------- GPU
def _edges(rng, n, lo, hi, obst):
    """Random segments: short and long, plus steep / vertical / endpoint-on-circle corner cases."""
    p1 = rng.uniform(lo, hi, (n, 2))
    L = np.exp(rng.uniform(np.log(1e-3), np.log(hi - lo), n))
    th = rng.uniform(0, 2 * np.pi, n)
    p2 = np.clip(p1 + L[:, None] * np.column_stack([np.cos(th), np.sin(th)]), lo, hi)
    k = n // 10
    p2[:k, 0] = p1[:k, 0]                                         # vertical (infinite-line quirk, all circles)
    p2[k:2 * k, 0] = p1[k:2 * k, 0] + rng.uniform(-1e-6, 1e-6, k)  # steeper than slope_max: falls back to the table
    ob = np.asarray(obst)
    j = rng.integers(0, len(ob), k)
    a = rng.uniform(0, 2 * np.pi, k)                              # an endpoint exactly on / just off a circle's rim
    p1[2 * k:3 * k] = ob[j, :2] + (ob[j, 2] * (1 + rng.choice([0, 1e-15, -1e-15, 1e-9], k)))[:, None] * np.column_stack([np.cos(a), np.sin(a)])
    return p1, p2


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["c2", "c3", "c1"])
def test_culled_edge_and_point_operators_equal_the_full_ones(workload):
    import torch
    from rrt_star_fnd_algorithm_b200 import _abi, workloads as W
    from rrt_star_fnd_algorithm_b200.batch import build_obstacle_grid, obstacle_table
    w = W.WORKLOADS[workload]()
    (x0, x1), (y0, y1) = w["xy_range"]
    tab = obstacle_table(w["obstacles"])
    g, arrs = build_obstacle_grid(tab[:, :3], w["node_radius"], (x0, x1, y0, y1))
    dev = torch.device("cuda:0")
    gt = [torch.from_numpy(a).to(dev) for a in arrs]
    g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in gt]
    rng = np.random.default_rng(3)
    n = 400_000
    p1, p2 = _edges(rng, n, x0, x1, w["obstacles"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tp = [d(p1[:, 0]), d(p1[:, 1]), d(p2[:, 0]), d(p2[:, 1])]
    tob = d(tab)
    full = torch.zeros(n, dtype=torch.uint8, device=dev)
    cull = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    _abi.check(_abi.lib.rrtfnd_edges_vs_circles(*[ptr(t) for t in tp], n, ptr(tob), len(tab), ptr(full), None), "full")
    _abi.check(_abi.lib.rrtfnd_edges_vs_circles_grid(*[ptr(t) for t in tp], n, ptr(tob), len(tab), C.byref(g), ptr(cull), None), "culled")
    torch.cuda.synchronize()
    full, cull = full.cpu().numpy(), cull.cpu().numpy()
    assert 0.02 < full.mean() < 0.98
    bad = np.flatnonzero(full != cull)
    assert bad.size == 0, (bad[:5], p1[bad[:5]], p2[bad[:5]])
    want = O.through_obstacle_batch(p1, p2, w["obstacles"])      # and both equal the CPU oracle (no pre-check, no culling)
    assert np.array_equal(full.astype(bool), want)
    # points
    pts = rng.uniform(x0, x1, (n, 2))
    tx, ty = d(pts[:, 0]), d(pts[:, 1])
    of = torch.zeros(n, dtype=torch.uint8, device=dev)
    oc = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    _abi.check(_abi.lib.rrtfnd_points_occupied(ptr(tx), ptr(ty), n, ptr(tob), len(tab), w["node_radius"], ptr(of), None), "full")
    _abi.check(_abi.lib.rrtfnd_points_occupied_grid(ptr(tx), ptr(ty), n, ptr(tob), len(tab), C.byref(g), w["node_radius"], ptr(oc), None), "culled")
    torch.cuda.synchronize()
    of, oc = of.cpu().numpy(), oc.cpu().numpy()
    assert np.array_equal(of, oc)
    for i in rng.integers(0, n, 300):
        assert bool(of[i]) == O.is_occupied(pts[i], w["obstacles"], w["node_radius"])


@pytest.mark.gpu
def test_root_precheck_borderline_segments_equal_the_oracle():
    """Segments whose endpoint x sits within ulps of an intersection root (the pre-check's margin decides nothing there:
    they must fall through to the exact arithmetic), plus tangent lines and segments wholly inside a circle."""
    import torch
    from rrt_star_fnd_algorithm_b200 import _abi
    from rrt_star_fnd_algorithm_b200.batch import obstacle_table
    rng = np.random.default_rng(17)
    obst = [(float(x), float(y), float(r)) for x, y, r in zip(rng.uniform(10, 190, 40), rng.uniform(10, 190, 40), rng.uniform(1, 9, 40))]
    ob = np.asarray(obst)
    n = 300_000
    j = rng.integers(0, len(ob), n)
    th = rng.uniform(0, 2 * np.pi, n)
    m = np.tan(rng.uniform(-1.5, 1.5, n))                             # line through a rim point with slope m
    rim = ob[j, :2] + ob[j, 2:3] * np.column_stack([np.cos(th), np.sin(th)])
    eps = rng.choice([0.0, 1e-16, -1e-16, 3e-16, -3e-16, 1e-13, -1e-13, 1e-10, -1e-10, 1e-7, 1e-3], n) * np.maximum(np.abs(rim[:, 0]), 1.0)
    L = np.exp(rng.uniform(np.log(1e-3), np.log(60.0), n)) * rng.choice([-1.0, 1.0], n)
    p1 = np.column_stack([rim[:, 0] + eps, rim[:, 1] + m * eps])      # an endpoint (almost) exactly on the circle, on the line
    p2 = np.column_stack([p1[:, 0] + L, p1[:, 1] + m * L])
    k = n // 6
    inside = ob[j[:k], :2] + 0.3 * ob[j[:k], 2:3] * rng.uniform(-1, 1, (k, 2))      # both endpoints inside one circle
    p1[:k], p2[:k] = inside, ob[j[:k], :2] + 0.3 * ob[j[:k], 2:3] * rng.uniform(-1, 1, (k, 2))
    tang = slice(k, 2 * k)                                             # tangent lines: delta ~ 0
    nrm = np.column_stack([np.cos(th[tang]), np.sin(th[tang])]); tan_dir = np.column_stack([-nrm[:, 1], nrm[:, 0]])
    c = ob[j[tang], :2] + ob[j[tang], 2:3] * nrm * (1 + rng.choice([0, 1e-15, -1e-15, 1e-12, -1e-12], k)[:, None])
    p1[tang], p2[tang] = c - tan_dir * rng.uniform(0.1, 30, (k, 1)), c + tan_dir * rng.uniform(0.1, 30, (k, 1))
    dev = torch.device("cuda:0")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tp = [d(p1[:, 0]), d(p1[:, 1]), d(p2[:, 0]), d(p2[:, 1])]
    tob = d(obstacle_table(obst))
    out = torch.zeros(n, dtype=torch.uint8, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    _abi.check(_abi.lib.rrtfnd_edges_vs_circles(*[ptr(t) for t in tp], n, ptr(tob), len(obst), ptr(out), None), "full")
    torch.cuda.synchronize()
    got, want = out.cpu().numpy().astype(bool), O.through_obstacle_batch(p1, p2, obst)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (bad[:5], p1[bad[:5]], p2[bad[:5]])
    assert 0.1 < want.mean() < 0.9
