"""FP64 roofline of the collision operators ("collision kernel: FP64 pipe"): python scripts/collision_roofline.py
Times rrtfnd_edges_vs_circles (one warp per edge, every circle of the table from shared memory; no early exit across lanes until a
hit) and rrtfnd_edges_vs_circles_grid (one lane per edge, culled) on the c3 map, counts the fp64 operations the tests performed
(DADD / DMUL / compare = 1; IEEE division and square root in DADD-equivalents as measured by _tools/fp64_rate) and prints one JSON line
with the fraction of the measured DADD / DMUL issue rate."""
import ctypes as C, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rrt_star_fnd_algorithm_b200 import _abi, workloads as W
from rrt_star_fnd_algorithm_b200.batch import build_obstacle_grid, obstacle_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rate = json.loads(subprocess.run([os.path.join(ROOT, "rrt_star_fnd_algorithm_b200", "_tools", "fp64_rate")], capture_output=True, text=True).stdout)
peak, div_eq, sqrt_eq = rate["dadd_per_s"], rate["ddiv_in_dadd_equivalents"], rate["dsqrt_in_dadd_equivalents"]
w = W.c3()
tab = obstacle_table(w["obstacles"])
M = len(tab)
rng = np.random.default_rng(1)
n = 1 << 20
p1 = rng.uniform(0, 200, (n, 2))
ang, length = rng.uniform(0, 2 * np.pi, n), rng.uniform(0.5, 10.0, n)         # planner-like segments: up to two steps long
p2 = p1 + np.column_stack([np.cos(ang), np.sin(ang)]) * length[:, None]
dev = torch.device("cuda:0")
cols = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1])]
tab_t = torch.from_numpy(tab).to(dev)
g, arrs = build_obstacle_grid(tab, w["node_radius"], (0.0, 200.0, 0.0, 200.0))
gt = [torch.from_numpy(a).to(dev) for a in arrs]
g.seg_start, g.seg_items, g.pt_start, g.pt_items = [t.data_ptr() for t in gt]
out = torch.zeros(n, dtype=torch.uint8, device=dev)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

full = lambda: _abi.check(_abi.lib.rrtfnd_edges_vs_circles(*[t.data_ptr() for t in cols], n, tab_t.data_ptr(), M, out.data_ptr(), None), "edges")
cull = lambda: _abi.check(_abi.lib.rrtfnd_edges_vs_circles_grid(*[t.data_ptr() for t in cols], n, tab_t.data_ptr(), M, C.byref(g), out.data_ptr(), None), "edges grid")
ms_full = timed(full); hit_full = out.cpu().numpy().astype(bool)
ms_cull = timed(cull); hit_cull = out.cpu().numpy().astype(bool)
assert np.array_equal(hit_full, hit_cull)
# operations of the un-culled operator: every lane of an edge's warp walks its stride of the table; lanes stop at the end of the
# tile in which any lane hit. Count exactly on the host for a sample of edges: per circle 11 add / mul (b, c, delta) + 1 compare;
# when delta >= 0 also sqrt + 2 div + 2 add + 4 compares; per edge the line set-up: 1 div + 9 add / mul.
k = 4096
a, b = p1[:k], p2[:k]
m = (b[:, 1] - a[:, 1]) / (b[:, 0] - a[:, 0]); kk = a[:, 1] - m * a[:, 0]
ox, oy, r, K = tab[:, 0][None], tab[:, 1][None], tab[:, 2][None], tab[:, 3][None]
bb = 2 * ((-ox + (m * kk)[:, None]) - m[:, None] * oy)
cc = (K - (2 * oy) * kk[:, None]) + (kk * kk)[:, None]
delta = bb * bb - (4 * (1 + m * m))[:, None] * cc
with np.errstate(invalid="ignore"):
    sq = np.sqrt(np.where(delta >= 0, delta, np.nan))
    a2 = (2 * (1 + m * m))[:, None]
    x1, x2 = (-bb + sq) / a2, (-bb - sq) / a2
    lo, hi = np.minimum(a[:, 0], b[:, 0])[:, None], np.maximum(a[:, 0], b[:, 0])[:, None]
    hit = (delta >= 0) & (((x1 > lo) & (x1 < hi)) | ((x2 > lo) & (x2 < hi)))
first = np.where(hit.any(1), hit.argmax(1), M - 1)
rows_run = first // 32 + 1                                  # the warp votes after every row of 32 circles and stops at the first hit
tests_sample = np.minimum(rows_run * 32, M)
idx = np.arange(M)[None, :] < tests_sample[:, None]
pos = float((delta >= 0)[idx].mean())
ops_per_test = 12 + pos * (sqrt_eq + 2 * div_eq + 6)
tests_full = float(n) * float(tests_sample.mean())          # estimated from a 4,096-edge sample evaluated on the host
ops_full = tests_full * ops_per_test + n * (div_eq + 9)
line = {"fp64_peak_dadd_per_s": peak, "ddiv_dadd_eq": div_eq, "dsqrt_dadd_eq": sqrt_eq, "edges": n, "circles": M, "delta_nonneg_fraction": pos,
        "dadd_equivalents_per_circle_test": ops_per_test,
        "uncull": {"ms": ms_full, "edges_per_s": n / ms_full * 1e3, "circle_tests_per_s": tests_full / ms_full * 1e3,
                   "dadd_eq_per_s": ops_full / ms_full * 1e3, "frac_of_fp64_issue_peak": ops_full / ms_full * 1e3 / peak},
        "culled": {"ms": ms_cull, "edges_per_s": n / ms_cull * 1e3, "speedup": ms_full / ms_cull}, "blocked_fraction": float(hit_full.mean())}
print(json.dumps(line))
