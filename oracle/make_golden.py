"""Generate tests/golden/*.npz by running the UNMODIFIED Python reference (TEST INFRASTRUCTURE).

Run in the build container only (`python -m oracle.make_golden`); it needs /root/reference. The
fixtures it writes are committed, so neither the CPU tests nor the GPU box need the reference.

Per case (oracle/cases.py) the fixture holds: the final tree (ids, x, y, parent id, edge cost, child
count), iter, best path ids + cost, number of random words consumed, attempts, through_obstacle call
count, and the per-iteration records (id_new, id_near, near-set size, choose_parent_kdtree's returned
edge under ascending ball order, rewired-node count, removed id).
`collision.npz` holds segments recorded from the reference's own through_obstacle calls with their
booleans; `occupancy.npz` the same for Map.is_occupied_c.
"""
import os
import sys

import numpy as np

from . import cases as K
from . import refharness as H
from .philox import philox_words

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def golden_for_case(case, keep_edges=0):
    words = philox_words(case["seed"], case["stream_id"], K.words_needed(case))
    ref = H.run_reference(case["mode"], start=case["start"], goal=case["goal"], xy_range=case["xy_range"],
                          obstacles=case["obstacles"], words=words, iter_num=case["iter_num"],
                          step_length=case["step_length"], radius=case["radius"], node_radius=case["node_radius"],
                          bias=case["bias"], max_nodes=case["max_nodes"], keep_edges=keep_edges)
    rec = ref["rec"]
    n = len(rec["id_new"])
    star = case["mode"] != "rrt"
    out = dict(
        ids=ref["ids"], x=ref["x"], y=ref["y"], parent=ref["parent"], cost=ref["cost"], nchild=ref["nchild"],
        iter=np.int64(ref["iter"]), best_path=ref["best_path"], best_cost=np.float64(ref["best_cost"]),
        words_used=np.int64(ref["words_used"]), last_id=np.int64(ref["last_id"]),
        attempts=np.int64(rec["attempts"]), edge_calls=np.int64(rec["edge_calls"]),
        rec_id_new=np.array(rec["id_new"], np.int32), rec_id_near=np.array(rec["id_near"], np.int32),
        rec_removed=np.array(rec["removed"], np.int32),
        rec_n_near=np.array(rec["n_near"] if star else [0] * n, np.int32),
        rec_cp_parent=np.array(rec["cp_parent"] if star else [-1] * n, np.int32),
        rec_cp_cost=np.array(rec["cp_cost"] if star else [0.0] * n, np.float64),
        rec_n_rewired=np.array([len(r) for r in rec["rewired"]] if star else [0] * n, np.int32),
    )
    return out, ref


def main():
    os.makedirs(OUT, exist_ok=True)
    edges = []
    for case in K.all_cases():
        g, ref = golden_for_case(case, keep_edges=4000)
        np.savez_compressed(os.path.join(OUT, f"case_{case['name']}.npz"), **g)
        e = np.array(ref["rec"]["edges"], dtype=np.float64).reshape(-1, 5)
        edges.append((case["name"], e))
        print(f"{case['name']:16s} mode={case['mode']:4s} iter={int(g['iter'])} live={len(g['ids'])} "
              f"attempts={int(g['attempts'])} edges={int(g['edge_calls'])} best={float(g['best_cost']):.4f} "
              f"words={int(g['words_used'])}")
    # collision booleans recorded from the reference's own calls (first 4000 per case)
    col = {}
    for name, e in edges:
        col[f"{name}_seg"] = e[:, :4]
        col[f"{name}_hit"] = e[:, 4].astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "collision.npz"), **col)

    # synthetic stress: random segments / points against the case maps, evaluated by the reference's own
    # Line + through_obstacle + Map.is_occupied_c (includes vertical and near-tangent segments)
    ref = H.load_reference()
    rng = np.random.default_rng(99)
    occ = {}
    for name in ("float_fn", "int_star", "main_star", "dense_fn"):
        case = K.case_by_name(name)
        m, _ = H.make_map_and_graph(ref, case["start"], case["goal"], case["xy_range"], case["node_radius"],
                                    case["obstacles"])
        (x0, x1), (y0, y1) = case["xy_range"]
        n = 3000
        p1 = np.stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)], 1)
        ang = rng.uniform(0, 2 * np.pi, n)
        ln = rng.uniform(0, 3 * case["step_length"], n)
        p2 = p1 + np.stack([np.cos(ang), np.sin(ang)], 1) * ln[:, None]
        p2[::17, 0] = p1[::17, 0]                      # exactly vertical segments (src/line.py:9)
        ob = np.array(case["obstacles"], dtype=np.float64)
        # near-tangent: aim a third of the segments at a circle's rim
        j = rng.integers(0, len(ob), n)
        rim = ob[j, :2] + ob[j, 2:3] * np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.98, 1.02, (n, 1))
        sel = np.arange(n) % 3 == 0
        p2[sel] = rim[sel]
        hits = np.zeros(n, np.uint8)
        for i in range(n):
            line = ref.line.Line((float(p1[i, 0]), float(p1[i, 1])), (float(p2[i, 0]), float(p2[i, 1])))
            hits[i] = ref.algorithm.through_obstacle(line, m.obstacles_c)
        col[f"synth_{name}_seg"] = np.concatenate([p1, p2], 1)
        col[f"synth_{name}_hit"] = hits
        pts = np.stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)], 1)
        pts[::2] = ob[j[::2], :2] + (ob[j[::2], 2:3] + case["node_radius"]) * np.stack(
            [np.cos(ang[::2]), np.sin(ang[::2])], 1) * rng.uniform(0.97, 1.03, (len(pts[::2]), 1))
        occ[f"{name}_pts"] = pts
        occ[f"{name}_occ"] = np.array([m.is_occupied_c((float(a), float(b))) for a, b in pts], np.uint8)
        print(f"synthetic {name}: {hits.sum()}/{n} segments hit, {occ[f'{name}_occ'].sum()}/{n} points occupied")
    np.savez_compressed(os.path.join(OUT, "collision.npz"), **col)
    np.savez_compressed(os.path.join(OUT, "occupancy.npz"), **occ)


if __name__ == "__main__":
    sys.exit(main())
