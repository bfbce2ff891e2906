"""Multi-GPU plumbing on CPU: world-size-2 gloo. Ranks own contiguous blocks of planners and meet once, after growth,
to gather the per-planner records and broadcast the globally best path (parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrt_star_fnd_algorithm_b200._structs import PLANNER_BYTES, PLANNER_DTYPE
from rrt_star_fnd_algorithm_b200.parallel import gather_records, shard_range


def test_shard_range_partitions_every_planner_once():
    for n in (1, 7, 8, 16384, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_workload_problems_do_not_depend_on_the_sharding():
    from rrt_star_fnd_algorithm_b200 import workloads as W
    w = W.c3()
    s_all, g_all, seeds, sids = W.planner_problems(w, 0, 12)
    for world in (2, 3):
        parts = [W.planner_problems(w, *shard_range(12, r, world)) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), s_all)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), g_all)
        assert np.array_equal(np.concatenate([p[3] for p in parts]), sids)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(rank, P, path_cap):
    """Synthetic planner records: planner i of rank r has best cost 100 - (r * P + i) % 7, rank 1 planner 2 has none."""
    pl = np.zeros(P, dtype=PLANNER_DTYPE)
    paths = np.zeros((P, path_cap), dtype=np.int32)
    for i in range(P):
        g = rank * P + i
        pl["best_cost"][i] = 100.0 - (g % 7)
        pl["best_len"][i] = 3 + g % 4
        pl["iter"][i] = 10 + g
        paths[i, :pl["best_len"][i]] = np.arange(pl["best_len"][i]) + 1000 * g
    if rank == 1:
        pl["best_cost"][2], pl["best_len"][2] = np.inf, 0
    return torch.from_numpy(pl.view(np.uint8).reshape(P, PLANNER_BYTES).copy()), torch.from_numpy(paths)


def _worker(rank, world, port, P, path_cap, out_dir, all_inf):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec, paths = _records(rank, P, path_cap)
        if all_inf:
            v = rec.numpy().reshape(-1).view(PLANNER_DTYPE)
            v["best_cost"], v["best_len"] = np.inf, 0
        allrec, gidx, path = gather_records(rec, paths)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), allrec=allrec.numpy(), gidx=gidx, path=path.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("all_inf", [False, True])
def test_gather_records_world_size_2_gloo(tmp_path, all_inf):
    world, P, path_cap = 2, 5, 16
    mp.spawn(_worker, args=(world, _free_port(), P, path_cap, str(tmp_path), all_inf), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    # every rank ends with the same gathered records, the same winner and the winner's path
    for k in ("allrec", "gidx", "path"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    allrec = outs[0]["allrec"].reshape(-1).view(PLANNER_DTYPE)
    assert len(allrec) == world * P and list(allrec["iter"]) == [10 + g for g in range(world * P)]
    if all_inf:
        assert int(outs[0]["gidx"]) == -1 and len(outs[0]["path"]) == 0
        return
    costs = allrec["best_cost"]
    want = int(np.flatnonzero(costs == costs.min())[0])          # lowest global index on ties
    assert int(outs[0]["gidx"]) == want == 6                     # 100 - 6 is the minimum, first reached by planner 6 (rank 1)
    assert list(outs[0]["path"]) == list(np.arange(3 + want % 4) + 1000 * want)


def test_workloads_are_deterministic_and_valid():
    """Maps and problems are pure functions of (workload, planner index); starts and goals are outside inflated obstacles."""
    import oracle as O
    from rrt_star_fnd_algorithm_b200 import workloads as W
    for name in ("c2", "c3", "c4"):
        a, b = W.WORKLOADS[name](), W.WORKLOADS[name]()
        assert a["obstacles"] == b["obstacles"] and len(a["obstacles"]) in (200, 500)
        s, g, seeds, sids = W.planner_problems(a, 100, 6)
        s2, g2, _, sids2 = W.planner_problems(b, 100, 6)
        assert np.array_equal(s, s2) and np.array_equal(g, g2) and list(sids) == list(range(100, 106)) == list(sids2)
        for p in range(6):
            assert not O.is_occupied(s[p], a["obstacles"], a["node_radius"]) and not O.is_occupied(g[p], a["obstacles"], a["node_radius"])


def test_bench_config_is_shared_by_both_arms():
    import argparse, importlib.util, os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    from rrt_star_fnd_algorithm_b200 import workloads as W
    args = argparse.Namespace(launch=0, near_cap=64)
    c = bench.bench_config(W.c3(), 16384, 1, args, 5002)
    assert c["workload"].startswith("c3: star x 16384 planners in total, 16384 per GPU (strong scaling), 500 circular obstacles, 5000 iterations")
    assert bench.DEFAULT_PLANNERS["c3"] == 16384 and bench.METRIC == "rrt_star_tree_extensions_per_sec"
    # config 3 as BASELINE.json words it: 16,384 planners SHARDED over the GPUs (strong scaling) unless a per-GPU batch is asked for
    ns = lambda **k: argparse.Namespace(workload="c3", planners=0, scaling="strong", **k)
    assert [bench.planner_layout(ns(), 8, r) for r in (0, 7)] == [(2048, 0, "strong"), (2048, 14336, "strong")]
    assert bench.planner_layout(ns(), 1, 0) == (16384, 0, "strong")
    assert bench.planner_layout(argparse.Namespace(workload="c3", planners=0, scaling="weak"), 4, 2) == (16384, 32768, "weak")
    assert bench.planner_layout(argparse.Namespace(workload="c3", planners=512, scaling="strong"), 4, 3) == (512, 1536, "weak")
    assert bench.planner_layout(argparse.Namespace(workload="c4", planners=0, scaling="strong"), 8, 5) == (1, 5, "weak")   # replicas
    c8 = bench.bench_config(W.c3(), 2048, 8, args, 5002, "strong")
    assert c8["planners_total"] == 16384 and c8["planners_per_gpu"] == 2048


def test_bench_cpu_legs_run_on_small_samples():
    """The CPU legs of bench.py (cpu_baseline / --impl reference): the C port on a few planners of c3 and the whole config-5
    step (planning + replanning ticks through oracle/fnd_loop.py) on a few planners."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    from rrt_star_fnd_algorithm_b200 import workloads as W
    ext, edge, dt, out = bench.cpu_port_run(W.c3(), 0, 4, 200, 2)
    assert int(out["iter"].sum()) == 800 and ext > 0 and edge > ext
    w = W.c5()
    ext, moves, dt, total = bench.cpu_c5_run(dict(w, iter_num=1500, max_nodes=400), 0, 4, 2)
    assert total >= 4 * 1500 and moves > 0
    cfg = bench.bench_config(w, 4096, 1, __import__("argparse").Namespace(launch=0, near_cap=64), 1258)
    assert "24 replanning ticks" in cfg["workload"]
