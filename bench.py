#!/usr/bin/env python
"""bench.py — RRT* tree extensions/s and edge collision checks/s on N B200s (see DESIGN.md "Measurement").

A *step* is one complete run of the workload's planners owned by this rank: every tree is re-initialised
to its root and grown for `iter_num` successful iterations (sample -> occupancy -> nearest visible ->
steer -> insert -> near set -> collision checks -> choose-parent -> rewire -> goal / best-path bookkeeping).
Scaling (`--scaling`, default strong): BASELINE.json's config 3 is 16,384 planners SHARDED over the GPUs, so by default the
workload's planners are split into contiguous blocks (parallel.shard_range: 16,384 / N per rank) and the total work is fixed;
for N > 1 the line also carries `weak_scaling`, the same measurement with a full 16,384-planner batch on EVERY GPU
(`--scaling weak` makes that the headline; `--planners` = planners per GPU, weak by definition). A planner's problem is a
function of its global index only, so its tree does not depend on the sharding. After its planners finish, a rank takes
part in one NCCL all_gather of the per-planner result records and one broadcast of the globally best path.

  value        extensions/s with inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e          the same metric through rrtfnd_plan_host(): HOST buffers in, HOST buffers out, copies timed
  roofline     algorithmic bytes of the fused planner kernel (16 B per node visited by the nearest / near-set
               scans + 12 B per parent-chain hop + 32 B per inserted node) / its CUDA-event duration, against
               the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle port (oracle/rrt_oracle.c) on all host cores, bounded sample of the same workload

`--impl reference` times that CPU port instead of the GPU path (the reference itself is Python and cannot
travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rrt_star_tree_extensions_per_sec"
UNIT = "extensions/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c2r", "c3", "c4", "c5"])
    ap.add_argument("--planners", type=int, default=0, help="planners per GPU (weak scaling; 0 = the workload's own count, see --scaling)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the workload's planners sharded over the GPUs (config as written); weak: that many on every GPU")
    ap.add_argument("--iters", type=int, default=0, help="iterations per planner (0 = workload default)")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--near-cap", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# default planners per GPU. c3 at N=1 is BASELINE.json's config 3 as written (16,384 planners on one B200, 3 GB of tree
# state); with weak scaling every further GPU gets a batch of the same size. 16,384 planners are ~7 waves of resident
# warps per SM, which also evens out the planner-to-planner variance that leaves SMs idle at the end of a 1-wave launch.
DEFAULT_PLANNERS = {"c1": 1, "c2": 4096, "c2r": 4096, "c3": 16384, "c4": 1, "c5": 4096}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_workload(args):
    from rrt_star_fnd_algorithm_b200 import workloads as W
    w = W.WORKLOADS[args.workload]()
    if args.iters:
        w["iter_num"] = args.iters
    P = args.planners or DEFAULT_PLANNERS[args.workload]
    return w, P


def oracle_params(w, flags):
    from rrt_star_fnd_algorithm_b200 import _structs as S
    p = S.Params()
    p.mode = {"rrt": 0, "star": 1, "fn": 2}[w["mode"]]
    p.iter_num, p.max_nodes, p.capacity = w["iter_num"], w["max_nodes"], w["iter_num"] + 2
    p.n_obst, p.near_cap, p.finish_cap, p.path_cap = len(w["obstacles"]), 4096, 4096, 4096
    p.stream_mode, p.flags = S.STREAM_PHILOX, flags
    p.step_length, p.radius, p.node_radius, p.bias = w["step_length"], w["radius"], w["node_radius"], w["bias"]
    (p.x_lo, p.x_hi), (p.y_lo, p.y_hi) = w["xy_range"]
    p.goal_node_radius = w["node_radius"]
    return p


def cpu_port_run(w, first, count, iters, threads):
    """Time the CPU oracle port on `count` planners x `iters` iterations. Returns (ext/s, edge/s, seconds, records)."""
    import oracle as O
    from rrt_star_fnd_algorithm_b200 import _structs as S
    from rrt_star_fnd_algorithm_b200 import workloads as W
    ww = dict(w, iter_num=iters)
    s, g, seeds, sids = W.planner_problems(ww, first, count)
    prm = oracle_params(ww, S.FLAG_EVAL_CHOOSE_PARENT)
    sg = np.concatenate([s, g], 1)
    t0 = time.perf_counter()
    out, bad = O.run_batch(prm, sg, np.stack([seeds, sids], 1), ww["obstacles"], n_threads=threads)
    dt = time.perf_counter() - t0
    return out["iter"].sum() / dt, out["n_edge_checks"].sum() / dt, dt, out


def cpu_c5_run(w, first, count, threads):
    """Config 5 on the CPU: plan `count` planners with the C port, then step each through the replanning ticks
    (oracle/fnd_loop.py). Returns (extensions/s, moves/s, seconds, extensions)."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    from oracle.fnd_loop import OracleReplanner
    from rrt_star_fnd_algorithm_b200 import workloads as W
    from rrt_star_fnd_algorithm_b200.fnd_loop import moved_obstacles
    s, g, seeds, sids = W.planner_problems(w, first, count)
    T = w["ticks"]
    tables = [moved_obstacles(w["obstacles"], w["velocity"], k) for k in range(1, T + 1)]

    def one(p):
        prm = oracle_params(w, 0)
        t = O.OracleTree(s[p], w["iter_num"] + 2 + T * w["regrow_iters"])
        t.run(prm, w["obstacles"], g[p], seed=int(seeds[p]), stream_id=int(sids[p]))
        ext, moves = t.planner().iter, 0
        r = OracleReplanner(t, prm, g[p], seeds[p], sids[p], w["node_radius"], w["reconnect_radius"], w["regrow_iters"],
                            edge_invalidation=w.get("edge_invalidation", False))
        for ob in tables:
            o = r.tick(ob)
            ext += sum(o["regrow_iters"]); moves += o["current"] >= 0
        t.close()
        return ext, moves

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, range(count)))
    dt = time.perf_counter() - t0
    ext, moves = sum(r[0] for r in res), sum(r[1] for r in res)
    return ext / dt, moves / dt, dt, ext


def size_cpu_sample(w, budget_s, cores):
    """(planners, iterations) of a CPU sample that takes roughly budget_s on `cores` threads."""
    cpu_port_run(w, 0, 1, 20, 1)                          # load / build the checker outside the calibration
    cal_it = min(w["iter_num"], 1000)
    _, _, dt, _ = cpu_port_run(w, 0, cores, cal_it, cores)
    # per-iteration cost of RRT* grows ~linearly with tree size (time ~ it^2); FN saturates at the node cap
    expo = 2.0 if w["mode"] == "star" else 1.0
    full = dt * (w["iter_num"] / cal_it) ** expo            # estimated seconds for `cores` planners at full length
    if full <= budget_s:
        return cores * int(max(1, min(64, budget_s // max(full, 1e-3)))), w["iter_num"]
    iters = int(max(cal_it, cal_it * (budget_s / max(dt, 1e-3)) ** (1.0 / expo)))
    return cores, min(iters, w["iter_num"])


def cpu_baseline(w, budget_s):
    """Bounded sample of the same workload on all host cores (C port of the reference algorithm)."""
    cores = len(os.sched_getaffinity(0))
    if w["name"] == "c5":
        _, _, dt, _ = cpu_c5_run(w, 0, cores, cores)
        count = int(min(4096, max(cores, cores * (budget_s // max(dt, 1e-3)))))
        ext, moves, dt, _ = cpu_c5_run(w, 0, count, cores)
        return {"value": ext, "unit": UNIT, "cores": cores, "kind": "port", "replan_moves_per_sec": moves, "seconds": dt,
                "sample": f"{count} planners (global index 0..{count - 1}) of workload c5, full length ({w['iter_num']} iterations + "
                          f"{w['ticks']} replanning ticks) on {cores} host threads, C port of the reference algorithm + "
                          "oracle/fnd_loop.py"}
    count, iters = size_cpu_sample(w, budget_s, cores)
    ext, edge, dt, _ = cpu_port_run(w, 0, count, iters, cores)
    if dt < 0.4 * budget_s:        # the calibration over-estimated: scale the sample up to the budget
        count = int(min(16384, count * max(1.0, 0.8 * budget_s / max(dt, 1e-3))))
        ext, edge, dt, _ = cpu_port_run(w, 0, count, iters, cores)
    return {"value": ext, "unit": UNIT, "cores": cores, "kind": "port",
            "edge_checks_per_sec": edge, "seconds": dt,
            "sample": f"{count} planners (global index 0..{count - 1}) x {iters} of {w['iter_num']} iterations of workload "
                      f"{w['name']} on {cores} host threads, C port of the reference algorithm (oracle/rrt_oracle.c)"}


def bench_config(w, P, world, args, capacity, scaling="strong"):
    """The `config` object of the JSON line: identical for the B200 arm and the reference arm."""
    return {"workload": f"{w['name']}: {w['mode']} x {P * world} planners in total, {P} per GPU ({scaling} scaling), "
                        f"{len(w['obstacles'])} circular obstacles, {w['iter_num']} iterations each"
                        + (f", then {w['ticks']} replanning ticks with moving obstacles" if "ticks" in w else ""),
            "planners_per_gpu": P, "planners_total": P * world, "iter_num": w["iter_num"],
            "l2": f"per-step working set {P * capacity * 37 / 1e6:.0f} MB of tree state per GPU is re-initialised every step "
                  "(larger than L2 for the default size; no explicit flush)",
            "launch": args.launch, "near_cap": args.near_cap}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, _ = load_workload(args)
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    P, _, scaling = planner_layout(args, world, 0)
    cores = len(os.sched_getaffinity(0))
    # bounded sample per step, sized so that warmup + steps stay within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    if w["name"] == "c5":
        _, _, dt1, _ = cpu_c5_run(w, 0, cores, cores)
        count, iters = int(min(4096, max(cores, cores * (budget // max(dt1, 1e-3))))), w["iter_num"]

        def run_once():
            _, _, _, ext = cpu_c5_run(w, 0, count, cores)
            return int(ext), 0
    else:
        count, iters = size_cpu_sample(w, budget, cores)
        _, _, dt1, _ = cpu_port_run(w, 0, count, iters, cores)
        if dt1 < 0.4 * budget:     # the calibration over-estimated: scale the sample up to the budget
            count = int(min(16384, count * max(1.0, 0.8 * budget / max(dt1, 1e-3))))

        def run_once():
            _, _, _, out = cpu_port_run(w, 0, count, iters, cores)
            return int(out["iter"].sum()), int(out["n_edge_checks"].sum())
    for _ in range(args.warmup):
        run_once()
    t0 = time.perf_counter()
    ext_total, edge_total = 0, 0
    for _ in range(args.steps):
        a, e = run_once()
        ext_total += a; edge_total += e
    dt = time.perf_counter() - t0
    val = ext_total / dt
    sample = (f"{count} planners x {iters} of {w['iter_num']} iterations of workload {w['name']} per step on {cores} host "
              f"threads, C port of the reference algorithm (the Python reference cannot travel to the GPU box)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(w, P, world, args,
                                   (w["max_nodes"] + 2 + max(256, w["max_nodes"] // 4)) if w["mode"] == "fn" else w["iter_num"] + 2, scaling),
            "edge_checks_per_sec": edge_total / dt,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def planner_layout(args, world, rank):
    """(planners of this rank, global index of its first planner, "strong" | "weak")."""
    from rrt_star_fnd_algorithm_b200.parallel import shard_range
    total = DEFAULT_PLANNERS[args.workload]
    if args.planners:                                    # an explicit per-GPU batch: weak by definition
        return args.planners, rank * args.planners, "weak"
    if args.scaling == "strong" and total >= world and total % world == 0:
        first, count = shard_range(total, rank, world)
        return count, first, "strong"
    return total, rank * total, "weak"                   # replicas (c1 / c4 hold a single planner) or --scaling weak


def measure(args, w, P, first, dev, world, rank, steps, warmup, sample_clocks):
    """Build this rank's batch of P planners (global indices first ..), run `warmup` + `steps` steps, return the device-timed
    results: every rank returns the same max-over-ranks time and the all-reduced counters."""
    import torch
    import torch.distributed as dist
    from rrt_star_fnd_algorithm_b200 import workloads as W
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    from rrt_star_fnd_algorithm_b200.parallel import gather_results
    starts, goals, seeds, sids = W.planner_problems(w, first, P)
    kw = dict(obstacles=w["obstacles"], xy_range=w["xy_range"], iter_num=w["iter_num"], step_length=w["step_length"],
              radius=w["radius"], node_radius=w["node_radius"], bias=w["bias"], max_nodes=w["max_nodes"],
              near_cap=args.near_cap, launch=args.launch)
    if w.get("node_grid"):
        kw.update(node_grid=True, finish_cap=65536, path_cap=65536)
    b = PlannerBatch(w["mode"], starts=starts, goals=goals, seeds=seeds, stream_ids=sids, device=dev, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kern_ms = []
    is_c5 = w["name"] == "c5"
    tight = b.prm.capacity
    replan = {"ms": [], "stats": None}
    if is_c5:
        from rrt_star_fnd_algorithm_b200.fnd_loop import DeviceFNDReplanner, moved_obstacles
        tables = [moved_obstacles(w["obstacles"], w["velocity"], k) for k in range(1, w["ticks"] + 1)]

    def step(timed):
        if is_c5:                         # back to the planning state: tight capacity, obstacles where they started
            b.restore_problems()
            b.reallocate(tight)
            b.set_obstacles(w["obstacles"])
        b.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.run()
        e1.record()
        if is_c5:                         # replanning ticks: robot steps, obstacles move, invalidate, reconnect / regrow
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            b.grow_capacity(tight + w["ticks"] * w["regrow_iters"])
            loop = DeviceFNDReplanner(b, w["reconnect_radius"], w["regrow_iters"], edge_invalidation=w.get("edge_invalidation", False))
            for ob in tables:                 # one C-ABI call per tick; nothing comes back to the host inside the loop
                loop.tick(ob, fetch=False)
            r1.record()
            if timed:
                replan["ms"].append((r0, r1)); replan["stats"] = dict(loop.stats)
        if world > 1:
            gather_results(b, dist.group.WORLD)     # all_gather of the records + broadcast of the best path
        if timed:
            kern_ms.append((e0, e1))

    for _ in range(warmup):
        step(False)
    barrier()
    clocks = ClockSampler(dev.index) if (sample_clocks and rank == 0) else None
    if clocks:
        clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clk = clocks.stop() if clocks else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())

    pl = b.planners()
    b.check_status()
    if is_c5:                             # extensions of a step = the planning run + what regrow() added
        pl["iter"][0] += replan["stats"]["regrow_extensions"]
    stats = torch.tensor([pl["iter"].sum(), pl["n_edge_checks"].sum(), pl["n_edge_checks_ref"].sum(),
                          pl["n_circle_tests"].sum(), pl["attempts"].sum(), pl["n_scan_nodes"].sum(),
                          pl["n_chain_hops"].sum(), pl["solution_found"].sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
    return dict(b=b, kw=kw, pl=pl, ms=ms, clk=clk, stats=[float(v) for v in stats.tolist()],
                k_ms=float(np.mean([a.elapsed_time(c) for a, c in kern_ms])),
                replan_ms=float(np.mean([a.elapsed_time(c) for a, c in replan["ms"]])) if replan["ms"] else None,
                replan_stats=replan["stats"])


def main():
    args = parse_args()
    if os.environ.get("BENCH_DUMP_AFTER"):      # debugging aid: print every thread's Python stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_DUMP_AFTER"]), exit=True)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        import datetime
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version / debug lines go to stderr: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    w, _ = load_workload(args)
    P, first, scaling = planner_layout(args, world, rank)
    is_c5 = w["name"] == "c5"
    m = measure(args, w, P, first, dev, world, rank, args.steps, args.warmup, True)
    b, pl, ms, clk = m["b"], m["pl"], m["ms"], m["clk"]
    ext, edges, edges_ref, circles, attempts, scan_nodes, hops, solved = m["stats"]
    # the counters are totals of ONE step (state is re-initialised each step); the timed region holds `steps` of them
    rate = args.steps / (ms / 1e3)

    # ---- e2e through the host-buffer C ABI (pinned host memory, copies inside the timed region); EVERY rank takes
    # part (its own planners, collectives inside), so this must come before the non-zero ranks leave
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e_c5(args, w, P, first, dev, m["kw"]) if is_c5 else run_e2e(args, w, P, first, b, dev)

    # ---- rank 0: roofline of the fused planner kernel (its own launches; algorithmic bytes) and the
    # parity spot check, while the headline batch is still alive
    roofline = parity = None
    capacity = b.prm.capacity
    import ctypes as C
    from rrt_star_fnd_algorithm_b200 import _abi
    plan_launches = int(_abi.lib.rrtfnd_plan_launches(C.byref(b.prm), P))     # init_kernel + this many plan_kernel launches per step
    if rank == 0:
        k_ms = m["k_ms"]
        alg_bytes = 16.0 * pl["n_scan_nodes"].sum() + 12.0 * pl["n_chain_hops"].sum() + 32.0 * pl["iter"].sum()
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (k_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                # measured for the workload's default single-GPU launch only
                traffic = json.load(open(tpath)).get(f"{args.workload}") if P == DEFAULT_PLANNERS[args.workload] and not args.iters else None
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "plan_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": peak_src,
                    "note": "algorithmic bytes (16 B per node visited by a scan + 12 B per parent-chain hop + 32 B per insert) of rank "
                            "0's launch over its CUDA-event time; the scans are served from L2 through the TMA ring (DRAM traffic is "
                            "`traffic`), the kernel is bound by per-warp instruction latency, not by HBM (profiles/)"}
        parity = spot_check(w, first, b)      # against the CPU oracle (not timed): planner 0 of this rank

    # ---- the other scaling mode beside the headline (N > 1): a full workload-sized batch on every GPU
    weak = None
    if world > 1 and scaling == "strong":
        total, n = DEFAULT_PLANNERS[args.workload], max(1, min(args.steps, 3))
        del b
        m["b"] = None
        torch.cuda.empty_cache()
        mw = measure(args, w, total, rank * total, dev, world, rank, n, 1, False)
        wrate = n / (mw["ms"] / 1e3)
        weak = {"value": mw["stats"][0] * wrate, "unit": UNIT, "planners_per_gpu": total, "steps": n, "ms_per_step": mw["ms"] / n,
                "edge_checks_per_sec": mw["stats"][1] * wrate,
                "note": "the same kernels with a full workload-sized batch on EVERY GPU (total work grows with N); device-timed, "
                        "max over ranks"}
        mw["b"] = None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(w, args.cpu_seconds)

    line = {
        "metric": METRIC, "value": ext * rate,
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(w, P, world, args, capacity, scaling),
        "edge_checks_per_sec": edges * rate,
        "edge_checks_ref_equiv_per_sec": edges_ref * rate,
        "circle_tests_per_sec": circles * rate,
        "attempts_per_sec": attempts * rate,
        "solved_planners": solved, "gpu_launches": (1 + plan_launches) * args.steps,
        "roofline": roofline, "clocks": clk, "parity_spot_check": parity,
    }
    if weak is not None:
        line["weak_scaling"] = weak
    if is_c5:
        st, rms = m["replan_stats"], m["replan_ms"]
        line["replanning"] = dict(st, ticks=w["ticks"], ms_per_step=rms, plan_ms_per_step=m["k_ms"],
                                  moves_per_sec=st["moves"] / (rms / 1e3), note="rank 0; one step = RRT*-FN planning "
                                  "(plan_kernel) + `ticks` replanning ticks (rrtfnd_fnd_tick: select_branch, valid_path, edge cut, "
                                  "reconnect, regrow kernels with the bookkeeping between them on the device; the moved obstacle "
                                  "table goes up once per tick, nothing comes back)")
        line["gpu_launches"] = int(args.steps * (1 + plan_launches + w["ticks"] * 23))
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, w, P, first, b, dev):
    """Same workload through rrtfnd_plan_host: host inputs -> device -> run -> trees + results back to host."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from rrt_star_fnd_algorithm_b200 import _abi, workloads as W
    from rrt_star_fnd_algorithm_b200._structs import PLANNER_BYTES
    world = int(os.environ.get("WORLD_SIZE", "1"))
    starts, goals, seeds, sids = W.planner_problems(w, first, P)
    prm = b.prm
    Cn = prm.capacity
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    sg = pin((P, 4), torch.float64); sg.copy_(torch.from_numpy(np.concatenate([starts, goals], 1)))
    sd = pin((P, 2), torch.int64); sd.copy_(torch.from_numpy(np.stack([seeds, sids], 1).astype(np.int64)))
    rows = np.asarray(w["obstacles"], np.float64)          # (ox, oy, r) circles, or table rows (cx, cy, -hx, hy) of the rectangle variant
    ob = pin((max(prm.n_obst, 1), rows.shape[1]), torch.float64); ob[:prm.n_obst].copy_(torch.from_numpy(rows))
    if rows.shape[1] == 4:                                 # rows that carry their fourth column: rrtfnd_plan_host takes them as they are
        import copy
        from rrt_star_fnd_algorithm_b200._structs import FLAG_OBST_HAS_K
        prm = copy.copy(prm)
        prm.flags |= FLAG_OBST_HAS_K
    o_pl = pin((P, PLANNER_BYTES), torch.uint8)
    o_xy, o_e = pin((P, Cn, 2), torch.float64), pin((P, Cn), torch.float64)
    o_par, o_nc, o_id = pin((P, Cn), torch.int32), pin((P, Cn), torch.int32), pin((P, Cn), torch.int32)
    o_bp = pin((P, prm.path_cap), torch.int32)
    h2d = sg.numel() * 8 + sd.numel() * 8 + prm.n_obst * 8 * rows.shape[1]
    d2h = o_pl.numel() + o_xy.numel() * 8 + o_e.numel() * 8 + 3 * o_par.numel() * 4 + o_bp.numel() * 4

    def call():
        rc = _abi.lib.rrtfnd_plan_host(C.byref(prm), P, sg.data_ptr(), sd.data_ptr(), ob.data_ptr(), None,
                                       o_pl.data_ptr(), o_xy.data_ptr(), o_e.data_ptr(),
                                       o_par.data_ptr(), o_nc.data_ptr(), o_id.data_ptr(), o_bp.data_ptr(), dev.index)
        _abi.check(rc, "rrtfnd_plan_host")

    call()  # warm-up (allocator, module load)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        call()           # synchronous: returns after the device->host copies have completed
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    from rrt_star_fnd_algorithm_b200._structs import PLANNER_DTYPE
    pl = o_pl.numpy().reshape(-1).view(PLANNER_DTYPE)
    ext = torch.tensor([float(pl["iter"].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ext)
    return {"value": float(ext.item()) * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": n, "ms_per_step": dt / n * 1e3,
            "api": "rrtfnd_plan_host (C ABI, pinned host buffers; whole trees + planner records + best paths copied back)"}


def run_e2e_c5(args, w, P, first, dev, kw):
    """Config 5 end to end: host problem arrays -> a fresh device batch -> planning -> replanning ticks -> whole trees,
    records and paths back on the host; wall clock around all of it."""
    import torch
    import torch.distributed as dist
    from rrt_star_fnd_algorithm_b200 import workloads as W
    from rrt_star_fnd_algorithm_b200.batch import PlannerBatch
    from rrt_star_fnd_algorithm_b200.fnd_loop import DeviceFNDReplanner, moved_obstacles
    world = int(os.environ.get("WORLD_SIZE", "1"))
    starts, goals, seeds, sids = W.planner_problems(w, first, P)
    tables = [moved_obstacles(w["obstacles"], w["velocity"], k) for k in range(1, w["ticks"] + 1)]
    sizes, host = {}, {}

    def call():
        b = PlannerBatch(w["mode"], starts=starts, goals=goals, seeds=seeds, stream_ids=sids, device=dev, **kw).run()
        b.grow_capacity(b.prm.capacity + w["ticks"] * w["regrow_iters"])
        loop = DeviceFNDReplanner(b, w["reconnect_radius"], w["regrow_iters"], edge_invalidation=w.get("edge_invalidation", False))
        for ob in tables:
            loop.tick(ob, fetch=False)
        dev_t = (b.planners_t, b.xy, b.ecost, b.parent, b.nchild, b.id, b.best_path_t, loop.t["phase"])
        if not host:                          # pinned result buffers, allocated by the warm-up call
            host["bufs"] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in dev_t]
        for h, t in zip(host["bufs"], dev_t):
            h.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        out = host["bufs"]
        sizes["d2h"] = sum(t.numel() * t.element_size() for t in out)
        sizes["h2d"] = starts.nbytes + goals.nbytes + seeds.nbytes + sids.nbytes + (1 + len(tables)) * len(w["obstacles"]) * 32
        from rrt_star_fnd_algorithm_b200._structs import PLANNER_DTYPE
        return float(out[0].numpy().reshape(-1).view(PLANNER_DTYPE)["iter"].sum() + loop.stats["regrow_extensions"])

    call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    ext = 0.0
    for _ in range(n):
        ext = call()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt, ext], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(tt)
        dt, ext = float(mx[0].item()), float(tt[1].item())
    return {"value": ext * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(sizes["h2d"]), "d2h_bytes_per_step": int(sizes["d2h"]),
            "steps": n, "ms_per_step": dt / n * 1e3,
            "api": "PlannerBatch + DeviceFNDReplanner (Python host over the C ABI: rrtfnd_plan_batch, rrtfnd_fnd_tick): problems from host "
                   "arrays, one obstacle table per tick, whole trees + planner records + best paths + phases copied back into pinned "
                   "buffers"}


def spot_check(w, first, b):
    """Planner 0 of this rank vs the CPU oracle on the same Philox stream (bit-exact tree comparison)."""
    try:
        import oracle as O
        from rrt_star_fnd_algorithm_b200 import _structs as S
        from rrt_star_fnd_algorithm_b200 import workloads as W
        s, g, seeds, sids = W.planner_problems(w, first, 1)
        t = O.OracleTree(s[0], w["iter_num"] + 2 + w.get("ticks", 0) * w.get("regrow_iters", 0))
        t.run(oracle_params(w, 0), w["obstacles"], g[0], seed=int(seeds[0]), stream_id=int(sids[0]))
        if w["name"] == "c5":
            from oracle.fnd_loop import OracleReplanner
            from rrt_star_fnd_algorithm_b200.fnd_loop import moved_obstacles
            r = OracleReplanner(t, oracle_params(w, 0), g[0], seeds[0], sids[0], w["node_radius"], w["reconnect_radius"],
                                w["regrow_iters"], edge_invalidation=w.get("edge_invalidation", False))
            for k in range(1, w["ticks"] + 1):
                r.tick(moved_obstacles(w["obstacles"], w["velocity"], k))
        got, want = b.tree(0), t.export()
        ok = all(np.array_equal(got[k], want[k]) for k in want)
        ok = ok and float(b.planners()[0]["best_cost"]) == t.planner().best_cost
        return "bit-exact vs oracle (planner 0)" if ok else "MISMATCH vs oracle (planner 0)"
    except Exception as e:  # the bench number does not depend on the checker
        return f"not checked: {type(e).__name__}: {e}"


if __name__ == "__main__":
    main()
