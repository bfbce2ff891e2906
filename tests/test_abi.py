"""CPU: the C-ABI library loads without a GPU and exports exactly what include/rrtfnd.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rrtfnd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rrtfnd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rrt_star_fnd_algorithm_b200 import _abi
    from rrt_star_fnd_algorithm_b200._structs import ABI_VERSION as S_ABI
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_abi.lib, n), f"{n} is declared in include/rrtfnd.h but missing from librrtfnd.so"
    assert sorted(_abi.EXPORTS) == names, "the ctypes binding list and the header disagree"
    assert _abi.lib.rrtfnd_abi_version() == S_ABI


def test_struct_mirrors_match_the_header_layout():
    """sizeof / field order of the ctypes mirrors against a C compile of the header."""
    import subprocess, tempfile
    from rrt_star_fnd_algorithm_b200 import _structs as S
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "rrtfnd.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu\n", sizeof(rrtfnd_params_t), sizeof(rrtfnd_planner_t), sizeof(rrtfnd_trace_t),
               sizeof(rrtfnd_batch_t), sizeof(rrtfnd_obst_grid_t));
        printf("%zu %zu %zu %zu\n", sizeof(rrtfnd_fnd_loop_t), offsetof(rrtfnd_fnd_loop_t, stats), offsetof(rrtfnd_batch_t, walk),
               sizeof(rrtfnd_walk_t));
        printf("%zu %zu %zu %zu\n", offsetof(rrtfnd_params_t, step_length), offsetof(rrtfnd_planner_t, n_slots),
               offsetof(rrtfnd_batch_t, grid), offsetof(rrtfnd_obst_grid_t, seg_start));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True).split()
    sizes = [int(v) for v in out]
    assert sizes[5:9] == [C.sizeof(S.FndLoop), S.FndLoop.stats.offset, S.Batch.walk.offset, 16]
    sizes = sizes[:5] + sizes[9:]
    assert sizes[:5] == [C.sizeof(S.Params), C.sizeof(S.Planner), C.sizeof(S.Trace), C.sizeof(S.Batch), C.sizeof(S.ObstGrid)]
    assert sizes[5:] == [S.Params.step_length.offset, S.Planner.n_slots.offset, S.Batch.grid.offset, S.ObstGrid.seg_start.offset]


def test_status_strings_and_bad_arguments_without_a_gpu():
    from rrt_star_fnd_algorithm_b200 import _abi, _structs as S
    assert _abi.status_string(S.ST_DONE) == "done"
    assert _abi.status_string(-100) == "bad argument"
    # NULL arguments are rejected before any CUDA call is made
    assert _abi.lib.rrtfnd_plan_batch(None, None, None) == -100
    assert _abi.lib.rrtfnd_edges_vs_circles(None, None, None, None, 5, None, 0, None, None) == -100
    assert _abi.lib.rrtfnd_edges_vs_circles(None, None, None, None, 0, None, 0, None, None) == 0   # empty input is fine


def test_obstacle_grid_builder_lists_every_circle_near_a_cell():
    """Host builder: every circle whose inflated box touches a cell is in that cell's list (brute-force check)."""
    from rrt_star_fnd_algorithm_b200.batch import build_obstacle_grid
    rng = np.random.default_rng(0)
    ob = np.column_stack([rng.uniform(0, 200, 300), rng.uniform(0, 200, 300), rng.uniform(0.5, 6, 300)])
    g, (ss, si, ps, pi) = build_obstacle_grid(ob, 2.0, (0.0, 200.0, 0.0, 200.0))
    assert g.nx > 1 and g.ny > 1 and ss[-1] == len(si) and ps[-1] == len(pi)
    cell = 1.0 / g.inv_cell
    assert g.margin >= 128 * 257 * 200 / 2 ** 26
    for which, start, items, infl in (("seg", ss, si, g.margin), ("pt", ps, pi, 2.0)):
        for cy in range(g.ny):
            for cx in range(g.nx):
                x0, x1, y0, y1 = cx * cell, (cx + 1) * cell, cy * cell, (cy + 1) * cell
                R = ob[:, 2] + infl
                # circles whose inflated box strictly overlaps the open cell must be listed
                must = (ob[:, 0] + R > x0) & (ob[:, 0] - R < x1) & (ob[:, 1] + R > y0) & (ob[:, 1] - R < y1)
                listed = set(items[start[cy * g.nx + cx]:start[cy * g.nx + cx + 1]].tolist())
                missing = set(np.flatnonzero(must).tolist()) - listed
                assert not missing, (which, cx, cy, sorted(missing)[:5])
    # no obstacles -> culling disabled
    g0, arrs0 = build_obstacle_grid(np.zeros((0, 3)), 1.0, (0.0, 1.0, 0.0, 1.0))
    assert g0.nx == 0 and arrs0 is None


def test_planner_kernel_sass_streams_the_tree_through_tma():
    """The built library's planner kernel holds the bulk-copy (TMA) and mbarrier instructions of the xy ring
    (csrc/stream.cuh): UBLKCP + SYNCS in the SASS of the lean RRT* instance, compiled for sm_100a."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    from rrt_star_fnd_algorithm_b200 import _abi
    out = subprocess.run([tool, "-sass", "-fun", "_ZN6rrtfnd11plan_kernelILi1ELi16ELb0ELi1ELb1ELb0ELb0ELb0EEEv13rrtfnd_params12rrtfnd_batch",
                          _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "SYNCS.ARRIVE.TRANS64" in out and "SYNCS.PHASECHK" in out
    assert out.count("DADD") > 100 and out.count("DMUL") > 50     # fp64 arithmetic as separately rounded adds and multiplies


def test_plan_launches_reports_the_slicing_rule():
    """rrtfnd_plan_launches: batches larger than the resident warp slots (148 SMs x 16 warps when no device answers) are grown in 8
    overlapped slices by the lean instances; everything else is one launch."""
    from rrt_star_fnd_algorithm_b200 import _abi, _structs as S
    p = S.Params()
    p.mode, p.iter_num, p.capacity, p.near_cap, p.finish_cap, p.path_cap = S.MODE_STAR, 5000, 5002, 64, 1024, 1024
    p.flags = S.FLAG_EVAL_CHOOSE_PARENT
    assert _abi.lib.rrtfnd_plan_launches(C.byref(p), 16384) == 8 and _abi.lib.rrtfnd_plan_launches(C.byref(p), 4096) == 8
    assert _abi.lib.rrtfnd_plan_launches(C.byref(p), 2048) == 1
    p.flags = S.FLAG_EVAL_CHOOSE_PARENT | S.FLAG_TRACE
    assert _abi.lib.rrtfnd_plan_launches(C.byref(p), 16384) == 1          # the generic instance (trace) is never sliced
    p.flags, p.mode = S.FLAG_EVAL_CHOOSE_PARENT, S.MODE_RRT
    assert _abi.lib.rrtfnd_plan_launches(C.byref(p), 16384) == 1          # RRT stops at the goal: no resumable slices
    assert _abi.lib.rrtfnd_plan_launches(None, 5) == -100
