"""ctypes mirrors of the structs in include/rrtfnd.h (kept field-for-field in the same order)."""
import ctypes as C

ABI_VERSION = 3

MODE_RRT, MODE_STAR, MODE_FN = 0, 1, 2
STREAM_WORDS, STREAM_PHILOX = 0, 1
FLAG_EVAL_CHOOSE_PARENT, FLAG_COMMIT_CHOOSE_PARENT, FLAG_TRACE, FLAG_NODE_GRID, FLAG_OBST_HAS_K, FLAG_RECTANGLES = 1, 2, 4, 8, 16, 32

ST_RUNNING, ST_DONE, ST_NEED_WORDS, ST_ATTEMPT_CAP = 0, 1, 2, 3
ST_CAPACITY, ST_NEAR_CAP, ST_FINISH_CAP, ST_PATH_CAP, ST_DUPLICATE, ST_NO_CHILDLESS, ST_STREAM_EXHAUSTED, ST_REGROW_NEED_WORDS = -1, -2, -3, -4, -5, -6, -7, -8

STATUS_NAMES = {
    ST_RUNNING: "running", ST_DONE: "done", ST_NEED_WORDS: "need-words", ST_ATTEMPT_CAP: "attempt-cap",
    ST_CAPACITY: "node capacity exhausted", ST_NEAR_CAP: "more nodes rewired in one iteration than near_cap",
    ST_FINISH_CAP: "finish list full", ST_PATH_CAP: "best path longer than path_cap",
    ST_DUPLICATE: "steered point coincides with an existing vertex",
    ST_NO_CHILDLESS: "forced_removal has no removable childless node",
    ST_STREAM_EXHAUSTED: "random word buffer exhausted inside forced_removal",
    ST_REGROW_NEED_WORDS: "regrow: random word buffer exhausted, refill and call again",
}

RESERVE_WORDS = 64  # words that must remain when an attempt starts in STREAM_WORDS mode


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("iter_num", C.c_int32), ("max_nodes", C.c_int32), ("capacity", C.c_int32),
        ("n_obst", C.c_int32), ("near_cap", C.c_int32), ("finish_cap", C.c_int32), ("path_cap", C.c_int32),
        ("stream_mode", C.c_int32), ("flags", C.c_int32), ("trace_cap", C.c_int32), ("reserved", C.c_int32),
        ("words_per_planner", C.c_int64), ("max_attempts", C.c_int64),
        ("step_length", C.c_double), ("radius", C.c_double), ("node_radius", C.c_double), ("bias", C.c_double),
        ("x_lo", C.c_double), ("x_hi", C.c_double), ("y_lo", C.c_double), ("y_hi", C.c_double),
        ("goal_node_radius", C.c_double),
    ]


class Planner(C.Structure):
    _fields_ = [
        ("start_x", C.c_double), ("start_y", C.c_double), ("goal_x", C.c_double), ("goal_y", C.c_double),
        ("best_cost", C.c_double),
        ("seed", C.c_uint64), ("stream_id", C.c_uint64), ("cursor", C.c_uint64),
        ("attempts", C.c_int64), ("n_edge_checks", C.c_int64), ("n_edge_checks_ref", C.c_int64),
        ("n_circle_tests", C.c_int64), ("n_scan_nodes", C.c_int64), ("n_chain_hops", C.c_int64),
        ("n_slots", C.c_int32), ("n_live", C.c_int32), ("last_id", C.c_int32), ("iter", C.c_int32),
        ("root_slot", C.c_int32), ("n_finish", C.c_int32), ("best_len", C.c_int32), ("status", C.c_int32),
        ("n_rewired", C.c_int32), ("n_removed", C.c_int32), ("solution_found", C.c_int32),
        ("n_compactions", C.c_int32), ("fn_count", C.c_int32), ("reserved", C.c_int32),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("id_new", C.c_int32), ("id_near", C.c_int32), ("n_near", C.c_int32), ("n_rewired", C.c_int32),
        ("removed_id", C.c_int32), ("cp_parent", C.c_int32), ("cp_cost", C.c_double),
    ]


class ObstGrid(C.Structure):
    _fields_ = [
        ("x0", C.c_double), ("y0", C.c_double), ("inv_cell", C.c_double),
        ("slope_max", C.c_double), ("coord_max", C.c_double), ("margin", C.c_double),
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("seg_start", C.c_void_p), ("seg_items", C.c_void_p), ("pt_start", C.c_void_p), ("pt_items", C.c_void_p),
    ]


class NodeGrid(C.Structure):
    _fields_ = [
        ("x0", C.c_double), ("y0", C.c_double), ("inv_cell", C.c_double), ("cell", C.c_double),
        ("nx", C.c_int32), ("ny", C.c_int32), ("min_nodes", C.c_int32), ("cand_cap", C.c_int32),
        ("head", C.c_void_p), ("next", C.c_void_p), ("cand", C.c_void_p),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_planners", C.c_int32), ("reserved", C.c_int32),
        ("planners", C.c_void_p),
        ("xy", C.c_void_p), ("ecost", C.c_void_p),
        ("parent", C.c_void_p), ("nchild", C.c_void_p), ("id", C.c_void_p), ("nflags", C.c_void_p),
        ("finish", C.c_void_p), ("best_path", C.c_void_p),
        ("obst", C.c_void_p), ("words", C.c_void_p), ("trace", C.c_void_p),
        ("grid", ObstGrid), ("node_grid", NodeGrid),
        ("walk", C.c_void_p),
    ]


class FndLoop(C.Structure):
    _fields_ = [
        ("max_rounds", C.c_int32), ("reserved", C.c_int32),
        ("phase", C.c_void_p), ("current", C.c_void_p), ("status", C.c_void_p), ("prev_root", C.c_void_p),
        ("separate", C.c_void_p), ("index_error", C.c_void_p), ("edge_cuts", C.c_void_p), ("todo", C.c_void_p),
        ("regrow_arg", C.c_void_p), ("link_tmp", C.c_void_p), ("regrow_tmp", C.c_void_p), ("iters_tmp", C.c_void_p),
        ("linked", C.c_void_p), ("regrow", C.c_void_p), ("regrow_iters", C.c_void_p), ("stats", C.c_void_p),
    ]


PLANNER_BYTES = C.sizeof(Planner)
TRACE_BYTES = C.sizeof(Trace)

# numpy structured dtypes with the same layout (for zero-copy views of downloaded buffers)
import numpy as _np  # noqa: E402

PLANNER_DTYPE = _np.dtype([(n, _np.dtype(t)) for n, t in Planner._fields_], align=True)
TRACE_DTYPE = _np.dtype([(n, _np.dtype(t)) for n, t in Trace._fields_], align=True)
assert PLANNER_DTYPE.itemsize == PLANNER_BYTES, (PLANNER_DTYPE.itemsize, PLANNER_BYTES)
assert TRACE_DTYPE.itemsize == TRACE_BYTES
