"""ctypes binding of librrtfnd.so (include/rrtfnd.h). There is no CPU fallback: if the library is missing
or cannot be loaded, importing this module raises."""
import ctypes as C
import os

from ._structs import ABI_VERSION, Batch, FndLoop, ObstGrid, Params, Planner

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRTFND_LIB") or os.path.join(_HERE, "librrtfnd.so")   # RRTFND_LIB: an A/B build of the same sources

# every symbol include/rrtfnd.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "rrtfnd_abi_version", "rrtfnd_status_string", "rrtfnd_plan_smem_bytes", "rrtfnd_obst_grid_build_host",
    "rrtfnd_edges_vs_circles", "rrtfnd_points_occupied", "rrtfnd_edges_vs_circles_grid", "rrtfnd_points_occupied_grid", "rrtfnd_nearest_visible", "rrtfnd_near_set",
    "rrtfnd_chain_cost", "rrtfnd_pow_libm", "rrtfnd_plan_batch", "rrtfnd_init_batch",
    "rrtfnd_fnd_select_branch", "rrtfnd_fnd_valid_path", "rrtfnd_fnd_reconnect", "rrtfnd_fnd_regrow",
    "rrtfnd_plan_host", "rrtfnd_obst_grid_build_host4", "rrtfnd_obst_grid_build_device", "rrtfnd_fnd_cut_blocked_edges",
    "rrtfnd_fnd_tick", "rrtfnd_plan_launches", "rrtfnd_host_release",
]


class LibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    PP, PB = C.POINTER(Params), C.POINTER(Batch)
    lib.rrtfnd_abi_version.restype = C.c_int
    lib.rrtfnd_status_string.restype = C.c_char_p
    lib.rrtfnd_status_string.argtypes = [C.c_int]
    lib.rrtfnd_plan_smem_bytes.restype = i64
    lib.rrtfnd_plan_smem_bytes.argtypes = [PP]
    lib.rrtfnd_plan_launches.argtypes = [PP, i32]
    lib.rrtfnd_edges_vs_circles.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, vp]
    lib.rrtfnd_points_occupied.argtypes = [vp, vp, i64, vp, i32, dbl, vp, vp]
    lib.rrtfnd_edges_vs_circles_grid.argtypes = [vp, vp, vp, vp, i64, vp, i32, C.POINTER(ObstGrid), vp, vp]
    lib.rrtfnd_points_occupied_grid.argtypes = [vp, vp, i64, vp, i32, C.POINTER(ObstGrid), dbl, vp, vp]
    lib.rrtfnd_nearest_visible.argtypes = [PB, i32, vp, i32, vp, vp, vp]
    lib.rrtfnd_near_set.argtypes = [PB, i32, vp, dbl, i32, vp, vp, vp]
    lib.rrtfnd_chain_cost.argtypes = [PB, i32, vp, vp, i64, vp, vp]
    lib.rrtfnd_pow_libm.argtypes = [vp, i64, i32, vp, vp]
    lib.rrtfnd_plan_batch.argtypes = [PP, PB, vp]
    lib.rrtfnd_init_batch.argtypes = [PP, PB, vp]
    lib.rrtfnd_fnd_select_branch.argtypes = [PP, PB, vp, vp, vp]
    lib.rrtfnd_fnd_valid_path.argtypes = [PP, PB, vp, vp, vp, vp]
    lib.rrtfnd_fnd_reconnect.argtypes = [PP, PB, vp, dbl, vp, vp]
    lib.rrtfnd_fnd_regrow.argtypes = [PP, PB, vp, i32, vp, vp, vp]
    lib.rrtfnd_plan_host.argtypes = [PP, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]
    lib.rrtfnd_obst_grid_build_host.argtypes = [vp, i32, dbl, dbl, dbl, dbl, dbl, dbl, C.POINTER(ObstGrid), vp, vp, vp, vp, vp]
    lib.rrtfnd_obst_grid_build_host4.argtypes = [vp, i32, dbl, dbl, dbl, dbl, dbl, dbl, C.POINTER(ObstGrid), vp, vp, vp, vp, vp]
    lib.rrtfnd_obst_grid_build_device.argtypes = [vp, i32, dbl, dbl, dbl, dbl, dbl, dbl, dbl, C.POINTER(ObstGrid), i32, i32, i32, vp, vp]
    lib.rrtfnd_fnd_cut_blocked_edges.argtypes = [PP, PB, vp, vp, vp, vp]
    lib.rrtfnd_fnd_tick.argtypes = [PP, PB, C.POINTER(FndLoop), dbl, i32, i32, vp]
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError here = the library does not match the header
        if name not in ("rrtfnd_status_string", "rrtfnd_plan_smem_bytes", "rrtfnd_abi_version"):
            getattr(lib, name).restype = C.c_int
    v = lib.rrtfnd_abi_version()
    if v != ABI_VERSION:
        raise LibraryError(f"librrtfnd.so ABI version {v}, expected {ABI_VERSION}")
    return lib


lib = _load()


def status_string(code: int) -> str:
    return lib.rrtfnd_status_string(int(code)).decode()


def check(rc: int, what: str):
    """Raise on a non-zero return code of a C-ABI call (cudaError_t > 0, library error < 0)."""
    if rc == 0:
        return
    if rc > 0:
        raise LibraryError(f"{what}: CUDA error {rc}")
    raise LibraryError(f"{what}: {status_string(rc)} ({rc})")
