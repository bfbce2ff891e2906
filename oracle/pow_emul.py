"""csrc/libm_pow.h compiled for the host (gcc) — TEST INFRASTRUCTURE: the same source the device operator
rrtfnd_pow_libm compiles, callable from Python, plus the bulk comparison against this machine's libm pow()."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpow_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "powcheck"])
        L = C.CDLL(_SO)
        L.orc_pow_emul_array.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
        L.orc_pow_emul_check.argtypes = [C.c_longlong, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def pow_emul(x, half=False) -> np.ndarray:
    """x ** 0.5 (half) or x ** 2 through the emulation, elementwise."""
    a = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(a)
    lib().orc_pow_emul_array(a.ctypes.data, a.size, int(bool(half)), out.ctypes.data)
    return out


def check(n, threads=4, seed=1):
    """Compare with libm pow() on n generated arguments: returns (pow2 mismatches, pow-half mismatches, how often pow(x, 2)
    != x*x, how often pow(x, 0.5) != sqrt(x), first bad argument)."""
    out = np.zeros(4, np.int64)
    fb = C.c_double(0.0)
    lib().orc_pow_emul_check(int(n), int(threads), int(seed), out.ctypes.data, C.byref(fb))
    return int(out[0]), int(out[1]), int(out[2]), int(out[3]), fb.value
