"""`x ** 2` / `x ** 0.5` as the reference's interpreter evaluates them (libm pow): csrc/libm_pow.h, one
source for the host checker (gcc) and the device operator rrtfnd_pow_libm (nvcc), against Python's own `**`."""
import platform

import numpy as np
import pytest

from oracle import pow_emul as E

GLIBC_239 = platform.libc_ver() == ("glibc", "2.39") and platform.machine() == "x86_64"
needs_this_libm = pytest.mark.skipif(not GLIBC_239, reason="the emulation restates glibc 2.39's x86-64 FMA pow")


def arguments(n, seed=0):
    rng = np.random.default_rng(seed)
    parts = [rng.uniform(-300, 300, n), rng.uniform(0, 1e6, n), rng.lognormal(0, 6, n), -rng.lognormal(0, 3, n // 4),
             1.0 + rng.uniform(-1e-9, 1e-9, n // 4), np.ldexp(1.0 + rng.uniform(0, 1e-12, n // 4), rng.integers(-60, 60, n // 4)),
             np.array([0.0, -0.0, 1.0, -1.0, 2.0, 0.5, 4.0, 1e-300, 1e300, 5e-324, np.inf, -np.inf, np.nan]),
             rng.uniform(0, 2.2e-308, 200)]                          # subnormals
    return np.concatenate(parts)


def python_pow(x, half):
    with np.errstate(all="ignore"):
        out = []
        for v in x.tolist():
            try:
                r = v ** (0.5 if half else 2)
            except OverflowError:                                  # float ** raises where libm returns inf
                r = float("inf")
            out.append(r if not isinstance(r, complex) else float("nan"))   # negative ** 0.5 is complex in Python
        return np.array(out, dtype=np.float64)


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    return (a.view(np.uint64) == b.view(np.uint64)) | both_nan


@needs_this_libm
def test_host_build_of_the_emulation_equals_python_pow():
    x = arguments(60000)
    sq, want = E.pow_emul(x), python_pow(x, False)
    assert same_bits(sq, want).all(), x[~same_bits(sq, want)][:5]
    assert (sq != x * x).sum() > 20                                 # ... and that is not x*x (the point of the exercise)
    xp = x[(x >= 0) | np.isnan(x)]             # includes -0.0 (pow gives +0, sqrt would give -0)
    rt, want = E.pow_emul(xp, half=True), python_pow(xp, True)
    assert same_bits(rt, want).all(), xp[~same_bits(rt, want)][:5]
    assert (rt != np.sqrt(xp)).sum() > 20
    # numpy float64 scalars take the same libm call
    for v in x[:2000]:
        assert same_bits(E.pow_emul([v])[0], float(np.float64(v) ** 2))


@needs_this_libm
def test_bulk_comparison_with_libm():
    bad2, badh, diff2, diffh, first = E.check(20_000_000, threads=4)
    assert (bad2, badh) == (0, 0), first
    assert diff2 > 5000 and diffh > 5000


@pytest.mark.gpu
def test_device_operator_equals_the_host_build_and_python_pow():
    import ctypes as C
    import torch
    from rrt_star_fnd_algorithm_b200 import _abi
    x = arguments(400000, seed=3)
    dev = torch.device("cuda", torch.cuda.current_device())
    xd = torch.from_numpy(x).to(dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for half in (False, True):
        xs = x if not half else x[(x >= 0) | np.isnan(x)]
        xd = torch.from_numpy(xs).to(dev)
        out = torch.empty_like(xd)
        _abi.check(_abi.lib.rrtfnd_pow_libm(C.c_void_p(xd.data_ptr()), xd.numel(), int(half), C.c_void_p(out.data_ptr()), st), "rrtfnd_pow_libm")
        got = out.cpu().numpy()
        want = E.pow_emul(xs, half=half)                           # the same source compiled by gcc
        assert same_bits(got, want).all(), xs[~same_bits(got, want)][:5]
        if GLIBC_239:                                               # and the interpreter of this very machine
            sub = slice(0, 50000)
            assert same_bits(got[sub], python_pow(xs[sub], half)).all()
    assert _abi.lib.rrtfnd_pow_libm(None, 5, 0, None, st) == -100   # RRTFND_E_BADARG
