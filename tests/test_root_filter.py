"""csrc/root_filter.h: the circle test's `is a root strictly inside the segment's x-interval`, decided without square root
and division whenever that is certain. One source for the device (arith.cuh)
and for this host build (oracle/root_filter_check.c, TEST INFRASTRUCTURE). The filter may decline (-1) but must never
disagree with the reference's expressions (src/algorithm.py:569-576, :594-604, :624-635)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", ORACLE, "-s", "rootcheck"])
    L = C.CDLL(os.path.join(ORACLE, "_build", "libroot_filter.so"))
    d = C.c_double
    L.orc_root_filter.argtypes = L.orc_root_slow.argtypes = [d, d, d, d, d]
    L.orc_root_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def quadratic(L, seg, circle):
    s, c, out = np.array(seg, np.float64), np.array(circle, np.float64), np.zeros(5)
    ok = L.orc_root_setup(s.ctypes.data, c.ctypes.data, out.ctypes.data)
    return ok, out


def test_filter_agrees_with_the_expressions_on_hand_made_cases(lib):
    circle = (10.0, 10.0, 2.0)
    cases = {                                   # segment -> reference's answer
        (0.0, 10.0, 20.0, 10.5): 1,             # straight through
        (0.0, 10.0, 7.0, 10.2): 0,              # stops short: both roots beyond hi
        (13.0, 10.1, 20.0, 10.3): 0,            # starts behind: both roots below lo
        (9.0, 9.9, 20.0, 10.3): 1,              # starts inside: one root in the interval
        (9.5, 9.8, 10.5, 10.1): 0,              # wholly inside the circle: x-only test finds no root in the interval
    }
    for seg, want in cases.items():
        ok, (b, delta, a2, lo, hi) = quadratic(lib, seg, circle)
        assert ok
        assert lib.orc_root_slow(b, delta, a2, lo, hi) == want
        assert lib.orc_root_filter(b, delta, a2, lo, hi) == want          # far from every threshold: decided, and right
    # an end of the segment exactly on the circle: the root meets lo - too close to call
    ok, (b, delta, a2, lo, hi) = quadratic(lib, (12.0, 10.0, 20.0, 10.0), circle)
    assert ok and lib.orc_root_filter(b, delta, a2, lo, hi) == -1
    # NaN / infinity / subnormal scale: declined
    assert lib.orc_root_filter(float("nan"), 1.0, 2.0, 0.0, 1.0) == -1
    assert lib.orc_root_filter(1.0, float("inf"), 2.0, 0.0, 1.0) in (0, -1)
    assert lib.orc_root_filter(1e-200, 1e-300, 2.0, 1e-200, 2e-200) == -1


def test_random_and_borderline_segments(lib):
    rng = np.random.default_rng(3)
    decided = total = 0
    for k in range(20000):
        ox, oy, r = rng.uniform(0, 200, 2).tolist() + [rng.uniform(1, 4)]
        p1 = rng.uniform(0, 200, 2)
        p2 = np.array([ox, oy]) + rng.uniform(-1.5, 1.5, 2) * r if k % 2 else rng.uniform(0, 200, 2)
        if k % 5 == 0:                          # an end on the circle, nudged by a few ulps
            th = rng.uniform(0, 2 * np.pi)
            p1 = np.array([ox + r * np.cos(th), oy + r * np.sin(th)])
            for _ in range(int(rng.integers(0, 4))):
                p1 = np.nextafter(p1, p1 + rng.choice([-1.0, 1.0], 2))
        ok, (b, delta, a2, lo, hi) = quadratic(lib, (p1[0], p1[1], p2[0], p2[1]), (ox, oy, r))
        if not ok:
            continue
        total += 1
        f = lib.orc_root_filter(b, delta, a2, lo, hi)
        if f >= 0:
            decided += 1
            assert f == lib.orc_root_slow(b, delta, a2, lo, hi), (p1, p2, ox, oy, r)
    assert decided > 0.7 * total and decided < total       # most are decided, the nudged ones are not


def test_recorded_through_obstacle_calls_of_the_reference(lib):
    """tests/golden/collision.npz: segments the UNMODIFIED reference passed to through_obstacle while planning, with its
    answers. Evaluated the way the device does (filter first, the plain expressions only when it declines) every answer is
    the reference's - and on planner segments, whose ends keep node_radius away from the circles, the filter practically never
    declines."""
    from oracle import cases as K
    from util import load_npz
    g = load_npz("collision.npz")
    names = sorted(k[:-4] for k in g if k.endswith("_seg"))      # synth_*: adversarial segments on the same maps
    assert len(names) >= 12
    lib.orc_root_through_obstacle.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    tests = declined = 0
    for name in names:
        case = K.case_by_name(name.replace("synth_", ""))
        circles = np.ascontiguousarray([[float(x), float(y), float(r)] for x, y, r in case["obstacles"]], dtype=np.float64)
        segs = np.ascontiguousarray(g[name + "_seg"], dtype=np.float64)
        hit = np.zeros(len(segs), np.uint8)
        counts = np.zeros(2, np.int64)
        lib.orc_root_through_obstacle(segs.ctypes.data, len(segs), circles.ctypes.data, len(circles), hit.ctypes.data, counts.ctypes.data)
        assert np.array_equal(hit, g[name + "_hit"].astype(np.uint8)), name
        tests += int(counts[0]); declined += int(counts[1])
    assert tests > 50_000 and declined <= tests // 1000, (tests, declined)      # measured: 60,131 tests, none declined


def test_bulk_checker_finds_no_mismatch(lib):
    """oracle/_build/root_filter_check: six families of generated input, ~12 million cases here (hundreds of millions when run
    by hand, see profiles/); exit status 1 on any disagreement."""
    out = subprocess.run([os.path.join(ORACLE, "_build", "root_filter_check"), "2", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("mismatches 0") == 6
