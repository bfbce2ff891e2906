// arith.cuh — the fp64 arithmetic contract of the reference, as device functions.
//
// Every operation is an explicitly rounded IEEE fp64 intrinsic (__dadd_rn, __dmul_rn, __ddiv_rn,
// __dsqrt_rn), so nothing here can be contracted into an FMA whatever the compiler flags are; the one
// fused multiply-add of the contract (calc_distance) is written as __fma_rn.
// Reference lines are relative to the reference root (src/...).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace rrtfnd {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
// IEEE division and square root expand to long instruction sequences (reciprocal / rsqrt seed, Newton steps, exact
// rounding fix-up). They are calls, not inline expansions, so that each exists once per kernel: the fused planner
// has about twenty call sites and must fit the instruction cache (see profiles/).
static __device__ __noinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __noinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// two quotients over one divisor: the two Newton chains are independent and interleave inside one body
static __device__ __noinline__ double2 ddiv2(double n1, double n2, double d) { return make_double2(__ddiv_rn(n1, d), __ddiv_rn(n2, d)); }

// calc_distance (src/algorithm.py:954-961): np.linalg.norm(p - q) = sqrt(ddot(v, v)); OpenBLAS' ddot tail
// loop is FMA-contracted, i.e. sqrt(fma(dy, dy, fl(dx*dx))).
__device__ __forceinline__ double dist_fma(double px, double py, double qx, double qy) {
    const double dx = dsub(px, qx), dy = dsub(py, qy);
    return dsqrt(__fma_rn(dy, dy, dmul(dx, dx)));
}

// cKDTree squared distance: k-NN order (src/algorithm.py:684) and ball membership d2 <= r*r (:837, :893)
__device__ __forceinline__ double dist2(double nx, double ny, double qx, double qy) {
    const double dx = dsub(nx, qx), dy = dsub(ny, qy);
    return dadd(dmul(dx, dx), dmul(dy, dy));
}

// np.linalg.norm(in_radius_pos - q_new, axis=1) (src/algorithm.py:842, :899)
__device__ __forceinline__ double dist_plain(double nx, double ny, double qx, double qy) {
    return dsqrt(dist2(nx, ny, qx, qy));
}

// The reference's `x ** 2` / `x ** 0.5` sites (src/algorithm.py:602-603, :617, :619-620; src/map.py:44). Default: fl(x*x) and
// sqrt(x). A build with -DRRTFND_LIBM_POW=1 evaluates them as the libm pow() the interpreter calls (libm_pow.h, bit-equal
// to glibc 2.39's pow on 2 x 10^10 checked arguments): bit-exact to the reference by construction at ~60 fp64 operations
// and two table reads per site. Compile-time so that the default kernels are untouched; see DESIGN.md section 2.
#ifndef RRTFND_LIBM_POW
#define RRTFND_LIBM_POW 0
#endif
#if RRTFND_LIBM_POW
}  // namespace rrtfnd
#include "libm_pow.h"
namespace rrtfnd {
static __device__ __noinline__ double pow2_ref(double x) { return rrtfnd_libm_pow2(x); }
static __device__ __noinline__ double pow_half_ref(double x) { return rrtfnd_libm_pow_half(x); }
#else
__device__ __forceinline__ double pow2_ref(double x) { return dmul(x, x); }
__device__ __forceinline__ double pow_half_ref(double x) { return dsqrt(x); }
#endif

// Decide the circle tests without the square root and the two divisions whenever that is certain (root_filter.h: exact; the
// reference's expressions are evaluated when a root is within 2^-47, relative, of an end of the segment's x-interval - which
// planner segments, whose ends keep node_radius away from every circle, practically never are). The lanes of a warp test
// different circles, so before this any lane with delta >= 0 sent the whole warp through the root and the quotients.
// Measured on B200 (profiles/r02_s17_raw.txt): +3.8 % on config 3, +8.4 % on its 2,048-planner shard, config 2 unchanged.
#ifndef RRTFND_ROOT_FILTER
#define RRTFND_ROOT_FILTER 1
#endif
#if RRTFND_ROOT_FILTER
}  // namespace rrtfnd
#include "root_filter.h"
namespace rrtfnd {
#endif

// Line(p1, p2) in slope / intercept form (src/line.py:5-16) plus the per-segment terms of calc_delta
// (src/algorithm.py:607-621) that do not depend on the circle.
struct Seg {
    double p1x;      // x_pos of a vertical line
    double p1y, ylo, yhi;   // only read by the rectangle extension (values the callers hold anyway)
    double m, k;     // Line.dir, Line.const_term
    double mk, kk;   // m*k, k*k (k**2)
    double a2, a4;   // 2*a, 4*a with a = 1 + m*m   (exact scalings)
    double lo, hi;   // min / max of the endpoint x (is_between, :632-633)
    bool vertical;   // p2x == p1x (src/line.py:9)
};

__device__ __forceinline__ Seg make_seg(double p1x, double p1y, double p2x, double p2y) {
    Seg s;
    s.p1x = p1x;
    s.vertical = (p2x == p1x);
    s.m = ddiv(dsub(p2y, p1y), dsub(p2x, p1x));            // src/line.py:15
    s.k = dsub(p1y, dmul(s.m, p1x));                       // src/line.py:16 (uses p1)
    s.mk = dmul(s.m, s.k);
    s.kk = pow2_ref(s.k);                                  // line.const_term ** 2 (src/algorithm.py:619)
    const double a = dadd(1.0, pow2_ref(s.m));             // 1 + line.dir ** 2 (:617)
    s.a2 = dadd(a, a);
    s.a4 = dmul(4.0, a);
    s.lo = p1x < p2x ? p1x : p2x;
    s.hi = p1x < p2x ? p2x : p1x;
    s.p1y = p1y;
    s.ylo = p1y < p2y ? p1y : p2y;
    s.yhi = p1y < p2y ? p2y : p1y;
    return s;
}

// ---- EXTENSION: axis-aligned rectangles -------------------------------------------------------------------------------
// The reference declares Map.obstacles_r (src/map.py:18, :54) and never implements it, so there is nothing to be faithful
// to: PARITY UNPINNED BY THE REFERENCE, pinned on the oracle's restatement of the same expressions (oracle/rrt_oracle.c).
// A rectangle shares the obstacle table with the circles as the row (cx, cy, -hx, hy): centre, NEGATED half width (a
// negative third column marks a rectangle, hx > 0), half height. Semantics, chosen like the circles' (segments against
// the raw shape, points against the shape inflated by node_radius):
//   segment: closed segment p1-p2 intersects the closed rectangle, by separating axes on the segment's midpoint m and
//            half vector e - no division, one rounding per written operation:
//            |m.x - cx| <= hx + |e.x|  and  |m.y - cy| <= hy + |e.y|  and  |m' x e| <= hx |e.y| + hy |e.x|
//   point:   distance from the point to the rectangle < node_radius (strict, like src/map.py:45), or the point lies in
//            the closed rectangle.
static __device__ __noinline__ bool seg_hits_rect(double p1x, double p1y, double p2x, double p2y, double cx, double cy, double hx,
                                                  double hy) {
    const double ex = dmul(dsub(p2x, p1x), 0.5), ey = dmul(dsub(p2y, p1y), 0.5);
    const double mx = dsub(dadd(p1x, ex), cx), my = dsub(dadd(p1y, ey), cy);
    const double ax = fabs(ex), ay = fabs(ey);
    if (fabs(mx) > dadd(hx, ax)) return false;
    if (fabs(my) > dadd(hy, ay)) return false;
    return !(fabs(dsub(dmul(mx, ey), dmul(my, ex))) > dadd(dmul(hx, ay), dmul(hy, ax)));
}
static __device__ __noinline__ bool point_in_inflated_rect(double px, double py, double cx, double cy, double hx, double hy,
                                                           double node_radius) {
    const double dx = fmax(dsub(fabs(dsub(px, cx)), hx), 0.0), dy = fmax(dsub(fabs(dsub(py, cy)), hy), 0.0);
    if (dx == 0.0 && dy == 0.0) return true;
    return dsqrt(dadd(dmul(dx, dx), dmul(dy, dy))) < node_radius;
}

// intersection_circle (src/algorithm.py:555-578) for one circle (ox, oy, r) with
// K = ((-(r*r)) + ox*ox) + oy*oy precomputed (first three terms of `c`, :619).
// RECT = false compiles the rectangle dispatch out: measured on B200, the never-taken branch and the call behind it cost
// the fused planner 5.8 % on config 3 (registers held across the circle loop), so the batch kernels of circle-only tables
// are instantiated without it (RRTFND_FLAG_RECTANGLES selects the instances with it).
template <bool RECT = true>
__device__ __forceinline__ bool seg_hits_circle(const Seg& s, double ox, double oy, double r, double K) {
    if (RECT && r < 0.0)                                                        // a rectangle row (cx, cy, -hx, hy): extension above
        return seg_hits_rect(s.p1x, s.p1y, s.p1x == s.lo ? s.hi : s.lo, s.p1y == s.ylo ? s.yhi : s.ylo, ox, oy, -r, K);
    if (s.vertical) return fabs(dsub(ox, s.p1x)) <= r;                          // :563-567
    const double b = dmul(2.0, dsub(dadd(-ox, s.mk), dmul(s.m, oy)));           // :618
    const double c = dadd(dsub(K, dmul(dadd(oy, oy), s.k)), s.kk);              // :619
    const double delta = dsub(pow2_ref(b), dmul(s.a4, c));                      // b ** 2 - 4 * a * c (:620)
    if (!(delta >= 0.0)) return false;                                          // :569 (NaN -> no hit as well)
#if RRTFND_ROOT_FILTER
    {
        // the filter's lo * 2a and hi * 2a are loop invariants of the callers' circle loops; hoisted, they cost six registers
        // across the loop and the planner 5.9 % (with them recomputed per circle it GAINS 3.8 %): hide 2a from the optimiser
        double a2 = s.a2;
        asm volatile("" : "+d"(a2));
        const int f = rrtfnd_root_filter(b, delta, a2, s.lo, s.hi);
        if (f >= 0) return f != 0;
    }
#endif
    const double sq = pow_half_ref(delta);                                      // delta ** 0.5 (:602-603)
    const double2 x = ddiv2(dadd(-b, sq), dsub(-b, sq), s.a2);                  // x1, x2 (:602-603)
    return (x.x > s.lo && x.x < s.hi) || (x.y > s.lo && x.y < s.hi);                // :576, :632-635
}

// Map.is_occupied_c for one circle (src/map.py:44-45)
template <bool RECT = true>
__device__ __forceinline__ bool point_in_inflated(double px, double py, double ox, double oy, double r,
                                                  double node_radius, double K = 0.0) {
    if (RECT && r < 0.0) return point_in_inflated_rect(px, py, ox, oy, -r, K, node_radius);   // rectangle row (extension)
    const double dx = dsub(px, ox), dy = dsub(py, oy);
    return dsqrt(dadd(pow2_ref(dx), pow2_ref(dy))) < dadd(node_radius, r);      // np.sqrt((..) ** 2 + (..) ** 2)
}

// lexicographic (d2, slot) order: nearest first, lowest slot (= lowest id) on ties
__device__ __forceinline__ bool key_less(double d1, int s1, double d2, int s2) {
    return d1 < d2 || (d1 == d2 && s1 < s2);
}

}  // namespace rrtfnd
