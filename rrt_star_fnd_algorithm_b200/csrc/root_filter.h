/* root_filter.h — decide `is a root of the line/circle quadratic strictly inside the segment's x-interval` without the square
 * root and the two divisions, whenever the decision is certain. One source for nvcc (arith.cuh) and gcc (the checker
 * oracle/root_filter_check.c), every operation an explicit single-rounding add / mul.
 *
 * The reference (src/algorithm.py:569-576, :594-604, :624-635) computes, in floating point,
 *     sq = delta ** 0.5;  x1 = (-b + sq) / (2a);  x2 = (-b - sq) / (2a);  hit = lo < x1 < hi  or  lo < x2 < hi
 * With L = lo * 2a + b and H = hi * 2a + b (L < H because lo < hi and 2a >= 2) the four comparisons are, in real arithmetic,
 *     x1 > lo <=> sq > L      x1 < hi <=> sq < H      x2 > lo <=> sq < -L      x2 < hi <=> sq > -H
 * so, sq being >= 0:   L, H of opposite sign:  hit <=> sq < max(|L|, |H|)
 *                      L, H of the same sign:  hit <=> min(|L|, |H|) < sq < max(|L|, |H|)
 * and sq can be compared through its square, delta. The floating-point pipeline above moves each threshold by at most
 * u * sq (rounding of the root) + 2 u * |lo * 2a| (roundings of the sum and the quotient), u = 2^-53; L and H as computed here
 * are off by at most 2 u * S, S = |lo * 2a| + |hi * 2a| + |b|; close to a threshold sq <= S. The filter therefore answers only
 * when sq is further than E = 2^-47 * S = 64 u S (that budget is 5 u S; the libm-pow build's 1-ulp root fits as well) from both |L|
 * and |H|, and |L|, |H| themselves exceed E (their signs are then certain); otherwise it returns -1 and the caller
 * evaluates the reference's expressions. NaN or infinite intermediates fail every comparison and end in -1 as well; so do
 * problems scaled below 2^-450, where a square could round as a subnormal.
 * (The same idea for Map.is_occupied_c's sqrt(S) < R was built and measured as well: no gain, not kept.)
 * Exactness is checked, not assumed: oracle/root_filter_check.c compares filter and expressions on random and on
 * adversarial inputs (segment ends on the circle to within a few ulps, tangent lines, huge and tiny scales), and the GPU
 * suite's 700,000 borderline segments run through it (tests/test_cull.py).
 */
#ifndef RRTFND_ROOT_FILTER_H
#define RRTFND_ROOT_FILTER_H

#include <math.h>

#ifdef __CUDACC__
#define RRTFND_RF_HD __device__ __forceinline__
#define RRTFND_RF_ADD(a, b) __dadd_rn((a), (b))
#define RRTFND_RF_SUB(a, b) __dsub_rn((a), (b))
#define RRTFND_RF_MUL(a, b) __dmul_rn((a), (b))
#else
#define RRTFND_RF_HD static inline
static inline double rrtfnd_rf_add(double a, double b) { volatile double r = a + b; return r; }
static inline double rrtfnd_rf_sub(double a, double b) { volatile double r = a - b; return r; }
static inline double rrtfnd_rf_mul(double a, double b) { volatile double r = a * b; return r; }
#define RRTFND_RF_ADD(a, b) rrtfnd_rf_add((a), (b))
#define RRTFND_RF_SUB(a, b) rrtfnd_rf_sub((a), (b))
#define RRTFND_RF_MUL(a, b) rrtfnd_rf_mul((a), (b))
#endif

#ifndef RRTFND_RF_MARGIN
#define RRTFND_RF_MARGIN 0x1p-47      /* E / S; the checker is also built with 0 and 2^-53 to show that it has teeth */
#endif

/* 1 = a root lies strictly inside (lo, hi), 0 = none does, -1 = too close to call. delta >= 0, a2 = 2a > 0, lo < hi. */
RRTFND_RF_HD int rrtfnd_root_filter(double b, double delta, double a2, double lo, double hi) {
    const double pl = RRTFND_RF_MUL(lo, a2), ph = RRTFND_RF_MUL(hi, a2);
    const double L = RRTFND_RF_ADD(pl, b), H = RRTFND_RF_ADD(ph, b);
    const double E = RRTFND_RF_MUL(RRTFND_RF_ADD(RRTFND_RF_ADD(fabs(pl), fabs(ph)), fabs(b)), RRTFND_RF_MARGIN);
    const double aL = fabs(L), aH = fabs(H);
    if (!(aL > E && aH > E && E > 0x1p-500)) return -1;         /* (far above the subnormal range: the squares below round normally) */
    const double big = aL > aH ? aL : aH, small = aL > aH ? aH : aL;
    const double up = RRTFND_RF_ADD(big, E), dn = RRTFND_RF_SUB(big, E);
    if (delta > RRTFND_RF_MUL(up, up)) return 0;                 /* sq beyond the far threshold: both roots outside */
    if (!(delta < RRTFND_RF_MUL(dn, dn))) return -1;
    if ((L < 0.0) != (H < 0.0)) return 1;                       /* thresholds on both sides of 0: one root inside */
    const double up2 = RRTFND_RF_ADD(small, E), dn2 = RRTFND_RF_SUB(small, E);
    if (delta > RRTFND_RF_MUL(up2, up2)) return 1;
    if (delta < RRTFND_RF_MUL(dn2, dn2)) return 0;
    return -1;
}

#endif
