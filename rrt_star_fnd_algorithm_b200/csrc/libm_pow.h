/* libm_pow.h — `x ** 2` and `x ** 0.5` as the reference's interpreter evaluates them, for host and device.
 *
 * The reference writes `b ** 2`, `line.dir ** 2`, `delta ** 0.5`, `(p[0] - ox) ** 2` ... (src/algorithm.py:602-603, :617,
 * :619-620; src/map.py:44). On Python floats and numpy float64 scalars `**` is libm pow(), and glibc's pow is not
 * correctly rounded: it differs from fl(x*x) in ~0.08 % and from sqrt(x) in ~0.06 % of inputs, by one ulp.
 * This header restates the main path of glibc 2.39's pow for x86-64 — the FMA build (__pow_fma) that libm's ifunc resolver
 * selects on every CPU with FMA and AVX2 — operation for operation, in the order and with the fused multiply-adds of that
 * binary (the compiler contracted more than the `__FP_FAST_FMA` source lines: t1, lo1, the polynomial steps, the exp
 * reduction and the final scale are all FMAs there). Tables: libm_pow_tables.h (generated from the same libm).
 *
 * Exactness is checked, not assumed: oracle/pow_emul_check.c compares this function with the container's pow() over
 * billions of arguments (tests/test_libm_pow.py holds a quick version), and the GPU operator rrtfnd_pow_libm is compared
 * with Python's own `**` on the GPU box. Outside the main path (zero, infinite or NaN argument, a subnormal argument of `** 2`,
 * |log-domain product| >= 512, i.e. results beyond ~2^+-738) the functions return x*x / sqrt(x) (+0 for sqrt of -0), which is
 * what libm returns for 0, inf and NaN; planner coordinates never leave the main path.
 *
 * One source for gcc (oracle, checker) and nvcc (device): every operation is an explicit mul / add / fma, so neither
 * compiler's contraction setting matters.
 */
#ifndef RRTFND_LIBM_POW_H
#define RRTFND_LIBM_POW_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__            /* nvcc: device-only functions, tables in device memory */
#define RRTFND_POW_HD __device__ __forceinline__
#define RRTFND_POW_TABLE static __device__ const
#define RRTFND_POW_DEVICE 1
#else                        /* gcc / g++: the oracle and the checker */
#define RRTFND_POW_HD static inline
#define RRTFND_POW_TABLE static const
#endif
#include "libm_pow_tables.h"

RRTFND_POW_HD double rrtfnd_pow_mul(double a, double b) {
#ifdef RRTFND_POW_DEVICE
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;      /* one rounding, never fused with a neighbouring add */
    return r;
#endif
}
RRTFND_POW_HD double rrtfnd_pow_add(double a, double b) {
#ifdef RRTFND_POW_DEVICE
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
RRTFND_POW_HD double rrtfnd_pow_fma(double a, double b, double c) {
#ifdef RRTFND_POW_DEVICE
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
RRTFND_POW_HD uint64_t rrtfnd_pow_bits(double x) {
#ifdef RRTFND_POW_DEVICE
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
RRTFND_POW_HD double rrtfnd_pow_from_bits(uint64_t u) {
#ifdef RRTFND_POW_DEVICE
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

/* pow(x, y) for the bits ix of a positive normal x and y in {2.0, 0.5} (any y with 2^-65 <= |y| < 2^63 works);
 * `fallback` is returned where glibc leaves its main path. */
RRTFND_POW_HD double rrtfnd_libm_pow_main(uint64_t ix, double y, double fallback) {
#define RRTFND_POW_LOGTAB(i, j) rrtfnd_pow_log_tab[i][j]
    const double* A = rrtfnd_pow_log_poly;
    /* ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1), as hi + tail */
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double z = rrtfnd_pow_from_bits(iz), kd = (double)k;
    const double invc = RRTFND_POW_LOGTAB(i, 0), logc = RRTFND_POW_LOGTAB(i, 1), logctail = RRTFND_POW_LOGTAB(i, 2);
    const double t1 = rrtfnd_pow_fma(kd, RRTFND_POW_LN2HI, logc);
    const double lo1 = rrtfnd_pow_fma(kd, RRTFND_POW_LN2LO, logctail);
    const double r = rrtfnd_pow_fma(z, invc, -1.0);
    const double ar = rrtfnd_pow_mul(r, A[0]);
    const double q12 = rrtfnd_pow_fma(r, A[2], A[1]);
    const double q34 = rrtfnd_pow_fma(r, A[4], A[3]);
    const double t2 = rrtfnd_pow_add(r, t1);
    const double lo2 = rrtfnd_pow_add(rrtfnd_pow_add(t1, -t2), r);
    const double ar2 = rrtfnd_pow_mul(r, ar);
    const double ar3 = rrtfnd_pow_mul(r, ar2);
    const double lo3 = rrtfnd_pow_fma(ar, r, -ar2);
    const double hi = rrtfnd_pow_add(t2, ar2);
    const double q56 = rrtfnd_pow_fma(r, A[6], A[5]);
    const double lo4 = rrtfnd_pow_add(rrtfnd_pow_add(t2, -hi), ar2);
    double q = rrtfnd_pow_fma(q56, ar2, q34);
    q = rrtfnd_pow_fma(ar2, q, q12);
    double lo = rrtfnd_pow_add(lo1, lo2);
    lo = rrtfnd_pow_add(lo, lo3);
    lo = rrtfnd_pow_add(lo, lo4);
    lo = rrtfnd_pow_fma(ar3, q, lo);
    const double lhi = rrtfnd_pow_add(hi, lo);
    const double ltail = rrtfnd_pow_add(rrtfnd_pow_add(hi, -lhi), lo);
    /* ---- y * log(x) as ehi + elo */
    const double ehi = rrtfnd_pow_mul(y, lhi);
    const double e1 = rrtfnd_pow_fma(lhi, y, -ehi);
    const double elo = rrtfnd_pow_fma(y, ltail, e1);
    /* ---- exp_inline(ehi, elo) */
    const unsigned abstop = (unsigned)(rrtfnd_pow_bits(ehi) >> 52) & 0x7ffu;
    if (abstop < 0x3c9u) return 1.0;                          /* |ehi| < 2^-54: 1.0 + ehi rounds to 1.0 */
    if (abstop > 0x407u) return fallback;                     /* |ehi| >= 512: glibc's special-case code */
    const double kds = rrtfnd_pow_fma(ehi, RRTFND_EXP_INVLN2N, RRTFND_EXP_SHIFT);
    const uint64_t ki = rrtfnd_pow_bits(kds);
    const double kd2 = rrtfnd_pow_add(kds, -RRTFND_EXP_SHIFT);
    const double r0 = rrtfnd_pow_fma(kd2, RRTFND_EXP_NEGLN2HIN, ehi);
    double rr = rrtfnd_pow_fma(kd2, RRTFND_EXP_NEGLN2LON, r0);
    rr = rrtfnd_pow_add(elo, rr);
    const unsigned idx = 2u * (unsigned)(ki & 127u);
    const uint64_t sbits = rrtfnd_exp_tab[idx + 1] + (ki << 45);
    const double tail = rrtfnd_pow_from_bits(rrtfnd_exp_tab[idx]);
    const double* C = rrtfnd_exp_poly;                        /* C[0] = C2 ... C[3] = C5 */
    const double p23 = rrtfnd_pow_fma(rr, C[1], C[0]);
    const double tr = rrtfnd_pow_add(rr, tail);
    const double r2 = rrtfnd_pow_mul(rr, rr);
    const double p45 = rrtfnd_pow_fma(rr, C[3], C[2]);
    double t = rrtfnd_pow_fma(p23, r2, tr);
    const double r4 = rrtfnd_pow_mul(r2, r2);
    t = rrtfnd_pow_fma(p45, r4, t);
    const double scale = rrtfnd_pow_from_bits(sbits);
    return rrtfnd_pow_fma(t, scale, scale);
#undef RRTFND_POW_LOGTAB
}

/* x ** 2: pow(x, 2.0). A negative base with an even integer exponent takes the same path with |x|. */
RRTFND_POW_HD double rrtfnd_libm_pow2(double x) {
    const uint64_t ix = rrtfnd_pow_bits(x) & 0x7fffffffffffffffull;
    const double sq = rrtfnd_pow_mul(x, x);
    if ((unsigned)(ix >> 52) - 1u > 0x7fdu) return sq;       /* zero, subnormal, inf, NaN */
    return rrtfnd_libm_pow_main(ix, 2.0, sq);
}

/* x ** 0.5: pow(x, 0.5) for x >= 0 (the reference only takes it of a non-negative delta, src/algorithm.py:569, :602). */
RRTFND_POW_HD double rrtfnd_libm_pow_half(double x) {
    uint64_t ix = rrtfnd_pow_bits(x);
    const double s = sqrt(x);
    if ((ix << 1) == 0) return 0.0;                           /* pow(+-0, 0.5) = +0 (sqrt keeps the sign of -0) */
    if ((unsigned)(ix >> 52) - 1u > 0x7fdu) {
        if ((ix >> 52) != 0) return s;                        /* negative, inf, NaN */
        ix = rrtfnd_pow_bits(rrtfnd_pow_mul(x, 4503599627370496.0)) - (52ull << 52);   /* subnormal: normalised as glibc does */
    }
    return rrtfnd_libm_pow_main(ix, 0.5, s);
}

#endif /* RRTFND_LIBM_POW_H */
