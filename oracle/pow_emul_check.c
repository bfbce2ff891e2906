/* pow_emul_check.c — TEST INFRASTRUCTURE: csrc/libm_pow.h against this machine's libm pow(), bit for bit.
 *
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-builtin-pow -o _build/pow_emul_check pow_emul_check.c -lm -lpthread
 *   _build/pow_emul_check [millions of arguments per exponent = 200] [threads = 8]
 *
 * Arguments: uniform bit patterns over the positive normal doubles (every binade), uniform values in the ranges planner
 * quantities live in, and the neighbourhoods of powers of two. Exit code 1 and a report on the first disagreement.
 * Also a library (orc_pow_emul_*) for tests/test_libm_pow.py. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../rrt_star_fnd_algorithm_b200/csrc/libm_pow.h"

static uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double as_double(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static uint64_t as_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }

/* argument number n of family f */
static double argument(int family, uint64_t* seed) {
    const uint64_t u = splitmix(seed);
    switch (family) {
        case 0: {                                         /* any positive normal double whose square / root stays normal */
            const uint64_t e = 1023 - 360 + (u >> 52) % 720, m = u & 0xfffffffffffffull;
            return as_double((e << 52) | m);
        }
        case 1: return (double)(u >> 11) * (1.0 / 9007199254740992.0) * 600.0 - 300.0;      /* coordinates, slopes, b */
        case 2: return (double)(u >> 11) * (1.0 / 9007199254740992.0) * 1.0e6;               /* deltas */
        case 3: {                                         /* next to a power of two */
            const uint64_t e = 1023 - 40 + (u >> 40) % 80;
            const int64_t d = (int64_t)(u & 0xffff) - 0x8000;
            return as_double(((e << 52)) + (uint64_t)d);
        }
        default: return (double)(u >> 11) * (1.0 / 9007199254740992.0) * 4.0 + 1e-3;           /* around 1 */
    }
}

typedef struct { uint64_t seed; long long n; long long bad2, badh, diff2, diffh; double first_bad; } job_t;

static void* worker(void* p) {
    job_t* j = (job_t*)p;
    uint64_t s = j->seed;
    for (long long i = 0; i < j->n; ++i) {
        const double x = argument((int)(i % 5), &s);
        const double a = pow(x, 2.0), b = rrtfnd_libm_pow2(x);
        if (as_bits(a) != as_bits(b)) { if (!j->bad2 && !j->badh) j->first_bad = x; ++j->bad2; }
        if (as_bits(a) != as_bits(x * x)) ++j->diff2;
        if (x >= 0.0) {
            const double c = pow(x, 0.5), d = rrtfnd_libm_pow_half(x);
            if (as_bits(c) != as_bits(d)) { if (!j->bad2 && !j->badh) j->first_bad = x; ++j->badh; }
            if (as_bits(c) != as_bits(sqrt(x))) ++j->diffh;
        }
    }
    return 0;
}

/* library entry for the Python test: counts over n arguments; out = {bad2, badh, diff2, diffh} */
void orc_pow_emul_check(long long n, int threads, uint64_t seed, long long* out, double* first_bad) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    job_t jobs[64];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){seed + 7919ull * (uint64_t)t, n / threads, 0, 0, 0, 0, 0.0};
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    out[0] = out[1] = out[2] = out[3] = 0;
    *first_bad = 0.0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], 0);
        out[0] += jobs[t].bad2; out[1] += jobs[t].badh; out[2] += jobs[t].diff2; out[3] += jobs[t].diffh;
        if ((jobs[t].bad2 || jobs[t].badh) && *first_bad == 0.0) *first_bad = jobs[t].first_bad;
    }
}
double orc_pow_emul_pow2(double x) { return rrtfnd_libm_pow2(x); }
double orc_pow_emul_half(double x) { return rrtfnd_libm_pow_half(x); }
void orc_pow_emul_array(const double* x, long long n, int half, double* out) {
    for (long long i = 0; i < n; ++i) out[i] = half ? rrtfnd_libm_pow_half(x[i]) : rrtfnd_libm_pow2(x[i]);
}

#ifndef ORC_POW_EMUL_LIB
int main(int argc, char** argv) {
    const long long n = (argc > 1 ? atoll(argv[1]) : 200) * 1000000ll;
    const int threads = argc > 2 ? atoi(argv[2]) : 8;
    long long out[4];
    double fb;
    orc_pow_emul_check(n, threads, 12345, out, &fb);
    printf("%lld arguments: pow(x, 2) != emulation in %lld, pow(x, 0.5) != emulation in %lld; for scale: pow(x, 2) != x*x in %lld, "
           "pow(x, 0.5) != sqrt(x) in %lld\n", n, out[0], out[1], out[2], out[3]);
    if (out[0] || out[1]) { printf("first disagreement at x = %a\n", fb); return 1; }
    return 0;
}
#endif
