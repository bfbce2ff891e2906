/* root_filter_check.c — TEST INFRASTRUCTURE: csrc/root_filter.h against the expressions it short-cuts.
 *
 *   make -C oracle rootcheck && oracle/_build/root_filter_check [millions of cases per family] [threads]
 *
 * For every generated (segment, circle) the line / circle quadratic is set up exactly as csrc/arith.cuh does (make_seg,
 * seg_hits_circle; src/line.py:5-16, src/algorithm.py:607-621), then the reference's decision - square root, two
 * divisions, four strict comparisons (src/algorithm.py:569-576, :594-604, :624-635) - is compared with the filter's
 * whenever the filter answers. Families: planner-like random input; a segment end on the circle to within 0..64 ulps;
 * tangent lines; thresholds forced to within 0..2000 ulps in delta-space; huge / tiny / negative / zero-crossing scales.
 * Prints the number of cases, how many the filter decided, and the mismatches (must be 0). Exit status 1 on a mismatch.
 * As a library (-DORC_ROOT_FILTER_LIB) it exports the filter for tests/test_root_filter.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../rrt_star_fnd_algorithm_b200/csrc/root_filter.h"

static inline double vadd(double a, double b) { volatile double r = a + b; return r; }
static inline double vsub(double a, double b) { volatile double r = a - b; return r; }
static inline double vmul(double a, double b) { volatile double r = a * b; return r; }
static inline double vdiv(double a, double b) { volatile double r = a / b; return r; }

/* the reference's decision on (b, delta >= 0, a2, lo, hi) */
static int slow_decision(double b, double delta, double a2, double lo, double hi) {
    const double sq = sqrt(delta);
    const double x1 = vdiv(vadd(-b, sq), a2), x2 = vdiv(vsub(-b, sq), a2);
    return (x1 > lo && x1 < hi) || (x2 > lo && x2 < hi);
}

/* quadratic of Line(p1, p2) against the circle (ox, oy, r), as arith.cuh sets it up; returns 0 for vertical / delta < 0 */
static int setup(double p1x, double p1y, double p2x, double p2y, double ox, double oy, double r, double* b, double* delta,
                 double* a2, double* lo, double* hi) {
    if (p2x == p1x) return 0;
    const double m = vdiv(vsub(p2y, p1y), vsub(p2x, p1x));
    const double k = vsub(p1y, vmul(m, p1x));
    const double mk = vmul(m, k), kk = vmul(k, k);
    const double a = vadd(1.0, vmul(m, m));
    *a2 = vadd(a, a);
    const double a4 = vmul(4.0, a);
    *lo = p1x < p2x ? p1x : p2x;
    *hi = p1x < p2x ? p2x : p1x;
    const double K = vadd(vadd(-vmul(r, r), vmul(ox, ox)), vmul(oy, oy));
    *b = vmul(2.0, vsub(vadd(-ox, mk), vmul(m, oy)));
    const double c = vadd(vsub(K, vmul(vadd(oy, oy), k)), kk);
    *delta = vsub(vmul(*b, *b), vmul(a4, c));
    return *delta >= 0.0;
}

#ifdef ORC_ROOT_FILTER_LIB
int orc_root_filter(double b, double delta, double a2, double lo, double hi) { return rrtfnd_root_filter(b, delta, a2, lo, hi); }
int orc_root_slow(double b, double delta, double a2, double lo, double hi) { return slow_decision(b, delta, a2, lo, hi); }
/* through_obstacle (src/algorithm.py:581-591) for n segments against M circles (x, y, r) the way the device evaluates a circle
 * (arith.cuh seg_hits_circle): vertical rule, discriminant, root filter, and the reference's expressions only when the filter
 * declines. hit[i] = the OR over the circles; counts[0] += circle tests with delta >= 0, counts[1] += those the filter declined. */
void orc_root_through_obstacle(const double* segs, long long n, const double* circles, int M, unsigned char* hit, long long* counts) {
    for (long long i = 0; i < n; ++i) {
        const double* s = segs + 4 * i;
        int h = 0;
        for (int j = 0; j < M && !h; ++j) {
            const double* c = circles + 3 * j;
            if (s[2] == s[0]) { h = fabs(vsub(c[0], s[0])) <= c[2]; continue; }
            double b, delta, a2, lo, hi;
            if (!setup(s[0], s[1], s[2], s[3], c[0], c[1], c[2], &b, &delta, &a2, &lo, &hi)) continue;
            ++counts[0];
            int f = rrtfnd_root_filter(b, delta, a2, lo, hi);
            if (f < 0) { ++counts[1]; f = slow_decision(b, delta, a2, lo, hi); }
            h = f;
        }
        hit[i] = (unsigned char)h;
    }
}
int orc_root_setup(const double* seg, const double* circle, double* out5) {
    return setup(seg[0], seg[1], seg[2], seg[3], circle[0], circle[1], circle[2], out5, out5 + 1, out5 + 2, out5 + 3, out5 + 4);
}
#else

typedef struct { uint64_t s; } rng_t;
static inline uint64_t next64(rng_t* g) {                 /* splitmix64 */
    uint64_t z = (g->s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double uni(rng_t* g) { return (double)(next64(g) >> 11) * 0x1p-53; }
static inline double range(rng_t* g, double a, double b) { return a + (b - a) * uni(g); }
static double step_ulps(double x, int n) {
    for (; n > 0; --n) x = nextafter(x, INFINITY);
    for (; n < 0; ++n) x = nextafter(x, -INFINITY);
    return x;
}

enum { N_FAM = 6 };
typedef struct { long long n[N_FAM], valid[N_FAM], decided[N_FAM], bad[N_FAM]; long long per_family; uint64_t seed; } job_t;

static void check(job_t* j, int fam, double b, double delta, double a2, double lo, double hi) {
    ++j->valid[fam];
    const int f = rrtfnd_root_filter(b, delta, a2, lo, hi);
    if (f < 0) return;
    ++j->decided[fam];
    if (f != slow_decision(b, delta, a2, lo, hi)) {
        if (++j->bad[fam] <= 3)
            fprintf(stderr, "MISMATCH family %d: b=%a delta=%a a2=%a lo=%a hi=%a filter=%d\n", fam, b, delta, a2, lo, hi, f);
    }
}

static void* run(void* arg) {
    job_t* j = (job_t*)arg;
    rng_t g = { j->seed };
    static const double scales[] = { 1.0, 1e-3, 1e3, 1e-30, 1e30, 1e-150, 1e150, 1e-160, 3e153 };
    double b, delta, a2, lo, hi;
    for (long long i = 0; i < j->per_family; ++i) {
        /* 0: planner-like: 200 x 200 map, r in [1, 4], short and long segments */
        {
            const double ox = range(&g, 0, 200), oy = range(&g, 0, 200), r = range(&g, 1, 4);
            const double len = (i & 1) ? 12.0 : 200.0;
            const double p1x = range(&g, 0, 200), p1y = range(&g, 0, 200);
            const double p2x = (i & 1) ? p1x + range(&g, -len, len) : range(&g, 0, 200);
            const double p2y = (i & 1) ? p1y + range(&g, -len, len) : range(&g, 0, 200);
            /* aim every other one at the circle so that delta >= 0 is common */
            const double tx = (i & 2) ? ox + range(&g, -1.5, 1.5) * r : p2x, ty = (i & 2) ? oy + range(&g, -1.5, 1.5) * r : p2y;
            ++j->n[0];
            if (setup(p1x, p1y, tx, ty, ox, oy, r, &b, &delta, &a2, &lo, &hi)) check(j, 0, b, delta, a2, lo, hi);
        }
        /* 1: one end of the segment on the circle to within 0..64 ulps (the root meets lo or hi) */
        {
            const double ox = range(&g, 0, 200), oy = range(&g, 0, 200), r = range(&g, 0.5, 8);
            const double th = range(&g, 0, 6.283185307179586);
            double ex = ox + r * cos(th), ey = oy + r * sin(th);
            ex = step_ulps(ex, (int)(next64(&g) % 129) - 64);
            ey = step_ulps(ey, (int)(next64(&g) % 129) - 64);
            const double qx = range(&g, 0, 200), qy = range(&g, 0, 200);
            ++j->n[1];
            const int ok = (i & 1) ? setup(ex, ey, qx, qy, ox, oy, r, &b, &delta, &a2, &lo, &hi)
                                   : setup(qx, qy, ex, ey, ox, oy, r, &b, &delta, &a2, &lo, &hi);
            if (ok) check(j, 1, b, delta, a2, lo, hi);
        }
        /* 2: (nearly) tangent lines: delta around 0 */
        {
            const double ox = range(&g, 0, 200), oy = range(&g, 0, 200), r = range(&g, 0.5, 8);
            const double th = range(&g, 0, 6.283185307179586), t = range(&g, 1, 50), u2 = range(&g, -50, 50);
            const double rr = r * (1.0 + range(&g, -1, 1) * ((i & 1) ? 1e-15 : 1e-9));
            const double cx = ox + rr * cos(th), cy = oy + rr * sin(th);         /* point of tangency; direction (-sin, cos) */
            ++j->n[2];
            if (setup(cx - t * -sin(th), cy - t * cos(th), cx + u2 * -sin(th), cy + u2 * cos(th), ox, oy, r, &b, &delta, &a2, &lo, &hi))
                check(j, 2, b, delta, a2, lo, hi);
        }
        /* 3: thresholds forced in delta-space: delta = (|L| or |H|)^2 moved by 0..2000 ulps */
        {
            const double s = scales[next64(&g) % 9];
            lo = range(&g, -200, 200) * s; hi = lo + range(&g, 1e-6, 200) * s;
            const double m = tan(range(&g, -1.57, 1.57));
            const double a = vadd(1.0, vmul(m, m));
            a2 = vadd(a, a);
            b = range(&g, -4, 4) * (fabs(lo) + fabs(hi)) * a;
            const double T = (i & 1) ? vadd(vmul(lo, a2), b) : vadd(vmul(hi, a2), b);
            delta = step_ulps(vmul(T, T), (int)(next64(&g) % 4001) - 2000);
            ++j->n[3];
            if (delta >= 0.0 && lo < hi) check(j, 3, b, delta, a2, lo, hi);
        }
        /* 4: planner geometry at other scales and around the origin (sign changes of lo, hi, L, H) */
        {
            const double s = scales[next64(&g) % 9];
            const double ox = range(&g, -5, 5) * s, oy = range(&g, -5, 5) * s, r = range(&g, 0.5, 4) * s;
            ++j->n[4];
            if (setup(range(&g, -8, 8) * s, range(&g, -8, 8) * s, range(&g, -8, 8) * s, range(&g, -8, 8) * s, ox, oy, r, &b, &delta, &a2, &lo, &hi))
                check(j, 4, b, delta, a2, lo, hi);
        }
        /* 5: raw doubles: random bit patterns in wide exponent ranges (NaN / inf / subnormal included) */
        {
            union { uint64_t u; double d; } v[5];
            for (int q = 0; q < 5; ++q) v[q].u = next64(&g);
            lo = v[0].d; hi = v[1].d; b = v[2].d; delta = fabs(v[3].d); a2 = vadd(2.0, fabs(v[4].d));
            if ((i & 3) == 0) { a2 = range(&g, 2, 1e6); }
            ++j->n[5];
            if (lo < hi && delta >= 0.0 && a2 >= 2.0) check(j, 5, b, delta, a2, lo, hi);
        }
    }
    return NULL;
}

int main(int argc, char** argv) {
    const long long millions = argc > 1 ? atoll(argv[1]) : 20;
    int threads = argc > 2 ? atoi(argv[2]) : 8;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    static job_t jobs[256];
    for (int t = 0; t < threads; ++t) {
        memset(&jobs[t], 0, sizeof(job_t));
        jobs[t].per_family = millions * 1000000ll / threads;
        jobs[t].seed = 0x1234567ull + 0x9e3779b97f4a7c15ull * (uint64_t)(t + 1);
        pthread_create(&th[t], NULL, run, &jobs[t]);
    }
    job_t sum;
    memset(&sum, 0, sizeof sum);
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        for (int f = 0; f < N_FAM; ++f) {
            sum.n[f] += jobs[t].n[f]; sum.valid[f] += jobs[t].valid[f]; sum.decided[f] += jobs[t].decided[f]; sum.bad[f] += jobs[t].bad[f];
        }
    }
    static const char* names[N_FAM] = { "planner-like", "end on circle", "tangent", "forced thresholds", "scales / origin", "raw doubles" };
    long long bad = 0;
    for (int f = 0; f < N_FAM; ++f) {
        printf("%-18s cases %12lld  delta>=0 %12lld  decided %12lld (%.4f %%)  mismatches %lld\n", names[f], sum.n[f], sum.valid[f],
               sum.decided[f], sum.valid[f] ? 100.0 * sum.decided[f] / sum.valid[f] : 0.0, sum.bad[f]);
        bad += sum.bad[f];
    }
    return bad ? 1 : 0;
}
#endif
