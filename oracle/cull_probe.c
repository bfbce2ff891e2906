/*
 * cull_probe.c — adversarial check of the obstacle-grid culling rule. TEST INFRASTRUCTURE ONLY.
 *
 * Claim under test (DESIGN.md "Obstacle grid"): for a non-vertical segment with |Line.dir| <= slope_max and all
 * coordinates / radii bounded by coord_max, the reference's intersection_circle (src/algorithm.py:555-578,
 * restated op-for-op below exactly as in rrt_oracle.c) is False for every circle whose centre lies farther than
 * r + margin from the segment's bounding box in x or in y.
 *
 * orc_cull_probe places circles at clearance `scale * margin` beyond the box, preferring the worst case (the
 * infinite line passes through the circle just outside an endpoint), and counts predicate hits. scale = 1 must
 * give zero hits; a tiny scale must give many (which shows the placement is adversarial).
 */
#include <math.h>
#include <stdint.h>

static int seg_hits_circle(double p1x, double p1y, double p2x, double p2y, double ox, double oy, double r) {
    if (p2x == p1x) return fabs(ox - p1x) <= r;
    double m = (p2y - p1y) / (p2x - p1x);
    double k = p1y - m * p1x;
    double a = 1.0 + m * m;
    double b = 2.0 * ((-ox + m * k) - m * oy);
    double c = ((((-(r * r)) + ox * ox) + oy * oy) - (2.0 * oy) * k) + k * k;
    double delta = b * b - (4.0 * a) * c;
    if (delta < 0.0) return 0;
    double s = sqrt(delta);
    double x1 = (-b + s) / (2.0 * a);
    double x2 = (-b - s) / (2.0 * a);
    double lo = p1x < p2x ? p1x : p2x;
    double hi = p1x < p2x ? p2x : p1x;
    return (x1 > lo && x1 < hi) || (x2 > lo && x2 < hi);
}

static uint64_t s_state;
static double u01(void) {  /* xorshift64* */
    s_state ^= s_state >> 12; s_state ^= s_state << 25; s_state ^= s_state >> 27;
    return (double)((s_state * 2685821237136412197ULL) >> 11) * (1.0 / 9007199254740992.0);
}
static double urange(double a, double b) { return a + (b - a) * u01(); }
static double logu(double a, double b) { return exp(urange(log(a), log(b))); }

/* returns the number of predicate hits among n_trials culled (segment, circle) pairs; *n_valid = pairs generated */
int64_t orc_cull_probe(uint64_t seed, int64_t n_trials, double coord_max, double slope_max, double margin,
                       double scale, int64_t* n_valid) {
    s_state = seed * 0x9E3779B97F4A7C15ULL + 1;
    const double S = coord_max;
    int64_t hits = 0, valid = 0;
    for (int64_t t = 0; t < n_trials; ++t) {
        const int positive = u01() < 0.5;                       /* maps like [0, S]^2 or [-S, S]^2 */
        const double lo_c = positive ? 0.0 : -S;
        double p1x = urange(lo_c, S), p1y = urange(lo_c, S);
        const double L = logu(1e-3, S);
        double dx, dy;
        if (u01() < 0.5) { const double th = urange(0.0, 6.283185307179586); dx = L * cos(th); dy = L * sin(th); }
        else {                                                  /* slope log-uniform up to the limit */
            const double m = logu(1e-6, slope_max) * (u01() < 0.5 ? -1.0 : 1.0);
            dx = L / sqrt(1.0 + m * m) * (u01() < 0.5 ? -1.0 : 1.0); dy = m * dx;
        }
        double p2x = p1x + dx, p2y = p1y + dy;
        if (p2x < lo_c || p2x > S) p2x = p1x - dx;
        if (p2y < lo_c || p2y > S) p2y = p1y - dy;
        if (p2x < lo_c || p2x > S || p2y < lo_c || p2y > S || p2x == p1x) continue;
        const double mh = (p2y - p1y) / (p2x - p1x);
        if (!(fabs(mh) <= slope_max)) continue;
        const double kh = p1y - mh * p1x;
        const double r = logu(1e-3, fmin(S / 4.0, 50.0));
        const double g = margin * scale;
        const double xlo = fmin(p1x, p2x), xhi = fmax(p1x, p2x), ylo = fmin(p1y, p2y), yhi = fmax(p1y, p2y);
        double ox, oy;
        const double sel = u01();
        const double far = (u01() < 0.7) ? 0.0 : logu(1e-9, 1.0) * r;   /* mostly right at the culling boundary */
        if (sel < 0.35) {                                       /* beyond an x end, on the infinite line */
            ox = (u01() < 0.5) ? xhi + (r + g + far) : xlo - (r + g + far);
            oy = mh * ox + kh + urange(-r, r) * sqrt(1.0 + mh * mh) * u01();
        } else if (sel < 0.7) {                                 /* beyond a y end, on the infinite line */
            oy = (u01() < 0.5) ? yhi + (r + g + far) : ylo - (r + g + far);
            ox = (mh != 0.0 ? (oy - kh) / mh : urange(xlo, xhi)) + urange(-r, r) * u01();
        } else if (sel < 0.85) {                                /* beside the box */
            ox = (u01() < 0.5) ? xhi + (r + g + far) : xlo - (r + g + far);
            oy = urange(ylo - r, yhi + r);
        } else {
            oy = (u01() < 0.5) ? yhi + (r + g + far) : ylo - (r + g + far);
            ox = urange(xlo - r, xhi + r);
        }
        if (!(fabs(ox) <= S) || !(fabs(oy) <= S)) continue;
        /* keep only pairs the rule culls: centre farther than r + g from the box in x or in y */
        if (!(ox > xhi + (r + g) || ox < xlo - (r + g) || oy > yhi + (r + g) || oy < ylo - (r + g))) {
            if (!(ox >= xhi + (r + g) || ox <= xlo - (r + g) || oy >= yhi + (r + g) || oy <= ylo - (r + g))) continue;
        }
        ++valid;
        hits += seg_hits_circle(p1x, p1y, p2x, p2y, ox, oy, r);
    }
    if (n_valid) *n_valid = valid;
    return hits;
}
