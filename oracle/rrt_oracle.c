/*
 * rrt_oracle.c — CPU restatement of the reference's per-iteration hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (rrt_star_fnd_algorithm_b200/) may import, link or
 * execute this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs have an independent checker and a CPU number. It is pinned against the
 * reference itself: oracle/make_golden.py runs the unmodified Python reference (oracle/refharness.py)
 * and tests/test_oracle_vs_golden.py requires bit-equal trees, ids, costs, removals and best paths.
 *
 * Brute force on purpose (no k-d tree, no grid): node ids index the arrays directly, every query is a
 * linear pass in ascending id. Arithmetic: one IEEE fp64 rounding per written operation
 * (-ffp-contract=off), fma() only inside dist_fma (see below).
 *
 * Each function cites the reference lines it follows (paths relative to the reference root).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rrtfnd.h"

#define ORC_RESERVE_WORDS 64 /* words that must be available when an attempt starts (WORDS mode) */

typedef struct orc_tree {
    int cap;          /* id capacity */
    double* x;        /* by id */
    double* y;
    double* ecost;    /* Graph.cost: edge cost to parent (src/graph.py:29) */
    int* parent;      /* by id, -1 = None */
    int* nchild;
    unsigned char* alive;
    int last_id, n_live, root;
    /* planner bookkeeping that survives between calls */
    int* finish;
    int n_finish, finish_cap;
    int* best_path;
    int best_len, path_cap;
    double best_cost;
    int solution_found;
    int iter;
    int status;
    int64_t attempts;
    uint64_t cursor;
    int64_t n_edge_checks, n_edge_checks_ref, n_circle_tests, n_scan_nodes, n_chain_hops;
    int n_rewired, n_removed;
    int fn_count;     /* RRT_star_FN's n_of_nodes (src/algorithm.py:209) */
    /* scratch */
    int* cand;
    double* cand_d;
    unsigned char* cand_hit;
} orc_tree_t;

/* ------------------------------------------------------------------------------------------- rng */

typedef struct orc_rng {
    int mode;
    const uint32_t* words;
    int64_t n_words;
    uint64_t seed, stream;
    uint64_t cursor;
    uint64_t blk_idx;
    uint32_t blk[4];
    int have_blk;
    int exhausted;
} orc_rng_t;

static void philox4x32_10(uint64_t blk, uint64_t stream, uint64_t seed, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t rng_next(orc_rng_t* r) {
    uint64_t i = r->cursor++;
    if (r->mode == RRTFND_STREAM_WORDS) {
        if ((int64_t)i >= r->n_words) { r->exhausted = 1; return 0; }
        return r->words[i];
    }
    uint64_t b = i >> 2;
    if (!r->have_blk || b != r->blk_idx) {
        philox4x32_10(b, r->stream, r->seed, r->blk);
        r->blk_idx = b;
        r->have_blk = 1;
    }
    return r->blk[i & 3];
}

/* CPython random_random(): 53-bit double from two MT words */
static double rng_random(orc_rng_t* r) {
    uint32_t a = rng_next(r) >> 5, b = rng_next(r) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

/* CPython Random._randbelow_with_getrandbits(n), n < 2**32 (random.choice, src/algorithm.py:813) */
static uint32_t rng_randbelow(orc_rng_t* r, uint32_t n) {
    int k = 0;
    for (uint32_t t = n; t; t >>= 1) ++k;
    uint32_t v = rng_next(r) >> (32 - k);
    while (v >= n && !r->exhausted) v = rng_next(r) >> (32 - k);
    return v;
}

/* --------------------------------------------------------------------------------- arithmetic */

/* `x ** 2` / `x ** 0.5` of the reference (src/algorithm.py:602-603, :617, :619-620; src/map.py:44): fl(x*x) / sqrt(x) by
 * default — what the CUDA planners compute — or, after orc_set_libm_pow(1), the restatement of the libm pow() those
 * expressions really call (csrc/libm_pow.h, bit-equal to glibc 2.39's pow on 2 x 10^10 checked arguments). */
#include "../rrt_star_fnd_algorithm_b200/csrc/libm_pow.h"
static int g_libm_pow = 0;
void orc_set_libm_pow(int on) { g_libm_pow = on; }
static inline double pow2_ref(double x) { return g_libm_pow ? rrtfnd_libm_pow2(x) : x * x; }
static inline double pow_half_ref(double x) { return g_libm_pow ? rrtfnd_libm_pow_half(x) : sqrt(x); }

/* calc_distance (src/algorithm.py:954-961): np.linalg.norm of a 1-D vector = sqrt(x.dot(x)); OpenBLAS
 * ddot's tail loop is compiled with FMA, giving sqrt(fma(dy, dy, fl(dx*dx))) (pinned by the goldens). */
static double dist_fma(double px, double py, double qx, double qy) {
    double dx = px - qx, dy = py - qy;
    return sqrt(fma(dy, dy, dx * dx));
}

/* np.linalg.norm(in_radius_pos - q_new, axis=1) (src/algorithm.py:842, :899): plain sum of squares */
static double dist_plain(double nx, double ny, double qx, double qy) {
    double dx = nx - qx, dy = ny - qy;
    return sqrt(dx * dx + dy * dy);
}

/* cKDTree's squared distance: orders k-NN (src/algorithm.py:684) and defines the ball d2 <= r*r (:837) */
static double dist2(double nx, double ny, double qx, double qy) {
    double dx = nx - qx, dy = ny - qy;
    return dx * dx + dy * dy;
}

/* intersection_circle(Line(p1, p2), circle) — src/line.py:5-16, src/algorithm.py:555-578, :594-635 */
/* K = ((-(r**2)) + ox**2) + oy**2, the circle-only part of `c` (:619), is evaluated by the Python side with the
 * reference's own expression on the caller's number types (`**` is libm pow there, which differs from r*r by an ulp
 * for a few radii), exactly as the product's obstacle table is built. */
#ifdef ORC_POW_AUDIT
/* Audit build (make audit): every predicate is ALSO evaluated with libm pow() at the reference's `**` sites
 * (src/algorithm.py:602-603, :617, :619-620; src/map.py:44) - the arithmetic CPython and numpy scalars really perform -
 * and compared with the x*x / sqrt form the oracle and the CUDA path use. Counters: [0] circle tests, [1] tests in which
 * some intermediate differs, [2] tests whose BOOLEAN differs, [3..5] the same for Map.is_occupied_c. */
static long long g_pow_audit[6];
void orc_pow_audit(long long* out, int reset) {
    for (int i = 0; i < 6; ++i) { out[i] = __atomic_load_n(&g_pow_audit[i], __ATOMIC_RELAXED); if (reset) __atomic_store_n(&g_pow_audit[i], 0, __ATOMIC_RELAXED); }
}
static int seg_hits_circle_plain(double p1x, double p1y, double p2x, double p2y, double ox, double oy, double r, double K);
static int seg_hits_circle(double p1x, double p1y, double p2x, double p2y, double ox, double oy, double r, double K) {
    const int plain = seg_hits_circle_plain(p1x, p1y, p2x, p2y, ox, oy, r, K);
    if (p2x == p1x) return plain;
    double m = (p2y - p1y) / (p2x - p1x), k = p1y - m * p1x;
    double a = 1.0 + pow(m, 2.0);
    double b = 2.0 * ((-ox + m * k) - m * oy);
    double c = (K - (2.0 * oy) * k) + pow(k, 2.0);
    double delta = pow(b, 2.0) - (4.0 * a) * c;
    int res = 0, differs = (a != 1.0 + m * m) || (c != (K - (2.0 * oy) * k) + k * k) || (delta != b * b - (4.0 * (1.0 + m * m)) * ((K - (2.0 * oy) * k) + k * k));
    if (!(delta < 0.0)) {
        double s = pow(delta, 0.5);
        differs = differs || (s != sqrt(delta));
        double x1 = (-b + s) / (2.0 * a), x2 = (-b - s) / (2.0 * a);
        double lo = p1x < p2x ? p1x : p2x, hi = p1x < p2x ? p2x : p1x;
        res = (x1 > lo && x1 < hi) || (x2 > lo && x2 < hi);
    }
    __atomic_fetch_add(&g_pow_audit[0], 1, __ATOMIC_RELAXED);
    if (differs) __atomic_fetch_add(&g_pow_audit[1], 1, __ATOMIC_RELAXED);
    if (res != plain) __atomic_fetch_add(&g_pow_audit[2], 1, __ATOMIC_RELAXED);
    return plain;
}
#define seg_hits_circle_impl seg_hits_circle_plain
#else
#define seg_hits_circle_impl seg_hits_circle
#endif
static int seg_hits_circle_impl(double p1x, double p1y, double p2x, double p2y, double ox, double oy, double r, double K) {
    if (p2x == p1x) {                         /* src/line.py:9; src/algorithm.py:563-567 */
        return fabs(ox - p1x) <= r;
    }
    double m = (p2y - p1y) / (p2x - p1x);     /* src/line.py:15 */
    double k = p1y - m * p1x;                 /* src/line.py:16 */
    double a = 1.0 + pow2_ref(m);             /* src/algorithm.py:617 */
    double b = 2.0 * ((-ox + m * k) - m * oy);            /* :618 */
    double c = (K - (2.0 * oy) * k) + pow2_ref(k); /* :619 */
    double delta = pow2_ref(b) - (4.0 * a) * c;     /* :620 */
    if (delta < 0.0) return 0;                /* :569 */
    double s = pow_half_ref(delta);           /* :602-603 */
    double x1 = (-b + s) / (2.0 * a);
    double x2 = (-b - s) / (2.0 * a);
    double lo = p1x < p2x ? p1x : p2x;        /* :632-633 */
    double hi = p1x < p2x ? p2x : p1x;
    return (x1 > lo && x1 < hi) || (x2 > lo && x2 < hi); /* :576 */
}

/* EXTENSION — rectangular obstacles. The reference declares Map.obstacles_r (src/map.py:18, :54) and never implements it:
 * PARITY UNPINNED BY THE REFERENCE. This restates the semantics documented in include/rrtfnd.h (a table row with a negative
 * third column is the rectangle (cx, cy, -hx, hy)), operation for operation as csrc/arith.cuh evaluates them, so that the
 * CUDA path has a checker. Segment: closed segment vs closed rectangle by separating axes. */
static int seg_hits_rect(double p1x, double p1y, double p2x, double p2y, double cx, double cy, double hx, double hy) {
    double ex = (p2x - p1x) * 0.5, ey = (p2y - p1y) * 0.5;
    double mx = (p1x + ex) - cx, my = (p1y + ey) - cy;
    double ax = fabs(ex), ay = fabs(ey);
    if (fabs(mx) > hx + ax) return 0;
    if (fabs(my) > hy + ay) return 0;
    double cross = mx * ey - my * ex, reach = hx * ay + hy * ax;
    return !(fabs(cross) > reach);
}
/* point: inside the closed rectangle, or closer to it than node_radius (strict, like src/map.py:45) */
static int point_in_inflated_rect(double px, double py, double cx, double cy, double hx, double hy, double node_radius) {
    double dx = fmax(fabs(px - cx) - hx, 0.0), dy = fmax(fabs(py - cy) - hy, 0.0);
    if (dx == 0.0 && dy == 0.0) return 1;
    return sqrt(dx * dx + dy * dy) < node_radius;
}

/* through_obstacle (src/algorithm.py:581-591); obst is [M][4] ox, oy, r, K */
static int through_obstacle(double p1x, double p1y, double p2x, double p2y, const double* obst, int M,
                            int64_t* n_tests) {
    for (int j = 0; j < M; ++j) {
        if (n_tests) ++*n_tests;
        if (obst[4 * j + 2] < 0.0) {          /* rectangle row (extension) */
            if (seg_hits_rect(p1x, p1y, p2x, p2y, obst[4 * j], obst[4 * j + 1], -obst[4 * j + 2], obst[4 * j + 3])) return 1;
            continue;
        }
        if (seg_hits_circle(p1x, p1y, p2x, p2y, obst[4 * j], obst[4 * j + 1], obst[4 * j + 2], obst[4 * j + 3])) return 1;
    }
    return 0;
}

/* Map.is_occupied_c (src/map.py:38-47) */
static int is_occupied(double px, double py, const double* obst, int M, double node_radius) {
    for (int j = 0; j < M; ++j) {
        if (obst[4 * j + 2] < 0.0) {          /* rectangle row (extension) */
            if (point_in_inflated_rect(px, py, obst[4 * j], obst[4 * j + 1], -obst[4 * j + 2], obst[4 * j + 3], node_radius)) return 1;
            continue;
        }
        double dx = px - obst[4 * j], dy = py - obst[4 * j + 1];
        double d = sqrt(pow2_ref(dx) + pow2_ref(dy));
#ifdef ORC_POW_AUDIT
        double d2 = sqrt(pow(dx, 2.0) + pow(dy, 2.0));     /* np.sqrt((..)**2 + (..)**2), src/map.py:44 */
        __atomic_fetch_add(&g_pow_audit[3], 1, __ATOMIC_RELAXED);
        if (d2 != d) __atomic_fetch_add(&g_pow_audit[4], 1, __ATOMIC_RELAXED);
        if ((d2 < node_radius + obst[4 * j + 2]) != (d < node_radius + obst[4 * j + 2])) __atomic_fetch_add(&g_pow_audit[5], 1, __ATOMIC_RELAXED);
#endif
        if (d < node_radius + obst[4 * j + 2]) return 1;
    }
    return 0;
}

/* Graph.get_cost (src/graph.py:34-46): leaf -> root, sequential adds; `first` replaces cost[from] when >= 0
 * is requested through use_first (choose_parent's mutated G.cost[id_new], src/algorithm.py:849) */
static double chain_cost(orc_tree_t* t, int from) {
    double s = 0.0;
    int n = from;
    while (t->parent[n] >= 0) {
        s = s + t->ecost[n];
        n = t->parent[n];
        ++t->n_chain_hops;
    }
    return s;
}

/* ------------------------------------------------------------------------------------ tree API */

orc_tree_t* orc_create(int id_capacity, int finish_cap, int path_cap, double sx, double sy) {
    orc_tree_t* t = (orc_tree_t*)calloc(1, sizeof(orc_tree_t));
    t->cap = id_capacity;
    t->x = (double*)calloc(id_capacity, sizeof(double));
    t->y = (double*)calloc(id_capacity, sizeof(double));
    t->ecost = (double*)calloc(id_capacity, sizeof(double));
    t->parent = (int*)calloc(id_capacity, sizeof(int));
    t->nchild = (int*)calloc(id_capacity, sizeof(int));
    t->alive = (unsigned char*)calloc(id_capacity, 1);
    t->cand = (int*)calloc(id_capacity, sizeof(int));
    t->cand_d = (double*)calloc(id_capacity, sizeof(double));
    t->cand_hit = (unsigned char*)calloc(id_capacity, 1);
    t->finish_cap = finish_cap;
    t->finish = (int*)calloc(finish_cap > 0 ? finish_cap : 1, sizeof(int));
    t->path_cap = path_cap;
    t->best_path = (int*)calloc(path_cap > 0 ? path_cap : 1, sizeof(int));
    t->best_cost = INFINITY;
    /* Graph.__init__ (src/graph.py:25-32) */
    t->x[0] = sx; t->y[0] = sy; t->parent[0] = -1; t->ecost[0] = 0.0; t->alive[0] = 1;
    t->last_id = 0; t->n_live = 1; t->root = 0; t->fn_count = 1;
    return t;
}

void orc_destroy(orc_tree_t* t) {
    if (!t) return;
    free(t->x); free(t->y); free(t->ecost); free(t->parent); free(t->nchild); free(t->alive);
    free(t->cand); free(t->cand_d); free(t->cand_hit); free(t->finish); free(t->best_path);
    free(t);
}

/* Load an arbitrary tree (used by the FND tests): arrays by ascending id. parent = id or -1. */
int orc_load(orc_tree_t* t, int n, const int* ids, const double* x, const double* y, const int* parent,
             const double* ecost, int root_id, int last_id) {
    memset(t->alive, 0, t->cap);
    memset(t->nchild, 0, t->cap * sizeof(int));
    for (int i = 0; i < n; ++i) {
        int id = ids[i];
        if (id < 0 || id >= t->cap) return -1;
        t->alive[id] = 1; t->x[id] = x[i]; t->y[id] = y[i]; t->parent[id] = parent[i]; t->ecost[id] = ecost[i];
    }
    for (int i = 0; i < n; ++i) if (parent[i] >= 0) t->nchild[parent[i]]++;
    t->n_live = n; t->root = root_id; t->last_id = last_id; t->fn_count = 1;   /* a loaded tree = a new planner call */
    return 0;
}

/* Export live nodes in ascending id. Returns the number written. */
int orc_export(const orc_tree_t* t, int* ids, double* x, double* y, int* parent, double* ecost, int* nchild) {
    int n = 0;
    for (int i = 0; i <= t->last_id; ++i) {
        if (!t->alive[i]) continue;
        ids[n] = i; x[n] = t->x[i]; y[n] = t->y[i]; parent[n] = t->parent[i]; ecost[n] = t->ecost[i];
        nchild[n] = t->nchild[i];
        ++n;
    }
    return n;
}

void orc_get_planner(const orc_tree_t* t, rrtfnd_planner_t* p) {
    p->best_cost = t->best_cost; p->cursor = t->cursor; p->attempts = t->attempts;
    p->n_edge_checks = t->n_edge_checks; p->n_edge_checks_ref = t->n_edge_checks_ref;
    p->n_circle_tests = t->n_circle_tests; p->n_scan_nodes = t->n_scan_nodes; p->n_chain_hops = t->n_chain_hops;
    p->n_slots = t->last_id + 1; p->n_live = t->n_live; p->last_id = t->last_id; p->iter = t->iter;
    p->root_slot = t->root; p->n_finish = t->n_finish; p->best_len = t->best_len; p->status = t->status;
    p->n_rewired = t->n_rewired; p->n_removed = t->n_removed; p->solution_found = t->solution_found;
    p->n_compactions = 0; p->fn_count = t->fn_count; p->reserved = 0;
}

int orc_get_best_path(const orc_tree_t* t, int* out) {
    for (int i = 0; i < t->best_len; ++i) out[i] = t->best_path[i];
    return t->best_len;
}

int orc_get_finish(const orc_tree_t* t, int* out) {
    for (int i = 0; i < t->n_finish; ++i) out[i] = t->finish[i];
    return t->n_finish;
}

/* ---------------------------------------------------------------------------- single operators */

int orc_through_obstacle(double p1x, double p1y, double p2x, double p2y, const double* obst, int M) {
    return through_obstacle(p1x, p1y, p2x, p2y, obst, M, 0);
}

/* many segments at once: seg is [n][4] (p1x, p1y, p2x, p2y) */
void orc_through_obstacle_batch(const double* seg, int64_t n, const double* obst, int M, unsigned char* out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = (unsigned char)through_obstacle(seg[4 * i], seg[4 * i + 1], seg[4 * i + 2], seg[4 * i + 3], obst, M, 0);
}

int orc_is_occupied(double px, double py, const double* obst, int M, double node_radius) {
    return is_occupied(px, py, obst, M, node_radius);
}

double orc_get_cost(orc_tree_t* t, int id) { return chain_cost(t, id); }

double orc_dist_fma(double px, double py, double qx, double qy) { return dist_fma(px, py, qx, qy); }
double orc_dist_plain(double px, double py, double qx, double qy) { return dist_plain(px, py, qx, qy); }

/* nearest_node_kdtree (src/algorithm.py:666-699). Returns id >= 0, -1 none visible, -2 exact-position hit. */
int orc_nearest_visible(orc_tree_t* t, double qx, double qy, const double* obst, int M, int* rounds) {
    int n_rounds = 0;
    for (int i = 0; i <= t->last_id; ++i)    /* G.id_vertex[vertex] (src/algorithm.py:676-678) */
        if (t->alive[i] && t->x[i] == qx && t->y[i] == qy) { if (rounds) *rounds = 0; return -2; }
    double prev_d = -1.0;
    int prev_id = -1;
    for (int round = 0; round < t->n_live; ++round) {
        /* the (round+1)-th nearest in (d2, id) order == kdtree.query(k=nn)[-1] (:684-688) */
        double bd = INFINITY;
        int bi = -1;
        for (int i = 0; i <= t->last_id; ++i) {
            if (!t->alive[i]) continue;
            double d = dist2(t->x[i], t->y[i], qx, qy);
            ++t->n_scan_nodes;
            if (d < prev_d || (d == prev_d && i <= prev_id)) continue;
            if (d < bd) { bd = d; bi = i; }
        }
        if (bi < 0) break;
        ++n_rounds;
        ++t->n_edge_checks; ++t->n_edge_checks_ref;
        /* Line(vertex, closest_pos): p1 = query, p2 = node (:690) */
        if (!through_obstacle(qx, qy, t->x[bi], t->y[bi], obst, M, &t->n_circle_tests)) {
            if (rounds) *rounds = n_rounds;
            return bi;
        }
        prev_d = bd; prev_id = bi;
    }
    if (rounds) *rounds = n_rounds;
    return -1;
}

/* query_ball_point (src/algorithm.py:837/:893) without `skip_id`; ascending id. */
int orc_near_set(orc_tree_t* t, double qx, double qy, double radius, int skip_id, int* out) {
    int n = 0;
    double r2 = radius * radius;
    for (int i = 0; i <= t->last_id; ++i) {
        if (!t->alive[i]) continue;
        ++t->n_scan_nodes;
        if (i == skip_id) continue;
        if (dist2(t->x[i], t->y[i], qx, qy) <= r2) out[n++] = i;
    }
    return n;
}

/* Graph.remove_vertex (src/graph.py:64-77) */
static void remove_vertex(orc_tree_t* t, int id) {
    int p = t->parent[id];
    if (p >= 0) t->nchild[p]--;
    t->alive[id] = 0;
    t->n_live--;
}

/* find_path (src/algorithm.py:778-797): ids leaf -> root and the cost summed in that order */
static double find_path(orc_tree_t* t, int from, int* path, int* len) {
    double cost = 0.0;
    int n = from, L = 0;
    while (n != t->root) {
        if (n < 0 || !t->alive[n]) { *len = L; return cost; } /* swallowed exception (:794-795) */
        if (path && L < t->path_cap) path[L] = n;
        ++L;
        cost = cost + t->ecost[n];
        n = t->parent[n];
        ++t->n_chain_hops;
    }
    if (path && L < t->path_cap) path[L] = t->root;
    ++L;
    *len = L;
    return cost;
}


/* ------------------------------------------------------------------- FND tree surgery (row a15) */

static int in_path(const orc_tree_t* t, int id) {
    for (int i = 0; i < t->best_len; ++i) if (t->best_path[i] == id) return 1;
    return 0;
}

/* remove_children (src/algorithm.py:328-341): delete every descendant of `id` reached through children that are not
 * on the path (children on the path are kept and not descended into). Children are visited in ascending id. */
static void remove_children(orc_tree_t* t, int id) {
    for (int c = 0; c <= t->last_id; ++c) {
        if (!t->alive[c] || t->parent[c] != id || in_path(t, c)) continue;
        if (t->nchild[c] != 0) remove_children(t, c);
        remove_vertex(t, c);
    }
}

/* select_branch (src/algorithm.py:294-325) with path = the stored best path (ids goal side first).
 * Returns 0, or -1 if `current` is not on the path (the reference raises ValueError at path.index). */
int orc_fnd_select_branch(orc_tree_t* t, int current) {
    int idx = -1;
    for (int i = 0; i < t->best_len; ++i) if (t->best_path[i] == current) { idx = i; break; }
    if (idx < 0 || !t->alive[current]) return -1;
    int parent = t->parent[current];
    if (parent < 0) return -2;                    /* G.children[None] -> KeyError in the reference */
    t->parent[current] = -1;                      /* :305 */
    t->nchild[parent]--;                          /* :306-307 */
    for (int i = t->best_len - 1; i > idx; --i)   /* reversed(stranded_nodes): root side first (:315-317) */
        remove_children(t, t->best_path[i]);
    for (int i = idx + 1; i < t->best_len; ++i) { /* :322-323 */
        int n = t->best_path[i];
        if (t->parent[n] >= 0 && !t->alive[t->parent[n]]) t->parent[n] = -1;   /* its parent entry is already gone */
        remove_vertex(t, n);
    }
    t->root = current;                            /* G.start = G.vertices[current] (:303) */
    t->best_len = idx + 1;                        /* new_path = path[:idx + 1] (:313) */
    int len;
    t->best_cost = find_path(t, t->best_path[0], 0, &len);
    int k = 0;                                    /* planner-local finish list: keep the surviving entries */
    for (int f = 0; f < t->n_finish; ++f) if (t->alive[t->finish[f]]) t->finish[k++] = t->finish[f];
    t->n_finish = k;
    return 0;
}

/* valid_path (src/algorithm.py:344-367) over the stored path. Returns the separate root id (previous_root if no
 * path node is occupied); *index_error = 1 if the reference would raise IndexError at :358 (occupied goal-side leaf). */
int orc_fnd_valid_path(orc_tree_t* t, const double* obst, int M, double node_radius, int previous_root, int* index_error) {
    int sep = previous_root;
    *index_error = 0;
    for (int i = t->best_len - 1; i >= 0; --i) {  /* reversed(nodes_in_path): root side first (:355) */
        int id = t->best_path[i];
        if (!t->alive[id]) continue;
        if (!is_occupied(t->x[id], t->y[id], obst, M, node_radius)) continue;   /* :356 */
        remove_children(t, id);                   /* :357 */
        int child = -1;                           /* G.children[id][0]: the remaining (on-path) child (:358) */
        for (int c = 0; c <= t->last_id; ++c) if (t->alive[c] && t->parent[c] == id) { child = c; break; }
        if (child < 0) { *index_error = 1; return sep; }
        if (t->parent[id] >= 0 && !t->alive[t->parent[id]]) t->parent[id] = -1;
        remove_vertex(t, id);                     /* :359 */
        t->parent[child] = -1;                    /* :360 */
        sep = child;                              /* :361 */
    }
    return sep;
}

/* EXTENSION — invalidation of path EDGES (DetectCollision of the RRT*FND paper; the reference's valid_path tests the path nodes
 * only, src/algorithm.py:356, and its own loop is a stub, :283-291): PARITY UNPINNED BY THE REFERENCE. Composed from the
 * reference's through_obstacle (:581-591): after valid_path, every still-linked pair (child = path[i], parent = path[i + 1]) of
 * the stored path whose Line(parent, child) is blocked is cut - the child becomes a separate root, nothing is deleted. Returns
 * the root of the tree that holds the goal-side leaf if anything was cut (and it is not the planner's root), else sep_in. */
int orc_fnd_cut_blocked_edges(orc_tree_t* t, const double* obst, int M, int sep_in, int* n_cut) {
    int cuts = 0;
    for (int i = t->best_len - 2; i >= 0; --i) {
        int c = t->best_path[i], q = t->best_path[i + 1];
        if (c < 0 || q < 0 || !t->alive[c] || !t->alive[q] || t->parent[c] != q) continue;
        ++t->n_edge_checks; ++t->n_edge_checks_ref;
        if (!through_obstacle(t->x[q], t->y[q], t->x[c], t->y[c], obst, M, &t->n_circle_tests)) continue;
        t->parent[c] = -1;
        t->nchild[q]--;
        ++cuts;
    }
    if (n_cut) *n_cut = cuts;
    if (cuts == 0 || t->best_len == 0 || !t->alive[t->best_path[0]]) return sep_in;
    int n = t->best_path[0];
    while (t->parent[n] >= 0) n = t->parent[n];
    return n != t->root ? n : sep_in;
}

/* reconnect (src/algorithm.py:370-404) + get_near_nodes (:423-434) + get_distance_dict (:938-951, called with the
 * harness' default for its unused third argument) + check_for_tree_associativity (:407-420).
 * Returns the id the separate root was linked to, or -1; on success the stored best path / cost are recomputed from
 * last_node (find_path(G, last_node, id_root), :400). */
int orc_fnd_reconnect(orc_tree_t* t, const double* obst, int M, int sep, double radius, int last_node) {
    const double qx = t->x[sep], qy = t->y[sep];
    const double y2 = qx * qx + qy * qy;
    for (int i = 0; i <= t->last_id; ++i) {
        if (!t->alive[i]) continue;
        /* np.sqrt(x2 - 2 * matmul + y2): the BLAS dot of a 2-vector is fma(px, qx, fl(py*qy)) (probed) */
        double x2 = t->x[i] * t->x[i] + t->y[i] * t->y[i];
        double dot = fma(t->x[i], qx, t->y[i] * qy);
        double d = sqrt((x2 - 2.0 * dot) + y2);
        if (!(d <= radius)) continue;             /* :432 (NaN from cancellation is not <= radius) */
        int n = i;                                /* check_for_tree_associativity(G, root, i) */
        while (t->parent[n] >= 0) n = t->parent[n];
        if (n != t->root) continue;
        ++t->n_edge_checks; ++t->n_edge_checks_ref;
        if (through_obstacle(t->x[i], t->y[i], qx, qy, obst, M, &t->n_circle_tests)) continue;   /* Line(node, separate) :387-388 */
        t->parent[sep] = i;                       /* G.add_edge(id_separate_root, node, cost) :389-390 */
        t->nchild[i]++;
        t->ecost[sep] = dist_fma(t->x[i], t->y[i], qx, qy);
        if (last_node >= 0) t->best_cost = find_path(t, last_node, t->best_path, &t->best_len);
        return i;
    }
    return -1;
}


/* reconnect_ancestors (src/algorithm.py:531-543): make `id` the root of its tree by reversing the parent pointers on
 * its chain. Edge costs are NOT touched (the reference leaves G.cost as it was). Child COUNTS are kept consistent with
 * the parents here; the reference's children lists are not, so they are not compared for regrow. */
static void reconnect_ancestors(orc_tree_t* t, int id) {
    int node = id, parent = t->parent[node];
    if (parent < 0) return;
    while (t->parent[parent] >= 0) {
        int pp = t->parent[parent];
        t->parent[parent] = node; t->nchild[node]++; t->nchild[parent]--;
        node = parent; parent = pp;
    }
    t->parent[parent] = node; t->nchild[node]++; t->nchild[parent]--;
}

/* regrow (src/algorithm.py:482-528): grow the root tree with plain extensions (brute-force nearest_node :702-728 that
 * skips the separate tree; choose_parent's return value is dropped at :505) until a new node sees a separate-tree node
 * closer than step_length, then hang the separate tree (re-rooted at that node) under the new node. At most max_iter
 * (500 in the reference) extensions. Returns 1 if reconnected, 0 if not, < 0 = status. *iters_out = extensions made. */
int orc_fnd_regrow(orc_tree_t* t, const rrtfnd_params_t* prm, const double* obst, double gx, double gy, int sep_root,
                   const uint32_t* words, int64_t n_words, uint64_t seed, uint64_t stream_id, int max_iter, int* iters_out) {
    const int M = prm->n_obst;
    unsigned char* in_sep = (unsigned char*)calloc((size_t)t->cap, 1);
    for (int i = 0; i <= t->last_id; ++i) {              /* check_for_tree_associativity per node (:483) */
        if (!t->alive[i]) continue;
        int n = i;
        while (t->parent[n] >= 0) n = t->parent[n];
        in_sep[i] = (unsigned char)(n == sep_root);
    }
    orc_rng_t rng;
    memset(&rng, 0, sizeof rng);
    rng.mode = prm->stream_mode; rng.words = words; rng.n_words = n_words; rng.seed = seed; rng.stream = stream_id;
    rng.cursor = t->cursor;
    int iter = 0, reconnected = 0, status = 0;
    const int64_t attempt_cap = t->attempts + (int64_t)max_iter * RRTFND_REGROW_ATTEMPTS_PER_ITER;
    while (iter < max_iter && !reconnected) {
        if (t->attempts >= attempt_cap) { status = RRTFND_ST_ATTEMPT_CAP; break; }   /* the reference would spin forever here */
        if (rng.mode == RRTFND_STREAM_WORDS && (int64_t)rng.cursor + ORC_RESERVE_WORDS > n_words) { status = RRTFND_ST_REGROW_NEED_WORDS; break; }
        if (t->last_id + 1 >= t->cap) { status = RRTFND_ST_CAPACITY; break; }
        ++t->attempts;
        double qx, qy;
        if (rng_random(&rng) < prm->bias) { qx = gx; qy = gy; }
        else {
            qx = prm->x_lo + (prm->x_hi - prm->x_lo) * rng_random(&rng);
            qy = prm->y_lo + (prm->y_hi - prm->y_lo) * rng_random(&rng);
        }
        if (is_occupied(qx, qy, obst, M, prm->node_radius)) continue;           /* :493 */
        int exact = 0, best = -1;
        double bestd = INFINITY;
        for (int i = 0; i <= t->last_id; ++i) if (t->alive[i] && t->x[i] == qx && t->y[i] == qy) exact = 1;   /* :712-714 then :495 */
        if (exact) continue;
        for (int i = 0; i <= t->last_id; ++i) {                                   /* :718-726 */
            if (!t->alive[i] || in_sep[i]) continue;
            if (through_obstacle(t->x[i], t->y[i], qx, qy, obst, M, 0)) continue; /* Line(ver, vertex): p1 = node */
            double d = dist_fma(t->x[i], t->y[i], qx, qy);
            if (d < bestd) { bestd = d; best = i; }
        }
        if (best < 0) continue;
        double nx = t->x[best], ny = t->y[best];
        double d = dist_fma(qx, qy, nx, ny);                                      /* steer (:496) */
        double ux = (qx - nx) / d, uy = (qy - ny) / d;
        double sx = nx + ux * prm->step_length, sy = ny + uy * prm->step_length;
        double px = qx, py = qy;
        if (d > prm->step_length) { px = sx; py = sy; }
        if (is_occupied(px, py, obst, M, prm->node_radius)) continue;            /* :497 */
        for (int i = 0; i <= t->last_id; ++i)
            if (t->alive[i] && t->x[i] == px && t->y[i] == py) { status = RRTFND_ST_DUPLICATE; goto out; }
        int id_new = ++t->last_id;
        t->x[id_new] = px; t->y[id_new] = py; t->alive[id_new] = 1; t->nchild[id_new] = 0; in_sep[id_new] = 0;
        t->n_live++;
        t->parent[id_new] = best; t->ecost[id_new] = dist_fma(px, py, nx, ny);    /* :500-506 */
        t->nchild[best]++;
        ++iter;
        for (int i = 0; i < id_new; ++i) {                                        /* :510-518, separate tree in ascending id */
            if (!t->alive[i] || !in_sep[i]) continue;
            if (through_obstacle(px, py, t->x[i], t->y[i], obst, M, 0)) continue; /* Line(q_new, node): p1 = q_new */
            double dd = dist_fma(px, py, t->x[i], t->y[i]);
            if (dd < prm->step_length) {
                reconnect_ancestors(t, i);
                t->parent[i] = id_new; t->nchild[id_new]++; t->ecost[i] = dd;
                reconnected = 1;
                break;
            }
        }
    }
out:
    free(in_sep);
    t->cursor = rng.cursor;
    if (iters_out) *iters_out = iter;
    if (status) return status;
    if (reconnected && t->best_len > 0) t->best_cost = find_path(t, t->best_path[0], t->best_path, &t->best_len);
    return reconnected;
}

int orc_set_best_path(orc_tree_t* t, const int* ids, int n) {
    if (n > t->path_cap) return -1;
    memcpy(t->best_path, ids, (size_t)n * sizeof(int));
    t->best_len = n;
    return 0;
}
int orc_get_root(const orc_tree_t* t) { return t->root; }
void orc_set_cursor(orc_tree_t* t, uint64_t c) { t->cursor = c; }

/* --------------------------------------------------------------------------------- the planner */

/* One call = the body of RRT / RRT_star / RRT_star_FN (src/algorithm.py:54-265) run until
 * iter == prm->iter_num. obst is [M][4]. trace may be NULL. Returns status. */
int orc_run(orc_tree_t* t, const rrtfnd_params_t* prm, const double* obst, double gx, double gy,
            const uint32_t* words, int64_t n_words, uint64_t seed, uint64_t stream_id, rrtfnd_trace_t* trace) {
    const int M = prm->n_obst;
    orc_rng_t rng;
    memset(&rng, 0, sizeof rng);
    rng.mode = prm->stream_mode; rng.words = words; rng.n_words = n_words; rng.seed = seed; rng.stream = stream_id;
    rng.cursor = t->cursor;
    t->status = RRTFND_ST_RUNNING;
    const double r2 = prm->radius * prm->radius;

    while (t->iter < prm->iter_num) {
        if (prm->max_attempts > 0 && t->attempts >= prm->max_attempts) { t->status = RRTFND_ST_ATTEMPT_CAP; break; }
        if (rng.mode == RRTFND_STREAM_WORDS && (int64_t)rng.cursor + ORC_RESERVE_WORDS > n_words) {
            t->status = RRTFND_ST_NEED_WORDS; break;
        }
        if (t->last_id + 1 >= t->cap) { t->status = RRTFND_ST_CAPACITY; break; }
        ++t->attempts;

        /* ---- new_node (src/algorithm.py:36-51) */
        double qx, qy;
        if (rng_random(&rng) < prm->bias) {           /* Graph.random_node (src/graph.py:96-98) */
            qx = gx; qy = gy;
        } else {
            qx = prm->x_lo + (prm->x_hi - prm->x_lo) * rng_random(&rng);
            qy = prm->y_lo + (prm->y_hi - prm->y_lo) * rng_random(&rng);
        }
        if (is_occupied(qx, qy, obst, M, prm->node_radius)) continue;       /* :39 */
        int id_near = orc_nearest_visible(t, qx, qy, obst, M, 0);             /* :44 */
        if (id_near < 0) continue;                                            /* :45-46 */
        double nx = t->x[id_near], ny = t->y[id_near];
        /* steer (src/algorithm.py:731-746) */
        double d = dist_fma(qx, qy, nx, ny);
        double ux = (qx - nx) / d, uy = (qy - ny) / d;
        double sx = nx + ux * prm->step_length, sy = ny + uy * prm->step_length;
        double px = qx, py = qy;
        if (d > prm->step_length) { px = sx; py = sy; }
        /* Graph.add_vertex (src/graph.py:48-62) */
        for (int i = 0; i <= t->last_id; ++i)
            if (t->alive[i] && t->x[i] == px && t->y[i] == py) { t->status = RRTFND_ST_DUPLICATE; goto out; }
        int id_new = ++t->last_id;
        t->x[id_new] = px; t->y[id_new] = py; t->alive[id_new] = 1; t->nchild[id_new] = 0;
        double cost_new_near = dist_fma(px, py, nx, ny);                      /* :49 */
        t->n_live++; t->fn_count++;                                           /* n_of_nodes += 1 (:221) */
        t->parent[id_new] = id_near;                                          /* :141-142 */
        t->ecost[id_new] = cost_new_near;

        rrtfnd_trace_t tr;
        tr.id_new = id_new; tr.id_near = id_near; tr.n_near = 0; tr.n_rewired = 0; tr.removed_id = -1;
        tr.cp_parent = -1; tr.cp_cost = 0.0;

        if (prm->mode == RRTFND_MODE_RRT) {
            t->nchild[id_near]++;                                             /* add_edge (:76) */
            if (dist_fma(px, py, gx, gy) < 2.0 * prm->goal_node_radius) {          /* check_solution (:78) */
                t->best_cost = find_path(t, id_new, t->best_path, &t->best_len);
                if (t->best_len > t->path_cap) { t->status = RRTFND_ST_PATH_CAP; goto out; }
                t->solution_found = 1;
                if (trace && t->iter < prm->trace_cap) trace[t->iter] = tr;
                t->status = RRTFND_ST_DONE;                                   /* break before iter += 1 (:81) */
                goto out;
            }
            if (trace && t->iter < prm->trace_cap) trace[t->iter] = tr;
            t->iter++;
            continue;
        }

        /* ---- near set shared by choose_parent_kdtree and rewire_kdtree (:837, :893) */
        int k = orc_near_set(t, px, py, prm->radius, id_new, t->cand);
        (void)r2;
        tr.n_near = k;
        t->n_edge_checks_ref += 2; /* the ball holds q_new itself: Line(q_new, q_new) is tested at :847 and :904 */
        for (int c = 0; c < k; ++c) {
            int id = t->cand[c];
            t->cand_d[c] = dist_plain(t->x[id], t->y[id], px, py);            /* :842 */
            ++t->n_edge_checks; t->n_edge_checks_ref += 2;
            /* Line(vertex, q_new): p1 = node, p2 = q_new (:846, :903) */
            t->cand_hit[c] = (unsigned char)through_obstacle(t->x[id], t->y[id], px, py, obst, M, &t->n_circle_tests);
        }

        /* ---- choose_parent_kdtree (:845-852): value RETURNED by the function; the planners drop it */
        if (prm->flags & (RRTFND_FLAG_EVAL_CHOOSE_PARENT | RRTFND_FLAG_COMMIT_CHOOSE_PARENT)) {
            int bp = id_near;
            double bc = cost_new_near;
            for (int c = 0; c < k; ++c) {
                if (t->cand_hit[c]) continue;
                int id = t->cand[c];
                /* G.get_cost(id_new) with the current G.cost[id_new] but still id_near's chain (:848-849) */
                if (chain_cost(t, id_new) > chain_cost(t, id) + t->cand_d[c]) {
                    t->ecost[id_new] = t->cand_d[c];
                    bp = id; bc = t->cand_d[c];
                }
            }
            tr.cp_parent = bp; tr.cp_cost = bc;
            if (prm->flags & RRTFND_FLAG_COMMIT_CHOOSE_PARENT) {
                t->parent[id_new] = bp; t->ecost[id_new] = bc;
            } else {
                t->ecost[id_new] = cost_new_near;                             /* G.add_edge(*best_edge) (:148) */
            }
        }
        t->nchild[t->parent[id_new]]++;                                       /* add_edge (:148) */

        /* ---- rewire_kdtree (:902-910), sequential and live like the reference, ascending id */
        for (int c = 0; c < k; ++c) {
            if (t->cand_hit[c]) continue;
            int id = t->cand[c];
            if (chain_cost(t, id) > chain_cost(t, id_new) + t->cand_d[c]) {
                t->nchild[t->parent[id]]--;
                t->parent[id] = id_new;
                t->nchild[id_new]++;
                t->ecost[id] = t->cand_d[c];
                ++tr.n_rewired; ++t->n_rewired;
            }
        }

        /* ---- forced_removal (:233-237, :800-819) */
        if (prm->mode == RRTFND_MODE_FN && t->fn_count > prm->max_nodes) {   /* n_of_nodes > max_nodes (:233) */
            int last_in_path = t->best_len > 0 ? t->best_path[0] : -1;
            int n_childless = 0, n_ok = 0;
            for (int i = 0; i <= t->last_id; ++i)
                if (t->alive[i] && t->nchild[i] == 0) {
                    t->cand[n_childless++] = i;
                    if (i != id_new && i != last_in_path) ++n_ok;
                }
            if (n_ok == 0) { t->status = RRTFND_ST_NO_CHILDLESS; goto out; }
            int pick = t->cand[rng_randbelow(&rng, (uint32_t)n_childless)];
            while (pick == id_new || pick == last_in_path) pick = t->cand[rng_randbelow(&rng, (uint32_t)n_childless)];
            remove_vertex(t, pick);
            --t->fn_count;                                                    /* :237 */
            ++t->n_removed;
            tr.removed_id = pick;
            for (int f = 0; f < t->n_finish; ++f)
                if (t->finish[f] == pick) {                                   /* :235-236 */
                    memmove(&t->finish[f], &t->finish[f + 1], (t->n_finish - f - 1) * sizeof(int));
                    t->n_finish--;
                    break;
                }
        }

        /* ---- check_solution (:155-161, :240-243) */
        if (dist_fma(px, py, gx, gy) < 2.0 * prm->goal_node_radius) {
            if (t->n_finish >= t->finish_cap) { t->status = RRTFND_ST_FINISH_CAP; goto out; }
            t->finish[t->n_finish++] = id_new;
            t->solution_found = 1;
            if (prm->mode == RRTFND_MODE_STAR) {                              /* unconditional overwrite (:160-161) */
                t->best_cost = find_path(t, id_new, t->best_path, &t->best_len);
                if (t->best_len > t->path_cap) { t->status = RRTFND_ST_PATH_CAP; goto out; }
            }
        }
        /* ---- best-path refresh (:166-170, :247-251) */
        for (int f = 0; f < t->n_finish; ++f) {
            int len;
            double c = find_path(t, t->finish[f], 0, &len);
            if (c < t->best_cost) {
                t->best_cost = find_path(t, t->finish[f], t->best_path, &t->best_len);
                if (t->best_len > t->path_cap) { t->status = RRTFND_ST_PATH_CAP; goto out; }
            }
        }
        if (trace && t->iter < prm->trace_cap) trace[t->iter] = tr;
        t->iter++;
    }
    if (t->status == RRTFND_ST_RUNNING && t->iter >= prm->iter_num) t->status = RRTFND_ST_DONE;
out:
    if (rng.exhausted) t->status = RRTFND_ST_STREAM_EXHAUSTED;
    t->cursor = rng.cursor;
    return t->status;
}

/* Batch driver for the CPU baseline: P independent planners on n_threads pthreads (work queue over
 * planners). Each planner p uses start/goal sg[p] = (sx, sy, gx, gy), Philox stream (seeds[2p], seeds[2p+1]).
 * Fills planners_out[p]. Returns the number of planners that ended with an error status. */
typedef struct orc_job {
    const rrtfnd_params_t* prm;
    int P;
    const double* sg;
    const uint64_t* seeds;
    const double* obst;
    rrtfnd_planner_t* out;
    int next; /* atomic work counter */
    int bad;
} orc_job_t;

static void* orc_worker(void* arg) {
    orc_job_t* j = (orc_job_t*)arg;
    for (;;) {
        int p = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (p >= j->P) break;
        const double* sg = j->sg + 4 * p;
        /* ids are never reused (src/graph.py:57-58), so the id space is iter_num + 2 whatever the mode */
        orc_tree_t* t = orc_create(j->prm->iter_num + 2, j->prm->finish_cap, j->prm->path_cap, sg[0], sg[1]);
        int st = orc_run(t, j->prm, j->obst, sg[2], sg[3], 0, 0, j->seeds[2 * p], j->seeds[2 * p + 1], 0);
        if (st < 0) __atomic_fetch_add(&j->bad, 1, __ATOMIC_RELAXED);
        orc_get_planner(t, &j->out[p]);
        j->out[p].start_x = sg[0]; j->out[p].start_y = sg[1]; j->out[p].goal_x = sg[2]; j->out[p].goal_y = sg[3];
        j->out[p].seed = j->seeds[2 * p]; j->out[p].stream_id = j->seeds[2 * p + 1];
        orc_destroy(t);
    }
    return 0;
}

int orc_run_batch(const rrtfnd_params_t* prm, int P, const double* sg, const uint64_t* seeds, const double* obst,
                  rrtfnd_planner_t* planners_out, int n_threads) {
    orc_job_t job = {prm, P, sg, seeds, obst, planners_out, 0, 0};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], 0, orc_worker, &job);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
    return job.bad;
}
