// planner.cu — fused persistent planner kernel: one CTA grows one tree from sample to best-path bookkeeping.
//
// Replaces the Python loops of RRT / RRT_star / RRT_star_FN (src/algorithm.py:54-265) together with
// new_node (:36-51), nearest_node_kdtree (:666-699), steer (:731-746), choose_parent_kdtree (:822-852),
// rewire_kdtree (:882-910), forced_removal (:800-819), check_solution (:749-760), find_path (:778-797)
// and the Graph bookkeeping of src/graph.py:34-87.
//
// Layout. Node coordinates of the planner live in shared memory (x[], y[] fp64, the two arrays every
// nearest / near-set scan streams) and are mirrored to HBM on insert; parent[], ecost[], nchild[], id[]
// stay in HBM/L1 (touched only by chain walks and commits). The obstacle table is staged once per CTA as
// SoA in shared memory. Slots are kept in ascending-id order: inserts append, FN removals leave a
// tombstone (x = y = NaN so every scan predicate is false, nchild = id = -1) and a stable compaction runs
// only when the slot array is full.
//
// The grid has one CTA per planner; the hardware block scheduler is the work queue (planners are
// independent and never communicate), so a batch larger than the resident-CTA count streams through.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "rng.cuh"

namespace rrtfnd {

constexpr int kReserveWords = 64;   // must equal RESERVE_WORDS in _structs.py / ORC_RESERVE_WORDS
constexpr unsigned kFull = 0xffffffffu;

struct Ctl {
    double qx, qy;          // current sample
    double px, py;          // new node position
    double e0;              // cost_new_near = calc_distance(q_new, q_near)
    double cost_new;        // Graph.get_cost(id_new)
    double best_cost;
    double gx, gy;
    long long attempts, n_edge_checks, n_edge_checks_ref, n_circle_tests, n_scan_nodes, n_chain_hops;
    int n_slots, n_live, last_id, iter, root_slot, n_finish, best_len, status;
    int n_rewired, n_removed, solution_found, n_compactions;
    int best_leaf_slot;     // slot of best_path["path"][0], -1 if none
    int go, need_compact, near_slot, new_slot, k, reached, rewired_iter, dup, pick;
    int cp_parent_id;
    double cp_cost;
    int removed_id;
};

struct SmemPlan {
    size_t off_x, off_y, off_obst, off_bm, off_cslot, off_cdpl, off_ccost, off_chit, off_red_d, off_red_s,
        off_map, off_ctl, total;
    int bm_words;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline SmemPlan plan_smem(const rrtfnd_params_t& prm, int nt, bool tree_in_smem, bool obst_in_smem = true) {
    SmemPlan s;
    size_t o = 0;
    const size_t C = (size_t)prm.capacity;
    s.off_x = o; o += tree_in_smem ? align_up(C * 8, 16) : 0;
    s.off_y = o; o += tree_in_smem ? align_up(C * 8, 16) : 0;
    s.off_obst = o; o += obst_in_smem ? align_up((size_t)prm.n_obst * 8, 16) * 4 : 0;   // ox | oy | r | K
    s.bm_words = (int)((C + nt) / 32 + 1);
    s.off_bm = o; o += align_up((size_t)s.bm_words * 4, 16);
    s.off_cdpl = o; o += align_up((size_t)prm.near_cap * 8, 16);
    s.off_ccost = o; o += align_up((size_t)prm.near_cap * 8, 16);
    s.off_cslot = o; o += align_up((size_t)prm.near_cap * 4, 16);
    s.off_chit = o; o += align_up((size_t)prm.near_cap * 4, 16);
    s.off_red_d = o; o += align_up((size_t)2 * 32 * 8, 16);
    s.off_red_s = o; o += align_up((size_t)2 * 32 * 4, 16);
    s.off_map = o; o += (prm.mode == RRTFND_MODE_FN) ? align_up(C * 4, 16) : 0;
    s.off_ctl = o; o += align_up(sizeof(Ctl), 16);
    s.total = o;
    return s;
}

// Leaf -> root sequential sum of edge costs (Graph.get_cost, src/graph.py:41-46). `acc` is the running
// value to start from (0 for get_cost; the new node's own edge cost when walking from its parent).
__device__ __forceinline__ double chain_walk(const int32_t* __restrict__ parent, const double* __restrict__ ecost,
                                             int slot, double acc, long long& hops) {
    int n = slot;
    int p = parent[n];
    while (p >= 0) {
        acc = dadd(acc, ecost[n]);
        n = p;
        p = parent[n];
        ++hops;
    }
    return acc;
}

// kTreeInSmem: x[]/y[] of the planner staged in shared memory (else streamed from HBM/L2 through L1).
// kObstInSmem: obstacle table staged per CTA as SoA (else read as 32-byte rows through L1; the table is shared
//              by every planner, so with many small CTAs per SM one L1-resident copy serves them all).
template <int NT, bool kTreeInSmem, bool kObstInSmem>
__global__ void __launch_bounds__(NT) plan_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x;
    const int C = prm.capacity;
    const int M = prm.n_obst;
    const SmemPlan sp = plan_smem(prm, NT, kTreeInSmem, kObstInSmem);

    // per-planner global arrays
    double* __restrict__ GX = bt.x + (size_t)p * C;
    double* __restrict__ GY = bt.y + (size_t)p * C;
    double* __restrict__ ECOST = bt.ecost + (size_t)p * C;
    int32_t* __restrict__ PARENT = bt.parent + (size_t)p * C;
    int32_t* __restrict__ NCHILD = bt.nchild + (size_t)p * C;
    int32_t* __restrict__ ID = bt.id + (size_t)p * C;
    int32_t* __restrict__ FINISH = bt.finish + (size_t)p * prm.finish_cap;
    int32_t* __restrict__ BPATH = bt.best_path + (size_t)p * prm.path_cap;
    rrtfnd_planner_t* __restrict__ PL = bt.planners + p;
    rrtfnd_trace_t* __restrict__ TRACE = bt.trace ? bt.trace + (size_t)p * prm.trace_cap : nullptr;

    double* sx = kTreeInSmem ? reinterpret_cast<double*>(smem + sp.off_x) : GX;
    double* sy = kTreeInSmem ? reinterpret_cast<double*>(smem + sp.off_y) : GY;
    const size_t ostride = align_up((size_t)M * 8, 16) / 8;
    double* so_x = reinterpret_cast<double*>(smem + sp.off_obst);
    double* so_y = so_x + ostride;
    double* so_r = so_y + ostride;
    double* so_K = so_r + ostride;
    const double2* __restrict__ gob = reinterpret_cast<const double2*>(bt.obst);
    auto load_obst = [&](int j, double& ox, double& oy, double& orad, double& oK) {
        if (kObstInSmem) { ox = so_x[j]; oy = so_y[j]; orad = so_r[j]; oK = so_K[j]; }
        else { const double2 a = __ldg(gob + 2 * j), b = __ldg(gob + 2 * j + 1); ox = a.x; oy = a.y; orad = b.x; oK = b.y; }
    };
    uint32_t* bm = reinterpret_cast<uint32_t*>(smem + sp.off_bm);
    int* cslot = reinterpret_cast<int*>(smem + sp.off_cslot);
    double* cdpl = reinterpret_cast<double*>(smem + sp.off_cdpl);
    double* ccost = reinterpret_cast<double*>(smem + sp.off_ccost);
    int* chit = reinterpret_cast<int*>(smem + sp.off_chit);
    double* red_d = reinterpret_cast<double*>(smem + sp.off_red_d);
    int* red_s = reinterpret_cast<int*>(smem + sp.off_red_s);
    int* smap = reinterpret_cast<int*>(smem + sp.off_map);
    Ctl& c = *reinterpret_cast<Ctl*>(smem + sp.off_ctl);

    // ---- load planner state
    WordSrc rng;  // meaningful in thread 0 only
    if (tid == 0) {
        c.gx = PL->goal_x; c.gy = PL->goal_y;
        c.best_cost = PL->best_cost;
        c.attempts = PL->attempts; c.n_edge_checks = PL->n_edge_checks; c.n_edge_checks_ref = PL->n_edge_checks_ref;
        c.n_circle_tests = PL->n_circle_tests; c.n_scan_nodes = PL->n_scan_nodes; c.n_chain_hops = PL->n_chain_hops;
        c.n_slots = PL->n_slots; c.n_live = PL->n_live; c.last_id = PL->last_id; c.iter = PL->iter;
        c.root_slot = PL->root_slot; c.n_finish = PL->n_finish; c.best_len = PL->best_len;
        c.status = RRTFND_ST_RUNNING;
        c.n_rewired = PL->n_rewired; c.n_removed = PL->n_removed; c.solution_found = PL->solution_found;
        c.n_compactions = PL->n_compactions;
        c.best_leaf_slot = -1;
        rng.mode = prm.stream_mode;
        rng.words = bt.words ? bt.words + (size_t)p * prm.words_per_planner : nullptr;
        rng.n_words = prm.words_per_planner;
        rng.seed = PL->seed; rng.stream = PL->stream_id; rng.cursor = PL->cursor;
        rng.have_blk = false; rng.exhausted = false; rng.blk_idx = 0;
    }
    if (kObstInSmem) {
        for (int j = tid; j < M; j += NT) {
            so_x[j] = bt.obst[4 * j + 0]; so_y[j] = bt.obst[4 * j + 1];
            so_r[j] = bt.obst[4 * j + 2]; so_K[j] = bt.obst[4 * j + 3];
        }
    }
    __syncthreads();
    if (kTreeInSmem) {
        for (int s = tid; s < c.n_slots; s += NT) { sx[s] = GX[s]; sy[s] = GY[s]; }
    }
    if (tid == 0 && c.best_len > 0) {
        // slot of the cached best path's leaf (ids ascend with slots: binary search)
        const int leaf_id = BPATH[0];
        int lo = 0, hi = c.n_slots - 1, found = -1;
        while (lo <= hi) {  // tombstones carry id -1: fall back to a linear probe around them
            const int mid = (lo + hi) >> 1;
            int m2 = mid;
            while (m2 <= hi && ID[m2] < 0) ++m2;
            if (m2 > hi) { hi = mid - 1; continue; }
            if (ID[m2] == leaf_id) { found = m2; break; }
            if (ID[m2] < leaf_id) lo = m2 + 1; else hi = mid - 1;
        }
        c.best_leaf_slot = found;
    }
    __syncthreads();

    long long my_hops = 0, my_tests = 0;   // per-thread counters, merged at the end
    const double r2 = dmul(prm.radius, prm.radius);
    const double goal_thresh = dmul(2.0, prm.node_radius);

    for (;;) {
        // ================================================================ A. attempt header (thread 0)
        if (tid == 0) {
            c.go = 1; c.need_compact = 0;
            if (c.iter >= prm.iter_num) { c.status = RRTFND_ST_DONE; c.go = 0; }
            else if (prm.max_attempts > 0 && c.attempts >= prm.max_attempts) { c.status = RRTFND_ST_ATTEMPT_CAP; c.go = 0; }
            else if (prm.stream_mode == RRTFND_STREAM_WORDS &&
                     (long long)rng.cursor + kReserveWords > prm.words_per_planner) { c.status = RRTFND_ST_NEED_WORDS; c.go = 0; }
            else if (c.n_slots >= C) {
                if (c.n_live < c.n_slots) c.need_compact = 1;
                else { c.status = RRTFND_ST_CAPACITY; c.go = 0; }
            }
            if (c.go) {
                ++c.attempts;
                // Graph.random_node (src/graph.py:96-98)
                if (next_random(rng) < prm.bias) { c.qx = c.gx; c.qy = c.gy; }
                else {
                    c.qx = dadd(prm.x_lo, dmul(dsub(prm.x_hi, prm.x_lo), next_random(rng)));
                    c.qy = dadd(prm.y_lo, dmul(dsub(prm.y_hi, prm.y_lo), next_random(rng)));
                }
            }
        }
        __syncthreads();
        if (!c.go) break;

        // ================================================================ A'. stable compaction (FN, rare)
        if (c.need_compact) {
            const int n_slots = c.n_slots;
            if (warp == 0) {
                int running = 0;
                for (int base = 0; base < n_slots; base += 32) {
                    const int s = base + lane;
                    const bool alive = s < n_slots && ID[s] >= 0;
                    const unsigned b = __ballot_sync(kFull, alive);
                    if (s < n_slots) smap[s] = alive ? running + __popc(b & ((1u << lane) - 1u)) : -1;
                    running += __popc(b);
                }
            }
            __syncthreads();
            for (int base = 0; base < n_slots; base += NT) {
                const int s = base + tid;
                const bool alive = s < n_slots && smap[s] >= 0;
                double vx = 0, vy = 0, ve = 0; int vp = -1, vn = 0, vi = -1;
                if (alive) {
                    vx = sx[s]; vy = sy[s]; ve = ECOST[s]; vp = PARENT[s]; vn = NCHILD[s]; vi = ID[s];
                    if (vp >= 0) vp = smap[vp];
                }
                __syncthreads();
                if (alive) {
                    const int d = smap[s];
                    sx[d] = vx; sy[d] = vy;
                    if (kTreeInSmem) { GX[d] = vx; GY[d] = vy; }
                    ECOST[d] = ve; PARENT[d] = vp; NCHILD[d] = vn; ID[d] = vi;
                }
                __syncthreads();
            }
            for (int f = tid; f < c.n_finish; f += NT) FINISH[f] = smap[FINISH[f]];
            __syncthreads();
            if (tid == 0) {
                c.root_slot = smap[c.root_slot];
                if (c.best_leaf_slot >= 0) c.best_leaf_slot = smap[c.best_leaf_slot];
                c.n_slots = c.n_live;
                ++c.n_compactions;
            }
            __syncthreads();
        }

        const double qx = c.qx, qy = c.qy;
        const int n_slots = c.n_slots;

        // ================================================================ B. Map.is_occupied_c (src/map.py:38-47)
        {
            bool occ = false;
            for (int j = tid; j < M; j += NT) {
                double ox, oy, orad, oK;
                load_obst(j, ox, oy, orad, oK);
                occ |= point_in_inflated(qx, qy, ox, oy, orad, prm.node_radius);
            }
            if (__syncthreads_or(occ)) continue;                         // src/algorithm.py:39-40
        }

        // ================================================================ C. nearest visible node (:666-699)
        int near_slot = -1;
        {
            double prev_d = -1.0; int prev_s = -1;
            bool reject = false;
            for (int round = 0; round < c.n_live; ++round) {
                double bd = CUDART_INF; int bs = 0x7fffffff; bool exact = false;
                for (int s = tid; s < n_slots; s += NT) {
                    const double x = sx[s], y = sy[s];
                    const double d = dist2(x, y, qx, qy);
                    if (round == 0 && x == qx && y == qy) exact = true;           // G.id_vertex[vertex] hit (:676-678)
                    const bool after_prev = d > prev_d || (d == prev_d && s > prev_s);
                    if (after_prev && d < bd) { bd = d; bs = s; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(kFull, bd, off);
                    const int os = __shfl_xor_sync(kFull, bs, off);
                    if (key_less(od, os, bd, bs)) { bd = od; bs = os; }
                }
                const int par = round & 1;
                if (lane == 0) { red_d[par * 32 + warp] = bd; red_s[par * 32 + warp] = bs; }
                if (round == 0) { if (__syncthreads_or(exact)) { reject = true; break; } }
                else __syncthreads();
                bd = red_d[par * 32]; bs = red_s[par * 32];
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    const double od = red_d[par * 32 + w]; const int os = red_s[par * 32 + w];
                    if (key_less(od, os, bd, bs)) { bd = od; bs = os; }
                }
                if (bs == 0x7fffffff) break;                                      // nothing left
                // through_obstacle(Line(q_rand, node)) — p1 = sample, p2 = node (:690-692)
                const Seg sg = make_seg(qx, qy, sx[bs], sy[bs]);
                bool hit = false;
                for (int j = tid; j < M; j += NT) {
                    double ox, oy, orad, oK;
                    load_obst(j, ox, oy, orad, oK);
                    hit |= seg_hits_circle(sg, ox, oy, orad, oK);
                }
                if (tid == 0) { ++c.n_edge_checks; ++c.n_edge_checks_ref; c.n_circle_tests += M; c.n_scan_nodes += c.n_live; }
                if (!__syncthreads_or(hit)) { near_slot = bs; break; }
                prev_d = bd; prev_s = bs;
            }
            if (reject || near_slot < 0) continue;                               // :45-46
        }

        // ================================================================ D. steer + insert (thread 0)
        if (tid == 0) {
            const double nx = sx[near_slot], ny = sy[near_slot];
            const double d = dist_fma(qx, qy, nx, ny);                            // steer (:739)
            const double ux = ddiv(dsub(qx, nx), d), uy = ddiv(dsub(qy, ny), d);
            const double tx = dadd(nx, dmul(ux, prm.step_length)), ty = dadd(ny, dmul(uy, prm.step_length));
            double px = qx, py = qy;
            if (d > prm.step_length) { px = tx; py = ty; }
            const int slot = c.n_slots++;
            const int id = ++c.last_id;                                           // Graph.add_vertex (src/graph.py:57-61)
            ++c.n_live;
            const double e0 = dist_fma(px, py, nx, ny);                           // :49
            sx[slot] = px; sy[slot] = py;
            if (kTreeInSmem) { GX[slot] = px; GY[slot] = py; }
            ID[slot] = id; PARENT[slot] = near_slot; ECOST[slot] = e0; NCHILD[slot] = 0;  // :141-142
            c.px = px; c.py = py; c.e0 = e0; c.near_slot = near_slot; c.new_slot = slot;
            c.reached = dist_fma(px, py, c.gx, c.gy) < goal_thresh;               // check_solution (:757-758)
            c.k = 0; c.rewired_iter = 0; c.dup = 0; c.removed_id = -1; c.cp_parent_id = -1; c.cp_cost = 0.0;
            if (prm.mode == RRTFND_MODE_RRT) {
                NCHILD[near_slot] += 1;                                           // add_edge (:76)
                if (c.reached) {                                                  // :78-81, break before iter += 1
                    double cost = 0.0; int n = slot, L = 0;
                    while (n != c.root_slot && n >= 0) {
                        if (L < prm.path_cap) BPATH[L] = ID[n];
                        ++L; cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++my_hops;
                    }
                    if (L < prm.path_cap) BPATH[L] = ID[c.root_slot];
                    ++L;
                    c.best_len = L; c.best_cost = cost; c.best_leaf_slot = slot; c.solution_found = 1;
                    c.status = L > prm.path_cap ? RRTFND_ST_PATH_CAP : RRTFND_ST_DONE;
                    c.go = 0;
                }
                if (TRACE && c.iter < prm.trace_cap) {
                    rrtfnd_trace_t t; t.id_new = id; t.id_near = ID[near_slot]; t.n_near = 0; t.n_rewired = 0;
                    t.removed_id = -1; t.cp_parent = -1; t.cp_cost = 0.0;
                    TRACE[c.iter] = t;
                }
                if (c.go) ++c.iter;
            }
        }
        __syncthreads();
        if (prm.mode == RRTFND_MODE_RRT) { if (!c.go) break; continue; }

        const double px = c.px, py = c.py;
        const int new_slot = c.new_slot;
        const int n_slots2 = c.n_slots;

        // ================================================================ E. near set (query_ball_point, :837 / :893)
        {
            bool dup = false;
            for (int base = 0; base < n_slots2; base += NT) {
                const int s = base + tid;
                bool in = false;
                if (s < n_slots2 && s != new_slot) {
                    const double x = sx[s], y = sy[s];
                    in = dist2(x, y, px, py) <= r2;
                    dup |= (x == px && y == py);
                }
                const unsigned b = __ballot_sync(kFull, in);
                if (lane == 0) bm[(base >> 5) + warp] = b;
            }
            if (__syncthreads_or(dup)) {                                          // src/graph.py:54-56 corner (B9)
                if (tid == 0) c.status = RRTFND_ST_DUPLICATE;
                break;
            }
            if (warp == 0) {                                                      // bitmap -> ascending slot list
                const int nwords = (n_slots2 + 31) >> 5;
                int count = 0;
                for (int w0 = 0; w0 < nwords; w0 += 32) {
                    uint32_t w = (w0 + lane < nwords) ? bm[w0 + lane] : 0u;
                    const int n = __popc(w);
                    int incl = n;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int v = __shfl_up_sync(kFull, incl, off);
                        if (lane >= off) incl += v;
                    }
                    int pos = count + incl - n;
                    while (w) {
                        const int bit = __ffs(w) - 1;
                        w &= w - 1;
                        if (pos < prm.near_cap) cslot[pos] = ((w0 + lane) << 5) + bit;
                        ++pos;
                    }
                    count += __shfl_sync(kFull, incl, 31);
                }
                if (lane == 0) {
                    c.k = count;
                    c.n_scan_nodes += c.n_live;
                    c.n_edge_checks += count;
                    c.n_edge_checks_ref += 2 * (long long)count + 2;   // each candidate twice + Line(q_new,q_new) twice
                }
            }
            __syncthreads();
            if (c.k > prm.near_cap) { if (tid == 0) c.status = RRTFND_ST_NEAR_CAP; break; }
        }
        const int k = c.k;

        // ---- E1. collision of Line(node, q_new) per candidate (:846-847 / :903-904): one warp per edge
        for (int ci = warp; ci < k; ci += NW) {
            const int s = cslot[ci];
            const double x = sx[s], y = sy[s];
            const Seg sg = make_seg(x, y, px, py);                                // p1 = node, p2 = q_new
            bool hit = false;
            int j0 = 0;
            for (; j0 < M; j0 += 32) {
                const int j = j0 + lane;
                bool h = false;
                if (j < M) {
                    double ox, oy, orad, oK;
                    load_obst(j, ox, oy, orad, oK);
                    h = seg_hits_circle(sg, ox, oy, orad, oK);
                }
                if (__any_sync(kFull, h)) { hit = true; j0 += 32; break; }
            }
            if (lane == 0) {
                chit[ci] = hit ? 1 : 0;
                cdpl[ci] = dist_plain(x, y, px, py);                              // :842 / :899
                my_tests += (j0 < M ? j0 : M);
            }
        }
        // ---- E2. Graph.get_cost per candidate (src/graph.py:34-46): one lane per chain
        for (int ci = tid; ci < k; ci += NT) ccost[ci] = chain_walk(PARENT, ECOST, cslot[ci], 0.0, my_hops);
        if (tid == NT - 1) c.cost_new = chain_walk(PARENT, ECOST, c.near_slot, c.e0, my_hops);  // get_cost(id_new)
        __syncthreads();

        // ---- E3. choose_parent_kdtree's RETURNED edge (:845-852); dropped by the planners (:146, :224)
        if (prm.flags & (RRTFND_FLAG_EVAL_CHOOSE_PARENT | RRTFND_FLAG_COMMIT_CHOOSE_PARENT)) {
            if (tid == 0) {
                int bp = c.near_slot; double bc = c.e0; double cn = c.cost_new;
                for (int ci = 0; ci < k; ++ci) {
                    if (chit[ci]) continue;
                    if (cn > dadd(ccost[ci], cdpl[ci])) {
                        bp = cslot[ci]; bc = cdpl[ci];
                        // G.cost[id_new] = cost, parent untouched: the chain is still id_near's (:849)
                        cn = chain_walk(PARENT, ECOST, c.near_slot, bc, my_hops);
                    }
                }
                c.cp_parent_id = ID[bp]; c.cp_cost = bc;
                if (prm.flags & RRTFND_FLAG_COMMIT_CHOOSE_PARENT) {   // EXTENSION (not the shipped behaviour)
                    PARENT[new_slot] = bp; ECOST[new_slot] = bc;
                    c.cost_new = chain_walk(PARENT, ECOST, bp, bc, my_hops);
                }
            }
            __syncthreads();
        }

        // ---- E4. rewire_kdtree (:902-910): decide against the snapshot, then commit
        {
            const double cost_new = c.cost_new;
            int n_rw = 0;
            for (int ci = tid; ci < k; ci += NT) {
                if (chit[ci]) continue;
                if (ccost[ci] > dadd(cost_new, cdpl[ci])) {
                    const int s = cslot[ci];
                    atomicSub(&NCHILD[PARENT[s]], 1);
                    PARENT[s] = new_slot;
                    ECOST[s] = cdpl[ci];
                    ++n_rw;
                }
            }
            if (n_rw) { atomicAdd(&NCHILD[new_slot], n_rw); atomicAdd(&c.rewired_iter, n_rw); }
            if (tid == 0) atomicAdd(&NCHILD[PARENT[new_slot]], 1);                // G.add_edge(*best_edge) (:148)
        }
        __syncthreads();

        // ================================================================ F. forced_removal (:233-237, :800-819)
        if (prm.mode == RRTFND_MODE_FN && c.n_live > prm.max_nodes) {
            for (int base = 0; base < n_slots2; base += NT) {
                const int s = base + tid;
                const bool childless = s < n_slots2 && NCHILD[s] == 0;             // tombstones hold -1
                const unsigned b = __ballot_sync(kFull, childless);
                if (lane == 0) bm[(base >> 5) + warp] = b;
            }
            __syncthreads();
            if (warp == 0) {
                const int nwords = (n_slots2 + 31) >> 5;
                int total = 0;
                for (int w0 = lane; w0 < nwords; w0 += 32) total += __popc(bm[w0]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(kFull, total, off);
                const int leaf = c.best_leaf_slot;
                int n_ok = total - (NCHILD[new_slot] == 0 ? 1 : 0);
                if (leaf >= 0 && leaf != new_slot && NCHILD[leaf] == 0) --n_ok;
                int pick = -1;
                if (n_ok <= 0) { if (lane == 0) c.status = RRTFND_ST_NO_CHILDLESS; }
                else {
                    for (;;) {
                        uint32_t r = 0; int bad = 0;
                        if (lane == 0) { r = next_randbelow(rng, (uint32_t)total); bad = rng.exhausted; }   // random.choice (:813/:816)
                        r = __shfl_sync(kFull, r, 0); bad = __shfl_sync(kFull, bad, 0);
                        if (bad) { if (lane == 0) c.status = RRTFND_ST_STREAM_EXHAUSTED; break; }
                        // r-th set bit of the bitmap, in ascending slot order
                        int seen = 0; pick = -1;
                        for (int w0 = 0; w0 < nwords && pick < 0; w0 += 32) {
                            const uint32_t w = (w0 + lane < nwords) ? bm[w0 + lane] : 0u;
                            const int n = __popc(w);
                            int incl = n;
#pragma unroll
                            for (int off = 1; off < 32; off <<= 1) {
                                const int v = __shfl_up_sync(kFull, incl, off);
                                if (lane >= off) incl += v;
                            }
                            const int excl = seen + incl - n;
                            const bool mine = (int)r >= excl && (int)r < excl + n;
                            const unsigned who = __ballot_sync(kFull, mine);
                            if (who) {
                                const int src = __ffs(who) - 1;
                                int cand = -1;
                                if (lane == src) cand = ((w0 + lane) << 5) + (__fns(w, 0, (int)r - excl + 1));
                                pick = __shfl_sync(kFull, cand, src);
                            }
                            seen += __shfl_sync(kFull, incl, 31);
                        }
                        if (pick != new_slot && pick != leaf) break;              // :815
                    }
                }
                if (lane == 0) { c.pick = (c.status == RRTFND_ST_RUNNING) ? pick : -1; c.k = 0; }
            }
            __syncthreads();
            if (c.status != RRTFND_ST_RUNNING) break;
            const int pick = c.pick;
            // finish_nodes_of_path.remove(id_removed) (:235-236), order preserving
            int fpos = -1;
            for (int f = tid; f < c.n_finish; f += NT) if (FINISH[f] == pick) fpos = f;
            if (fpos >= 0) c.k = fpos + 1;      // c.k doubles as a mailbox (zeroed above; candidates are done)
            __syncthreads();
            if (c.k > 0) {
                const int f0 = c.k - 1;
                // shift left by one, chunked so reads precede writes
                for (int base = f0; base < c.n_finish - 1; base += NT) {
                    const int f = base + tid;
                    int v = 0; const bool ok = f < c.n_finish - 1;
                    if (ok) v = FINISH[f + 1];
                    __syncthreads();
                    if (ok) FINISH[f] = v;
                    __syncthreads();
                }
            }
            if (tid == 0) {                                                       // Graph.remove_vertex (src/graph.py:64-77)
                if (c.k > 0) --c.n_finish;
                const int par = PARENT[pick];
                if (par >= 0) NCHILD[par] -= 1;
                c.removed_id = ID[pick];
                ID[pick] = -1; NCHILD[pick] = -1; PARENT[pick] = -1;
                const double qnan = CUDART_NAN;
                sx[pick] = qnan; sy[pick] = qnan;
                if (kTreeInSmem) { GX[pick] = qnan; GY[pick] = qnan; }
                --c.n_live; ++c.n_removed;
            }
            __syncthreads();
        }

        // ================================================================ G. goal test and best path (:155-170, :240-251)
        if (c.reached) {
            if (tid == 0) {
                if (c.n_finish >= prm.finish_cap) c.status = RRTFND_ST_FINISH_CAP;
                else {
                    FINISH[c.n_finish++] = new_slot;
                    c.solution_found = 1;
                    if (prm.mode == RRTFND_MODE_STAR) {                           // unconditional overwrite (:160-161)
                        double cost = 0.0; int n = new_slot, L = 0;
                        while (n != c.root_slot && n >= 0) {
                            if (L < prm.path_cap) BPATH[L] = ID[n];
                            ++L; cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++my_hops;
                        }
                        if (L < prm.path_cap) BPATH[L] = ID[c.root_slot];
                        ++L;
                        c.best_len = L; c.best_cost = cost; c.best_leaf_slot = new_slot;
                        if (L > prm.path_cap) c.status = RRTFND_ST_PATH_CAP;
                    }
                }
            }
            __syncthreads();
            if (c.status != RRTFND_ST_RUNNING) break;
        }
        // Path costs only change through a rewire, and the list only grows through a new finish node, so the
        // reference's per-iteration refresh (:166-170) can only have an effect in those two cases.
        if (c.reached || c.rewired_iter > 0) {
            const int nf = c.n_finish;
            if (nf > 0) {
                double bd = CUDART_INF; int bf = 0x7fffffff;
                for (int f = tid; f < nf; f += NT) {
                    // find_path's running sum starts at 0 and adds leaf -> root (:787-792)
                    double cost = 0.0; int n = FINISH[f];
                    while (n != c.root_slot && n >= 0) { cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++my_hops; }
                    if (cost < bd) { bd = cost; bf = f; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(kFull, bd, off);
                    const int of = __shfl_xor_sync(kFull, bf, off);
                    if (key_less(od, of, bd, bf)) { bd = od; bf = of; }
                }
                if (lane == 0) { red_d[warp] = bd; red_s[warp] = bf; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < NW; ++w)
                        if (key_less(red_d[w], red_s[w], bd, bf)) { bd = red_d[w]; bf = red_s[w]; }
                    if (bd < c.best_cost) {                                       // strict (:168, :249)
                        int n = FINISH[bf], L = 0;
                        const int leaf = n;
                        while (n != c.root_slot && n >= 0) {
                            if (L < prm.path_cap) BPATH[L] = ID[n];
                            ++L; n = PARENT[n]; ++my_hops;
                        }
                        if (L < prm.path_cap) BPATH[L] = ID[c.root_slot];
                        ++L;
                        c.best_len = L; c.best_cost = bd; c.best_leaf_slot = leaf;
                        if (L > prm.path_cap) c.status = RRTFND_ST_PATH_CAP;
                    }
                }
                __syncthreads();
                if (c.status != RRTFND_ST_RUNNING) break;
            }
        }

        // ================================================================ H. end of iteration
        if (tid == 0) {
            c.n_rewired += c.rewired_iter;
            if (TRACE && c.iter < prm.trace_cap) {
                rrtfnd_trace_t t;
                t.id_new = c.last_id; t.id_near = ID[c.near_slot]; t.n_near = k; t.n_rewired = c.rewired_iter;
                t.removed_id = c.removed_id; t.cp_parent = c.cp_parent_id; t.cp_cost = c.cp_cost;
                TRACE[c.iter] = t;
            }
            ++c.iter;                                                             // :173 / :254
        }
        // the next header's __syncthreads orders these writes before anyone reads them
    }

    // ---- write back
    __syncthreads();
    if (my_hops) atomicAdd((unsigned long long*)&c.n_chain_hops, (unsigned long long)my_hops);
    if (my_tests) atomicAdd((unsigned long long*)&c.n_circle_tests, (unsigned long long)my_tests);
    __syncthreads();
    if (tid == 0) {
        PL->best_cost = c.best_cost; PL->cursor = rng.cursor; PL->attempts = c.attempts;
        PL->n_edge_checks = c.n_edge_checks; PL->n_edge_checks_ref = c.n_edge_checks_ref;
        PL->n_circle_tests = c.n_circle_tests; PL->n_scan_nodes = c.n_scan_nodes; PL->n_chain_hops = c.n_chain_hops;
        PL->n_slots = c.n_slots; PL->n_live = c.n_live; PL->last_id = c.last_id; PL->iter = c.iter;
        PL->root_slot = c.root_slot; PL->n_finish = c.n_finish; PL->best_len = c.best_len; PL->status = c.status;
        PL->n_rewired = c.n_rewired; PL->n_removed = c.n_removed; PL->solution_found = c.solution_found;
        PL->n_compactions = c.n_compactions;
    }
}

// Graph.__init__ (src/graph.py:25-32): a tree holding only the start node, id 0.
__global__ void init_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const size_t o = (size_t)p * prm.capacity;
    rrtfnd_planner_t* PL = bt.planners + p;
    bt.x[o] = PL->start_x; bt.y[o] = PL->start_y; bt.ecost[o] = 0.0;
    bt.parent[o] = -1; bt.nchild[o] = 0; bt.id[o] = 0;
    PL->best_cost = CUDART_INF;
    PL->cursor = 0; PL->attempts = 0; PL->n_edge_checks = 0; PL->n_edge_checks_ref = 0; PL->n_circle_tests = 0;
    PL->n_scan_nodes = 0; PL->n_chain_hops = 0;
    PL->n_slots = 1; PL->n_live = 1; PL->last_id = 0; PL->iter = 0; PL->root_slot = 0; PL->n_finish = 0;
    PL->best_len = 0; PL->status = RRTFND_ST_RUNNING; PL->n_rewired = 0; PL->n_removed = 0;
    PL->solution_found = 0; PL->n_compactions = 0;
}

constexpr size_t kMaxSmem = 227 * 1024;

template <int NT, bool kTreeInSmem, bool kObstInSmem>
static int launch_plan(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, cudaStream_t st) {
    const SmemPlan sp = plan_smem(*prm, NT, kTreeInSmem, kObstInSmem);
    auto kern = plan_kernel<NT, kTreeInSmem, kObstInSmem>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total);
    if (e != cudaSuccess) return (int)e;
    kern<<<bt->n_planners, NT, sp.total, st>>>(*prm, *bt);
    return (int)cudaGetLastError();
}

// Launch geometry. prm->reserved encodes an explicit choice as threads + 1000 * layout
// (layout 0: tree + obstacles in shared memory, 1: obstacles only, 2: neither); 0 = automatic.
struct PlanCfg { int nt; int layout; };

static PlanCfg choose_cfg(const rrtfnd_params_t& prm, int n_planners) {
    if (prm.reserved > 0) return PlanCfg{prm.reserved % 1000, prm.reserved / 1000};
    // Batches that can fill the GPU several CTAs deep run light CTAs (tree streamed through L1/L2, obstacle table
    // L1-resident): occupancy hides the parent-chain latency. Small batches keep one fat CTA per planner with
    // the coordinates in shared memory (lowest latency per tree).
    if (n_planners >= 296 && plan_smem(prm, 128, false, false).total <= 24 * 1024) return PlanCfg{128, 2};
    if (plan_smem(prm, 256, true, true).total <= kMaxSmem) return PlanCfg{256, 0};
    if (plan_smem(prm, 256, false, true).total <= kMaxSmem) return PlanCfg{256, 1};
    return PlanCfg{256, 2};
}

static size_t cfg_smem(const rrtfnd_params_t& prm, PlanCfg c) {
    return plan_smem(prm, c.nt, c.layout == 0, c.layout <= 1).total;
}

template <int NT>
static int launch_nt(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, int layout, cudaStream_t st) {
    switch (layout) {
        case 0: return launch_plan<NT, true, true>(prm, bt, st);
        case 1: return launch_plan<NT, false, true>(prm, bt, st);
        case 2: return launch_plan<NT, false, false>(prm, bt, st);
        default: return RRTFND_E_BADARG;
    }
}

static int launch_cfg(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, PlanCfg c, cudaStream_t st) {
    if (cfg_smem(*prm, c) > kMaxSmem) return RRTFND_E_SMEM;
    switch (c.nt) {
        case 32: return launch_nt<32>(prm, bt, c.layout, st);
        case 64: return launch_nt<64>(prm, bt, c.layout, st);
        case 128: return launch_nt<128>(prm, bt, c.layout, st);
        case 256: return launch_nt<256>(prm, bt, c.layout, st);
        case 512: return launch_nt<512>(prm, bt, c.layout, st);
        default: return RRTFND_E_BADARG;
    }
}

}  // namespace rrtfnd

using namespace rrtfnd;

static int check_params(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt) {
    if (!prm || !bt) return RRTFND_E_BADARG;
    if (bt->n_planners <= 0 || prm->capacity < 2 || prm->n_obst < 0) return RRTFND_E_BADARG;
    if (prm->mode < RRTFND_MODE_RRT || prm->mode > RRTFND_MODE_FN) return RRTFND_E_BADARG;
    if (prm->near_cap < 1 || prm->finish_cap < 1 || prm->path_cap < 2) return RRTFND_E_BADARG;
    if (prm->stream_mode == RRTFND_STREAM_WORDS && !bt->words) return RRTFND_E_BADARG;
    if ((prm->flags & RRTFND_FLAG_TRACE) && (!bt->trace || prm->trace_cap < 1)) return RRTFND_E_BADARG;
    return 0;
}

extern "C" int64_t rrtfnd_plan_smem_bytes(const rrtfnd_params_t* prm) {
    if (!prm) return RRTFND_E_BADARG;
    const size_t b = cfg_smem(*prm, choose_cfg(*prm, 1 << 20));
    return b <= kMaxSmem ? (int64_t)b : (int64_t)RRTFND_E_SMEM;
}

extern "C" int rrtfnd_plan_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, void* stream) {
    int rc = check_params(prm, bt);
    if (rc) return rc;
    rrtfnd_params_t q = *prm;
    if (!(q.flags & RRTFND_FLAG_TRACE)) q.trace_cap = 0;
    rrtfnd_batch_t b = *bt;
    if (!(q.flags & RRTFND_FLAG_TRACE)) b.trace = nullptr;
    return launch_cfg(&q, &b, choose_cfg(q, b.n_planners), (cudaStream_t)stream);
}

extern "C" int rrtfnd_init_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, void* stream) {
    if (!prm || !bt || bt->n_planners <= 0) return RRTFND_E_BADARG;
    const int nt = 128;
    init_kernel<<<(bt->n_planners + nt - 1) / nt, nt, 0, (cudaStream_t)stream>>>(*prm, *bt);
    return (int)cudaGetLastError();
}
