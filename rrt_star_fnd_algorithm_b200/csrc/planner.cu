// planner.cu — fused persistent planner kernel: one WARP grows one tree from sample to best-path bookkeeping.
//
// Replaces the Python loops of RRT / RRT_star / RRT_star_FN (src/algorithm.py:54-265) together with
// new_node (:36-51), nearest_node_kdtree (:666-699), steer (:731-746), choose_parent_kdtree (:822-852),
// rewire_kdtree (:882-910), forced_removal (:800-819), check_solution (:749-760), find_path (:778-797)
// and the Graph bookkeeping of src/graph.py:34-87.
//
// Execution model. Planners are independent and strictly sequential inside (iteration i+1 needs the tree of
// iteration i), and after obstacle culling one iteration is a few thousand fp64 operations plus several
// latency-bound pointer chases. So a planner gets one warp, not a CTA: no block barriers, reductions are
// warp votes / redux, warp-uniform values live in registers of every lane, and the SM hides one planner's
// latency (parent-chain walks, div / sqrt chains) behind the other 16-64 planners resident on it. A batch larger
// than the resident-warp count streams through the hardware block scheduler.
//
// Layout (HBM, per planner, fixed capacity, slots in ascending-id order): xy[] (double2, the array every
// nearest / near-set scan streams with one coalesced 16-byte load per lane), ecost[], parent[], nchild[], id[],
// nflags[]. Inserts append; FN removals leave a tombstone (x = y = NaN so every scan predicate is false,
// nchild = id = -1) and a stable in-place compaction runs only when the slot array is full. The obstacle table
// and its culling grid (grid.cuh) are shared by all planners and stay L1 resident.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "cull.cuh"
#include "grid.cuh"
#include "rng.cuh"
#include "stream.cuh"

namespace rrtfnd {

constexpr int kReserveWords = 64;   // must equal RESERVE_WORDS in _structs.py / ORC_RESERVE_WORDS
constexpr int kStage = 160;         // slots staged per warp: < 32 left over + one 128-node block of the scan
constexpr int kIntMax = 0x7fffffff;
constexpr int kChainCap = 96;       // edge costs of the new node's parent chain cached per warp (choose-parent re-sums them)

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int bitmap_words(const rrtfnd_params_t& prm) { return prm.capacity / 32 + 1; }
// Compaction keeps a live bitmap + prefix counts in shared memory: always for FN; for the other modes (tombstones only
// after FND surgery) while the capacity is moderate. Without it a full slot array is a capacity error.
__host__ __device__ inline bool can_compact(const rrtfnd_params_t& prm) { return prm.mode == RRTFND_MODE_FN || prm.capacity <= 32768; }

// ------------------------------------------------------------------------------------ warp helpers

// (value, index) argmin over the warp for NON-NEGATIVE doubles (their bit patterns order like unsigned integers),
// lowest index on equal values. Three redux instructions instead of a five-step shuffle tree.
__device__ __forceinline__ void warp_argmin(double& bd, int& bs) {
    const unsigned hi = (unsigned)__double2hiint(bd), lo = (unsigned)__double2loint(bd);
    const unsigned mh = __reduce_min_sync(kFull, hi);
    const unsigned ml = __reduce_min_sync(kFull, hi == mh ? lo : 0xffffffffu);
    bs = __reduce_min_sync(kFull, (hi == mh && lo == ml) ? bs : kIntMax);
    bd = __hiloint2double((int)mh, (int)ml);
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// Leaf -> root sequential sum of edge costs (Graph.get_cost, src/graph.py:41-46). `acc` is the running
// value to start from (0 for get_cost; the new node's own edge cost when walking from its parent).
__device__ __forceinline__ double chain_walk(const int32_t* parent, const double* ecost, int slot, double acc,
                                             long long& hops) {
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

// find_path's cost (src/algorithm.py:787-792): like get_cost but the walk ends at the planner's root slot
__device__ __forceinline__ double path_cost(const int32_t* parent, const double* ecost, int from, int root, long long& hops) {
    double cost = 0.0;
    int n = from;
    while (n != root && n >= 0) { cost = dadd(cost, ecost[n]); n = parent[n]; ++hops; }
    return cost;
}

// ---- rare paths, kept out of line so that the per-iteration code stays small ---------------------------------

// Stable in-place compaction of a planner's slot arrays (FN only, when the slot array is full of tombstones).
static __device__ __noinline__ void compact_tree(double2* XY, double* ECOST, int32_t* PARENT, int32_t* NCHILD, int32_t* ID,
                                                 uint8_t* NFLAG, int32_t* FINISH, int n_finish, int n_slots, uint32_t* bm,
                                                 uint32_t* pre, int* root_slot, int* best_leaf_slot, int lane) {
    const int nwords = (n_slots + 31) >> 5;
    for (int w = 0; w < nwords; ++w) {
        const int s = w * 32 + lane;
        const unsigned b = __ballot_sync(kFull, s < n_slots && ID[s] >= 0);
        if (lane == 0) bm[w] = b;
    }
    __syncwarp();
    if (lane == 0) { int run = 0; for (int w = 0; w < nwords; ++w) { pre[w] = run; run += __popc(bm[w]); } }
    __syncwarp();
    auto rank = [&](int s) { return (int)pre[s >> 5] + __popc(bm[s >> 5] & ((1u << (s & 31)) - 1u)); };
    for (int base = 0; base < n_slots; base += 32) {
        const int s = base + lane;
        const bool alive = s < n_slots && ((bm[s >> 5] >> (s & 31)) & 1u);
        double2 v = make_double2(0.0, 0.0); double ve = 0; int vp = -1, vn = 0, vi = -1; uint8_t vf = 0;
        if (alive) {
            v = XY[s]; ve = ECOST[s]; vp = PARENT[s]; vn = NCHILD[s]; vi = ID[s]; vf = NFLAG[s];
            if (vp >= 0) vp = rank(vp);
        }
        __syncwarp();   // the whole row is read before any of it is overwritten (destinations are <= sources)
        if (alive) {
            const int d = rank(s);
            XY[d] = v; ECOST[d] = ve; PARENT[d] = vp; NCHILD[d] = vn; ID[d] = vi; NFLAG[d] = vf;
        }
        __syncwarp();
    }
    for (int f = lane; f < n_finish; f += 32) FINISH[f] = rank(FINISH[f]);
    *root_slot = rank(*root_slot);
    if (*best_leaf_slot >= 0) *best_leaf_slot = rank(*best_leaf_slot);
    __syncwarp();
}

// Is any vertex bit-identical to q? (only reached when a squared distance underflowed to 0 without equality)
static __device__ __noinline__ bool any_exact_match(const double2* XY, int n_slots, double qx, double qy, int lane) {
    bool exact = false;
    for (int s = lane; s < n_slots; s += 32) { const double2 t = XY[s]; exact |= (t.x == qx && t.y == qy); }
    return __any_sync(kFull, exact);
}

// Rank of key (vd, vs) among the nodes in (d2, slot) order = the number of k-NN queries the reference makes (:683-697)
static __device__ __noinline__ int key_rank(const double2* XY, int n_slots, double qx, double qy, double vd, int vs, int lane) {
    int rank = 0;
    for (int base = 0; base < n_slots; base += 32) {
        const int s = base + lane;
        bool le = false;
        if (s < n_slots) { const double2 t = XY[s]; const double d = dist2(t.x, t.y, qx, qy); le = d < vd || (d == vd && s <= vs); }
        rank += __popc(__ballot_sync(kFull, le));
    }
    return rank;
}

// ---- node grid (RRTFND_FLAG_NODE_GRID): candidate lists instead of full scans for large trees -------------------

// Gather the slots of every node in cells [cx0, cx1] x [cy0, cy1] into cand[0 .. cap). Lanes take cells round-robin and
// walk their linked lists in lock step. Returns the number of nodes found (> cap = overflow, list unusable).
static __device__ __noinline__ int ng_gather(const int32_t* __restrict__ head, const int32_t* __restrict__ next, int nx,
                                             int cx0, int cx1, int cy0, int cy1, int32_t* cand, int cap, int lane) {
    const int w = cx1 - cx0 + 1, n_cells = w * (cy1 - cy0 + 1);
    const unsigned lt = (1u << lane) - 1u;
    int n = 0;
    for (int k0 = 0; k0 < n_cells && n <= cap; k0 += 32) {
        const int k = k0 + lane;
        int cur = k < n_cells ? head[(cy0 + k / w) * nx + cx0 + k % w] : -1;
        for (;;) {
            const unsigned m = __ballot_sync(kFull, cur >= 0);
            if (!m) break;
            if (cur >= 0) {
                const int pos = n + __popc(m & lt);
                if (pos < cap) cand[pos] = cur;
                cur = next[cur];
            }
            n += __popc(m);
        }
    }
    __syncwarp();
    return n;
}

// ascending sort of n (<= 512) distinct slots by rank counting: out[rank(x)] = x
static __device__ __noinline__ void ng_sort(const int32_t* in, int32_t* out, int n, int lane) {
    for (int i = lane; i < n; i += 32) {
        const int x = in[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += in[j] < x;
        out[r] = x;
    }
    __syncwarp();
}

// key_rank over a candidate list that holds every node with a key <= (vd, vs)
static __device__ __noinline__ int ng_key_rank(const double2* XY, const int32_t* list, int n, double qx, double qy, double vd,
                                               int vs, int lane) {
    int rank = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool le = false;
        if (i < n) { const int s = list[i]; const double2 t = XY[s]; const double d = dist2(t.x, t.y, qx, qy); le = d < vd || (d == vd && s <= vs); }
        rank += __popc(__ballot_sync(kFull, le));
    }
    return rank;
}

// One tile of 128 (slot, position) pairs for a scan body: from the TMA ring when scanning the whole tree, gathered
// through a candidate list otherwise. Padding has slot -1 / position NaN, for which every predicate is false.
template <bool NG>
__device__ __forceinline__ void fetch_tile(XYStream& xs, const double2* XY, const int32_t* list, int n_items, int t, int lane,
                                           double2 (&v)[4], int (&sl)[4]) {
    const int base = t * kTileNodes;
    const double2 nan2 = make_double2(CUDART_NAN, CUDART_NAN);
    if (NG && list) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = base + 32 * u + lane; sl[u] = i < n_items ? list[i] : -1; }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = sl[u] >= 0 ? XY[sl[u]] : nan2;
    } else if (xs.n_stages == 0) {      // plain coalesced 16-byte loads, four in flight per lane
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int s = base + 32 * u + lane; sl[u] = s; v[u] = s < n_items ? XY[s] : nan2; }
    } else {
        const double2* tile = stream_wait(xs, t);
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int s = base + 32 * u + lane; sl[u] = s; v[u] = s < n_items ? tile[32 * u + lane] : nan2; }
        stream_release(xs, XY, n_items, t, lane);
    }
}

// ---- the planner of one warp ---------------------------------------------------------------------------------------------
//
// The kernel body is a short loop over PHASES, each a non-inlined function. Everything a planner carries from one phase
// to the next - its counters, the current sample, the new node - lives in a per-warp shared-memory record (WarpCtx) that
// every lane reads by broadcast and writes with the same value, so no phase inherits live registers from another: the
// register allocation is the maximum over the phases, not their union, and 32 warps fit an SM without spills.

struct WarpCtx {
    XYStream xs;
    // planner state (rrtfnd_planner_t fields, loaded at the start and written back at the end)
    double gx, gy, best_cost;
    long long attempts, n_edge_checks, n_edge_checks_ref, n_scan_nodes, n_hops, n_tests;
    int n_slots, n_live, last_id, iter, root_slot, n_finish, best_len, status;
    int n_rewired, n_removed, solution_found, n_compactions, fn_count, best_leaf_slot, n_childless;
    // the current attempt / iteration
    double qx, qy, px, py, e0, cp_cost;
    int near_slot, new_slot, id_new, reached, k_total, n_rw, cp_parent_id, removed_id, need_refresh, pad;
};

// shared memory of one warp: xy ring (n_stages tiles) | mbarriers | WarpCtx | word source | chain[kChainCap] | stage[kStage] |
// rwlist[near_cap] (re-parented slots of one iteration; overflow goes to nflags) | live bitmap + prefix counts (compaction)
__host__ __device__ inline size_t warp_smem_bytes(const rrtfnd_params_t& prm, int n_stages) {
    size_t o = (size_t)n_stages * kTileBytes + align_up((size_t)n_stages * 8, 16) + align_up(sizeof(WarpCtx), 16) +
               align_up(sizeof(WordSrc), 16) + kChainCap * 8 + kStage * 4 + align_up((size_t)prm.near_cap * 4, 16);
    if (can_compact(prm)) o += 2 * align_up((size_t)bitmap_words(prm) * 4, 16);
    return align_up(o, 128);
}

// Everything a phase needs, re-derived from the kernel parameters and the warp's shared-memory base (cheap integer math).
struct Warp {
    const rrtfnd_params_t* prm;
    const rrtfnd_batch_t* bt;
    WarpCtx* c;
    WordSrc* rng;
    double* chain;
    int* stage; int* rwlist; uint32_t* bm; uint32_t* pre;
    double2* XY; double* ECOST; int32_t* PARENT; int32_t* NCHILD; int32_t* ID; uint8_t* NFLAG; int32_t* FINISH; int32_t* BPATH;
    int32_t* HEAD; int32_t* NEXT; int32_t* CAND;
    int lane, p;
};

__device__ __forceinline__ Warp make_warp(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    Warp w;
    w.prm = prm; w.bt = bt;
    w.lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    w.p = blockIdx.x * (blockDim.x >> 5) + wib;
    const int n_stages = prm->reserved;                  // ring depth, resolved by the host wrapper (0 = plain loads)
    unsigned char* wp = smem_base + (size_t)wib * warp_smem_bytes(*prm, n_stages) + (size_t)n_stages * kTileBytes +
                        align_up((size_t)n_stages * 8, 16);
    w.c = reinterpret_cast<WarpCtx*>(wp); wp += align_up(sizeof(WarpCtx), 16);
    w.rng = reinterpret_cast<WordSrc*>(wp); wp += align_up(sizeof(WordSrc), 16);
    w.chain = reinterpret_cast<double*>(wp); wp += kChainCap * 8;
    w.stage = reinterpret_cast<int*>(wp);
    w.rwlist = w.stage + kStage;
    w.bm = reinterpret_cast<uint32_t*>(w.rwlist + align_up((size_t)prm->near_cap * 4, 16) / 4);
    w.pre = w.bm + align_up((size_t)bitmap_words(*prm) * 4, 16) / 4;
    const size_t o = (size_t)w.p * prm->capacity;
    w.XY = reinterpret_cast<double2*>(bt->xy) + o; w.ECOST = bt->ecost + o; w.PARENT = bt->parent + o; w.NCHILD = bt->nchild + o;
    w.ID = bt->id + o; w.NFLAG = bt->nflags + o;
    w.FINISH = bt->finish + (size_t)w.p * prm->finish_cap; w.BPATH = bt->best_path + (size_t)w.p * prm->path_cap;
    const rrtfnd_node_grid_t& ng = bt->node_grid;
    const bool has_ng = (prm->flags & RRTFND_FLAG_NODE_GRID) != 0;
    w.HEAD = has_ng ? ng.head + (size_t)w.p * ng.nx * ng.ny : nullptr;
    w.NEXT = has_ng ? ng.next + o : nullptr;
    w.CAND = has_ng ? ng.cand + (size_t)w.p * 2 * ng.cand_cap : nullptr;
    return w;
}

// per-phase work counters -> the warp's totals
__device__ __forceinline__ void add_counters(const Warp& w, int hops, int tests) {
    hops = __reduce_add_sync(kFull, hops); tests = __reduce_add_sync(kFull, tests);
    __syncwarp();
    if (w.lane == 0) { w.c->n_hops += hops; w.c->n_tests += tests; }
    __syncwarp();
}

// all nodes within rho cells' width of a point lie in the (2 rho + 1)^2 cells around the point's cell
__device__ __forceinline__ bool ng_window(const rrtfnd_node_grid_t& ng, double x, double y, int rho, int& cx0, int& cx1, int& cy0, int& cy1) {
    const int cx = grid_cell(x, ng.x0, ng.inv_cell, ng.nx), cy = grid_cell(y, ng.y0, ng.inv_cell, ng.ny);
    cx0 = max(cx - rho, 0); cx1 = min(cx + rho, ng.nx - 1); cy0 = max(cy - rho, 0); cy1 = min(cy + rho, ng.ny - 1);
    return cx0 == 0 && cy0 == 0 && cx1 == ng.nx - 1 && cy1 == ng.ny - 1;
}
__device__ __forceinline__ double ng_complete_d2(const rrtfnd_node_grid_t& ng, int rho) { const double w = rho * ng.cell; return w * w * (1.0 - 1e-9); }

enum { kNext = 0, kRetry = 1, kStop = 2 };

// ================================================================ phase 1: one attempt up to the nearest visible node
// A. header checks  B. sample (Graph.random_node) + Map.is_occupied_c  C. nearest_node_kdtree (:666-699)
// The reference asks the k-d tree for the 1st, 2nd, 3rd ... nearest node until one is visible. Round 0 here is that first
// query (plain argmin). If it is blocked, the remaining order is consumed in annuli of doubling squared radius: every node
// of an annulus is tested (one lane per segment) and the lowest (d2, slot) key among the visible ones is the node the
// reference would have stopped at, because everything in earlier annuli was blocked. A boxed-in sample therefore costs
// O(log N) scans instead of O(N).
template <bool NG>
static __device__ __noinline__ int phase_attempt(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    WordSrc& rng = *w.rng;
    const int lane = w.lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    double2* XY = w.XY;
    int* stage = w.stage;
    const rrtfnd_node_grid_t& ng = bt->node_grid;

    if (c.iter >= prm->iter_num) { c.status = RRTFND_ST_DONE; return kStop; }
    if (prm->max_attempts > 0 && c.attempts >= prm->max_attempts) { c.status = RRTFND_ST_ATTEMPT_CAP; return kStop; }
    if (prm->stream_mode == RRTFND_STREAM_WORDS && (long long)rng.cursor + kReserveWords > prm->words_per_planner) {
        c.status = RRTFND_ST_NEED_WORDS; return kStop;
    }
    if (c.n_slots >= prm->capacity) {
        if (c.n_live >= c.n_slots || !can_compact(*prm)) { c.status = RRTFND_ST_CAPACITY; return kStop; }
        // stable in-place compaction (FN, or tombstones left by FND surgery; rare)
        int root = c.root_slot, leaf = c.best_leaf_slot;
        compact_tree(XY, w.ECOST, w.PARENT, w.NCHILD, w.ID, w.NFLAG, w.FINISH, c.n_finish, c.n_slots, w.bm, w.pre, &root, &leaf, lane);
        c.root_slot = root; c.best_leaf_slot = leaf;
        c.n_slots = c.n_live;
        c.n_compactions += 1;
    }
    c.attempts += 1;
    const int n_slots = c.n_slots, n_live = c.n_live;
    // Graph.random_node (src/graph.py:96-98)
    double qx, qy;
    if (next_random(rng) < prm->bias) { qx = c.gx; qy = c.gy; }
    else {
        qx = dadd(prm->x_lo, dmul(dsub(prm->x_hi, prm->x_lo), next_random(rng)));
        qy = dadd(prm->y_lo, dmul(dsub(prm->y_hi, prm->y_lo), next_random(rng)));
    }
    const Obst ob = make_obst(bt->obst, prm->n_obst, &bt->grid);
    int tests = 0;
    {   // Map.is_occupied_c (src/map.py:38-47); rejected at src/algorithm.py:39-40
        long long t = 0;
        const bool occ = point_occupied_warp(ob, qx, qy, prm->node_radius, lane, t);
        tests += (int)t;
        if (occ) { add_counters(w, 0, tests); return kRetry; }
    }

    int near_slot = -1, rounds = 1;
    long long scanned = 0;
    double bd = CUDART_INF; int bs = kIntMax;
    // what the scans below run over: the whole tree (TMA ring), or - node grid - the nodes of a window of cells
    const int32_t* list = nullptr;
    int n_items = n_slots, rho = 1;
    bool covers_all = true;
    const bool use_ng = NG && n_live >= ng.min_nodes;
    for (;;) {      // (node grid: widen the window until the nearest node is certain; one pass when scanning)
        if (use_ng) {
            int cx0, cx1, cy0, cy1;
            covers_all = ng_window(ng, qx, qy, rho, cx0, cx1, cy0, cy1);
            const int n = ng_gather(w.HEAD, w.NEXT, ng.nx, cx0, cx1, cy0, cy1, w.CAND, ng.cand_cap, lane);
            if (n > ng.cand_cap) { list = nullptr; n_items = n_slots; covers_all = true; }
            else { list = w.CAND; n_items = n; }
        }
        bd = CUDART_INF; bs = kIntMax;
        const int n_tiles = (n_items + kTileNodes - 1) / kTileNodes;
        if (!list) stream_begin(c.xs, XY, n_slots, lane);
        for (int t = 0; t < n_tiles; ++t) {
            double2 v[4]; int sl[4];
            fetch_tile<NG>(c.xs, XY, list, n_items, t, lane, v, sl);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double d = dist2(v[u].x, v[u].y, qx, qy);                 // NaN for tombstones / padding
                if (d < bd || (NG && d == bd && sl[u] < bs)) { bd = d; bs = sl[u]; }
            }
        }
        warp_argmin(bd, bs);
        scanned += list ? n_items : n_live;
        if (!list || covers_all || (bs != kIntMax && bd <= ng_complete_d2(ng, rho))) break;
        rho *= 2;
    }
    bool reject = bs == kIntMax;                                              // empty tree (cannot happen)
    if (!reject && bd == 0.0) {                                               // G.id_vertex[vertex] hit (:676-678), rejected at :45
        const double2 e = XY[bs];
        reject = (e.x == qx && e.y == qy);
        if (!reject) reject = any_exact_match(XY, n_slots, qx, qy, lane);     // d2 underflowed to 0 without equality
    }
    if (reject) { c.n_scan_nodes += scanned; add_counters(w, 0, tests); return kRetry; }
    {   // through_obstacle(Line(q_rand, node)) — p1 = sample, p2 = node (:690-692); the warp shares the one segment
        const double2 nb = XY[bs];
        const int r = seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, qx, qy, nb.x, nb.y, lane, 32);
        tests += r >> 1;
        if (!__any_sync(kFull, r & 1)) near_slot = bs;
    }
    if (near_slot < 0 && n_live > 1) {
        double lo_d = bd; int lo_s = bs;                                      // keys <= (lo_d, lo_s) are known to be blocked
        double T = bd;
        const double t_floor = dmul(dmul(prm->step_length, prm->step_length), 0.0625);
        rounds = n_live;                                                      // if nothing is visible every node was tried
        for (;;) {
            T = dadd(T, T);
            if (T < t_floor) T = t_floor;
            if (NG && list && !covers_all && T > ng_complete_d2(ng, rho)) {   // the window must hold every node within sqrt(T)
                rho = max(2 * rho, (int)(sqrt(T) * ng.inv_cell) + 1);
                int cx0, cx1, cy0, cy1;
                covers_all = ng_window(ng, qx, qy, rho, cx0, cx1, cy0, cy1);
                const int n = ng_gather(w.HEAD, w.NEXT, ng.nx, cx0, cx1, cy0, cy1, w.CAND, ng.cand_cap, lane);
                if (n > ng.cand_cap) { list = nullptr; n_items = n_slots; covers_all = true; }
                else n_items = n;
            }
            double vd = CUDART_INF; int vs = kIntMax;                          // this lane's best visible key
            bool beyond = NG && list && !covers_all;
            int count = 0;
            const int n_tiles = (n_items + kTileNodes - 1) / kTileNodes;
            if (!list) stream_begin(c.xs, XY, n_slots, lane);
            for (int t = 0; t < n_tiles; ++t) {
                double2 v[4]; int sl[4];
                fetch_tile<NG>(c.xs, XY, list, n_items, t, lane, v, sl);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int s = sl[u];
                    const double d = dist2(v[u].x, v[u].y, qx, qy);
                    beyond |= d > T;
                    const bool in = (d > lo_d || (d == lo_d && s > lo_s)) && d <= T;
                    const unsigned m = __ballot_sync(kFull, in);
                    if (in) stage[count + __popc(m & lt_mask)] = s;
                    count += __popc(m);
                }
                const bool last = t + 1 >= n_tiles;
                int head = 0;
                while (count - head >= 32 || (last && count > head)) {         // one lane per staged segment
                    const int n = min(count - head, 32);
                    __syncwarp();
                    if (lane < n) {
                        const int s = stage[head + lane];
                        const double2 nb = XY[s];
                        const int r = seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, qx, qy, nb.x, nb.y, 0, 1);
                        tests += r >> 1;
                        if (!(r & 1)) {
                            const double d = dist2(nb.x, nb.y, qx, qy);
                            if (d < vd || (d == vd && s < vs)) { vd = d; vs = s; }
                        }
                    }
                    __syncwarp();
                    head += n;
                }
                if (head > 0) {                                               // keep the (< 32) leftover at the front
                    const int rem = count - head;
                    const int t2 = lane < rem ? stage[head + lane] : 0;
                    __syncwarp();
                    if (lane < rem) stage[lane] = t2;
                    count = rem;
                    __syncwarp();
                }
            }
            scanned += list ? n_items : n_live;
            warp_argmin(vd, vs);
            if (vs != kIntMax) {
                near_slot = vs;
                // the reference's query count = rank of the winner in (d2, slot) order
                rounds = (NG && list) ? ng_key_rank(XY, list, n_items, qx, qy, vd, vs, lane) : key_rank(XY, n_slots, qx, qy, vd, vs, lane);
                break;
            }
            if (!__any_sync(kFull, beyond)) break;                            // every node tried, none visible
            lo_d = T; lo_s = kIntMax;
        }
    }
    c.n_scan_nodes += scanned;
    c.n_edge_checks += rounds; c.n_edge_checks_ref += rounds;
    add_counters(w, 0, tests);
    if (near_slot < 0) return kRetry;                                         // :45-46
    c.qx = qx; c.qy = qy; c.near_slot = near_slot;
    return kNext;
}

// ================================================================ phase 2: steer + insert (every lane computes, lane 0 stores)
template <bool NG>
static __device__ __noinline__ int phase_insert(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    const int lane = w.lane, near_slot = c.near_slot;
    const double qx = c.qx, qy = c.qy;
    const double2 nb = w.XY[near_slot];
    double px = qx, py = qy;
    {
        const double d = dist_fma(qx, qy, nb.x, nb.y);                        // steer (:739)
        const double ux = ddiv(dsub(qx, nb.x), d), uy = ddiv(dsub(qy, nb.y), d);
        const double tx = dadd(nb.x, dmul(ux, prm->step_length)), ty = dadd(nb.y, dmul(uy, prm->step_length));
        if (d > prm->step_length) { px = tx; py = ty; }
    }
    const int new_slot = c.n_slots;
    const int id_new = c.last_id + 1;                                         // Graph.add_vertex (src/graph.py:57-61)
    c.n_slots = new_slot + 1; c.last_id = id_new;
    c.n_live += 1; c.fn_count += 1;                                           // n_of_nodes += 1 (:221)
    const double e0 = dist_fma(px, py, nb.x, nb.y);                           // :49
    const bool reached = dist_fma(px, py, c.gx, c.gy) < dmul(2.0, prm->goal_node_radius);   // check_solution (:757-758)
    c.px = px; c.py = py; c.e0 = e0; c.new_slot = new_slot; c.id_new = id_new; c.reached = reached;
    if (lane == 0) {
        w.XY[new_slot] = make_double2(px, py);
        w.ID[new_slot] = id_new; w.PARENT[new_slot] = near_slot; w.ECOST[new_slot] = e0;   // :141-142
        w.NCHILD[new_slot] = 0; w.NFLAG[new_slot] = 0;
        if (NG) {                                                             // push onto its cell's list
            const rrtfnd_node_grid_t& ng = bt->node_grid;
            const int cell = grid_cell(py, ng.y0, ng.inv_cell, ng.ny) * ng.nx + grid_cell(px, ng.x0, ng.inv_cell, ng.nx);
            w.NEXT[new_slot] = w.HEAD[cell]; w.HEAD[cell] = new_slot;
        }
    }
    __syncwarp();
    if (prm->mode != RRTFND_MODE_RRT) return kNext;

    // ---- RRT (:76-84): add_edge, goal test, no near set
    int L = 0; double cost = 0.0; int hops = 0;
    rrtfnd_trace_t* TRACE = bt->trace ? bt->trace + (size_t)w.p * prm->trace_cap : nullptr;
    if (lane == 0) {
        w.NCHILD[near_slot] += 1;                                             // add_edge (:76)
        if (reached) {                                                        // :78-81, break before iter += 1
            int n = new_slot;
            while (n != c.root_slot && n >= 0) {
                if (L < prm->path_cap) w.BPATH[L] = w.ID[n];
                ++L; cost = dadd(cost, w.ECOST[n]); n = w.PARENT[n]; ++hops;
            }
            if (L < prm->path_cap) w.BPATH[L] = w.ID[c.root_slot];
            ++L;
        }
        if (TRACE && c.iter < prm->trace_cap) {
            rrtfnd_trace_t t; t.id_new = id_new; t.id_near = w.ID[near_slot]; t.n_near = 0; t.n_rewired = 0;
            t.removed_id = -1; t.cp_parent = -1; t.cp_cost = 0.0;
            TRACE[c.iter] = t;
        }
    }
    __syncwarp();
    add_counters(w, hops, 0);
    if (reached) {
        const int len = __shfl_sync(kFull, L, 0);
        c.best_len = len; c.best_cost = __shfl_sync(kFull, cost, 0);
        c.best_leaf_slot = new_slot; c.solution_found = 1;
        c.status = len > prm->path_cap ? RRTFND_ST_PATH_CAP : RRTFND_ST_DONE;
        return kStop;
    }
    c.iter += 1;
    return kRetry;                                                            // next attempt
}

// ================================================================ phase 3: near set -> collision, costs, choose-parent, rewire
// The ball (query_ball_point, :837 / :893) is produced in ascending slot order by the scan and handed to the lanes in chunks
// of up to 32 candidates: lane i owns candidate i (its own segment test, its own parent-chain walk). In the first chunk
// lane 31 walks the new node's chain (Graph.get_cost(id_new)). Rewire decisions are taken against the pre-rewire tree and
// committed after the scan.
template <bool NG>
static __device__ __noinline__ int phase_near(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    const int lane = w.lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    double2* XY = w.XY; double* ECOST = w.ECOST; int32_t* PARENT = w.PARENT; int32_t* NCHILD = w.NCHILD; uint8_t* NFLAG = w.NFLAG;
    int* stage = w.stage; int* rwlist = w.rwlist; double* chain = w.chain;
    const rrtfnd_node_grid_t& ng = bt->node_grid;
    const Obst ob = make_obst(bt->obst, prm->n_obst, &bt->grid);
    const double px = c.px, py = c.py, e0 = c.e0;
    const int near_slot = c.near_slot, new_slot = c.new_slot, n_slots = c.n_slots, n_live = c.n_live;
    const double r2 = dmul(prm->radius, prm->radius);
    const bool eval_cp = (prm->flags & (RRTFND_FLAG_EVAL_CHOOSE_PARENT | RRTFND_FLAG_COMMIT_CHOOSE_PARENT)) != 0;
    const bool commit_cp = (prm->flags & RRTFND_FLAG_COMMIT_CHOOSE_PARENT) != 0;

    int n_rw = 0, k_total = 0, hops = 0, tests = 0;
    int cp_slot = near_slot; double cp_cost = e0;
    bool dup = false;
    double cost_new = 0.0;
    // node grid: the ball lies inside the window of cells within `radius` of q_new; its nodes, sorted by slot, replace the scan
    const int32_t* nlist = nullptr;
    int n_near_items = n_slots;
    if (NG && n_live >= ng.min_nodes) {
        int cx0, cx1, cy0, cy1;
        ng_window(ng, px, py, (int)(prm->radius * ng.inv_cell) + 1, cx0, cx1, cy0, cy1);
        const int n = ng_gather(w.HEAD, w.NEXT, ng.nx, cx0, cx1, cy0, cy1, w.CAND, ng.cand_cap, lane);
        if (n <= ng.cand_cap && n <= 512) { ng_sort(w.CAND, w.CAND + ng.cand_cap, n, lane); nlist = w.CAND + ng.cand_cap; n_near_items = n; }
    }
    const int n_pass = commit_cp ? 2 : 1;          // committing choose-parent needs its result before any rewire decision
#pragma unroll 1
    for (int pass = 0; pass < n_pass; ++pass) {
        const bool do_cp = eval_cp && pass == 0;
        const bool do_rw = pass == n_pass - 1;
        const int walk_from = PARENT[new_slot];
        const double walk_acc = ECOST[new_slot];
        double cp_cn = 0.0;
        int chain_depth = 0;
        bool first = true;
        int count = 0;
        k_total = 0;
        const int n_tiles = (n_near_items + kTileNodes - 1) / kTileNodes;
        if (!nlist) stream_begin(c.xs, XY, n_slots, lane);
        for (int t = 0; t < n_tiles; ++t) {
            double2 v[4]; int sl[4];
            fetch_tile<NG>(c.xs, XY, nlist, n_near_items, t, lane, v, sl);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = sl[u];
                const bool in = dist2(v[u].x, v[u].y, px, py) <= r2 && s != new_slot;
                const unsigned m = __ballot_sync(kFull, in);
                if (in) stage[count + __popc(m & lt_mask)] = s;
                count += __popc(m); k_total += __popc(m);
            }
            const bool last = t + 1 >= n_tiles;
            int head = 0;
            while (count - head >= (first ? 31 : 32) || (last && count > head)) {
                const int n = min(count - head, first ? 31 : 32);
                __syncwarp();
                // ---- one chunk: lane i < n owns candidate i; in the first chunk lane 31 walks the new node's chain
                int s = -1; double dpl = 0.0, cc = 0.0; bool hit = true;
                int w_from = -1;
                if (lane < n) {
                    s = stage[head + lane];
                    const double2 cpos = XY[s];
                    dpl = dist_plain(cpos.x, cpos.y, px, py);                 // :842 / :899
                    if (dpl == 0.0 && cpos.x == px && cpos.y == py) dup = true;   // src/graph.py:54-56 corner (B9)
                    // Line(node, q_new): p1 = node (:846 / :903)
                    const int r = seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, cpos.x, cpos.y, px, py, 0, 1);
                    tests += r >> 1;
                    hit = r & 1;
                    if (!hit) w_from = s;                                     // Graph.get_cost(id) (:848 / :905)
                }
                // parent-chain walks, one loop for all lanes (Graph.get_cost, src/graph.py:41-46): lane 31 of the first
                // chunk walks the new node's chain, starts from its edge cost and records the addends
                const bool rec = first && lane == 31;
                int depth = 0;
                if (rec) { w_from = walk_from; cc = walk_acc; }
                if (w_from >= 0) {
                    int wn = w_from, q = PARENT[wn];
                    while (q >= 0) {
                        const double e = ECOST[wn];
                        if (rec && depth < kChainCap) chain[depth] = e;
                        cc = dadd(cc, e);
                        wn = q; q = PARENT[wn];
                        ++depth;
                    }
                    hops += depth;
                }
                __syncwarp();
                if (first) { cost_new = __shfl_sync(kFull, cc, 31); cp_cn = cost_new; chain_depth = __shfl_sync(kFull, depth, 31); }
                if (do_cp) {
                    // choose_parent_kdtree's RETURNED edge (:845-852), candidates in ascending id
                    unsigned open = __ballot_sync(kFull, lane < n && !hit);
                    while (open) {
                        const int i = __ffs(open) - 1;
                        open &= open - 1;
                        const double ci = __shfl_sync(kFull, cc, i), di = __shfl_sync(kFull, dpl, i);
                        if (cp_cn > dadd(ci, di)) {
                            cp_slot = __shfl_sync(kFull, s, i); cp_cost = di;
                            // G.cost[id_new] = cost, parent untouched: the chain is still id_near's (:849)
                            if (chain_depth <= kChainCap) {                   // re-sum the cached addends
                                cp_cn = di;
                                for (int h = 0; h < chain_depth; ++h) cp_cn = dadd(cp_cn, chain[h]);
                            } else {
                                long long h = 0;
                                cp_cn = chain_walk(PARENT, ECOST, near_slot, di, h);   // every lane walks the same chain
                                if (lane == 0) hops += (int)h;
                            }
                        }
                    }
                }
                if (do_rw) {
                    const bool rw = lane < n && !hit && cc > dadd(cost_new, dpl);   // :905, strict
                    const unsigned m = __ballot_sync(kFull, rw);
                    if (m) {   // remember the decision: short list in shared memory, overflow as a per-node flag bit
                        const int pos = n_rw + __popc(m & lt_mask);
                        if (rw) { if (pos < prm->near_cap) rwlist[pos] = s; else NFLAG[s] |= 2; }
                        n_rw += __popc(m);
                    }
                }
                first = false;
                __syncwarp();
                head += n;
            }
            if (head > 0) {                                                   // keep the (< 32) leftover at the front
                const int rem = count - head;
                const int t2 = lane < rem ? stage[head + lane] : 0;
                __syncwarp();
                if (lane < rem) stage[lane] = t2;
                count = rem;
                __syncwarp();
            }
        }
        if (pass == 0 && commit_cp && k_total > 0) {   // EXTENSION (not the shipped behaviour): adopt the returned edge
            if (lane == 0) { PARENT[new_slot] = cp_slot; ECOST[new_slot] = cp_cost; }
            __syncwarp();
        }
    }
    add_counters(w, hops, tests);
    if (__any_sync(kFull, dup)) { c.status = RRTFND_ST_DUPLICATE; return kStop; }
    c.n_scan_nodes += nlist ? n_near_items : n_live;
    c.n_edge_checks += k_total;
    c.n_edge_checks_ref += 2 * (long long)k_total + 2;   // each candidate twice + Line(q_new, q_new) twice

    // ---- commit: rewire_kdtree's re-parenting (:906-910) and G.add_edge(*best_edge) (:148), serial on lane 0
    int d_childless = 0, flagged = 0;
    if (lane == 0) {
        auto reparent = [&](int s) {
            const int oldp = PARENT[s];
            if (--NCHILD[oldp] == 0) ++d_childless;
            const double2 v = XY[s];
            PARENT[s] = new_slot;
            ECOST[s] = dist_plain(v.x, v.y, px, py);
            flagged |= NFLAG[s] & 1;
        };
        const int n_list = n_rw < prm->near_cap ? n_rw : prm->near_cap;
        for (int i = 0; i < n_list; ++i) reparent(rwlist[i]);
        if (n_rw > n_list) {          // more re-parented nodes than the list holds (rare): find the flagged rest
            for (int s = 0; s < new_slot; ++s)
                if (NFLAG[s] & 2) { NFLAG[s] &= (uint8_t)~2; reparent(s); }
        }
        NCHILD[new_slot] = n_rw;
        if (n_rw == 0) ++d_childless;
        if (NCHILD[PARENT[new_slot]]++ == 0) --d_childless;
        if (flagged) {   // the moved subtree holds a goal-reaching node: mark its new ancestors
            int n = new_slot;
            while (n >= 0 && !(NFLAG[n] & 1)) { NFLAG[n] |= 1; n = PARENT[n]; }
        }
    }
    __syncwarp();
    c.n_childless += __shfl_sync(kFull, d_childless, 0);
    c.need_refresh = __shfl_sync(kFull, flagged, 0) != 0;
    c.n_rewired += n_rw;
    c.k_total = k_total; c.n_rw = n_rw; c.cp_cost = cp_cost;
    c.cp_parent_id = eval_cp ? w.ID[cp_slot] : -1;     // read before forced_removal can delete that node
    return kNext;
}

// ================================================================ phase 4: forced_removal (:233-237, :800-819)
static __device__ __noinline__ int phase_fn_remove(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    WordSrc& rng = *w.rng;
    const int lane = w.lane, new_slot = c.new_slot, n_slots = c.n_slots;
    int32_t* NCHILD = w.NCHILD; int32_t* FINISH = w.FINISH;
    const int leaf = c.best_leaf_slot;
    const int total = c.n_childless;
    int n_ok = total - (NCHILD[new_slot] == 0 ? 1 : 0);
    if (leaf >= 0 && leaf != new_slot && NCHILD[leaf] == 0) --n_ok;
    if (n_ok <= 0) { c.status = RRTFND_ST_NO_CHILDLESS; return kStop; }
    int pick = -1;
    for (;;) {
        const uint32_t r = next_randbelow(rng, (uint32_t)total);              // random.choice (:813 / :816)
        if (rng.exhausted) { c.status = RRTFND_ST_STREAM_EXHAUSTED; return kStop; }
        // r-th childless node in ascending slot order (tombstones hold nchild = -1)
        int running = 0; pick = -1;
        for (int base = 0; base < n_slots; base += 32) {
            const int s = base + lane;
            const unsigned m = __ballot_sync(kFull, s < n_slots && NCHILD[s] == 0);
            const int cnt = __popc(m);
            if ((int)r < running + cnt) { pick = base + (int)__fns(m, 0, (int)r - running + 1); break; }
            running += cnt;
        }
        if (pick != new_slot && pick != leaf) break;                          // :815
    }
    // finish_nodes_of_path.remove(id_removed) (:235-236), order preserving
    const int n_finish = c.n_finish;
    int fpos = -1;
    for (int base = 0; base < n_finish; base += 32) {
        const int f = base + lane;
        const unsigned m = __ballot_sync(kFull, f < n_finish && FINISH[f] == pick);
        if (m) { fpos = base + __ffs(m) - 1; break; }
    }
    if (fpos >= 0) {
        for (int base = fpos; base < n_finish - 1; base += 32) {
            const int f = base + lane;
            const bool ok = f < n_finish - 1;
            const int v = ok ? FINISH[f + 1] : 0;
            __syncwarp();
            if (ok) FINISH[f] = v;
            __syncwarp();
        }
        c.n_finish = n_finish - 1;
    }
    int d_childless = -1;                                                     // the removed node was childless
    c.removed_id = w.ID[pick];
    __syncwarp();
    if (lane == 0) {                                                          // Graph.remove_vertex (src/graph.py:64-77)
        const int par = w.PARENT[pick];
        if (par >= 0 && --NCHILD[par] == 0) ++d_childless;
        w.ID[pick] = -1; NCHILD[pick] = -1; w.PARENT[pick] = -1;
        w.XY[pick] = make_double2(CUDART_NAN, CUDART_NAN);
    }
    __syncwarp();
    c.n_childless += __shfl_sync(kFull, d_childless, 0);
    c.n_live -= 1; c.fn_count -= 1; c.n_removed += 1;
    return kNext;
}

// ================================================================ phase 5: goal test and best path (:155-170, :240-251), end of iteration
static __device__ __noinline__ int phase_goal(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    const int lane = w.lane, new_slot = c.new_slot, root_slot = c.root_slot;
    int32_t* PARENT = w.PARENT; double* ECOST = w.ECOST; int32_t* FINISH = w.FINISH; int32_t* BPATH = w.BPATH; int32_t* ID = w.ID;
    int hops = 0;
    const bool reached = c.reached != 0;
    if (reached) {
        if (c.n_finish >= prm->finish_cap) { c.status = RRTFND_ST_FINISH_CAP; return kStop; }
        int L = 0; double cost = 0.0;
        if (lane == 0) {
            FINISH[c.n_finish] = new_slot;
            int n = new_slot;                                                 // mark the ancestors of the new finish node
            while (n >= 0 && !(w.NFLAG[n] & 1)) { w.NFLAG[n] |= 1; n = PARENT[n]; }
            if (prm->mode == RRTFND_MODE_STAR) {                              // unconditional overwrite (:160-161)
                n = new_slot;
                while (n != root_slot && n >= 0) {
                    if (L < prm->path_cap) BPATH[L] = ID[n];
                    ++L; cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++hops;
                }
                if (L < prm->path_cap) BPATH[L] = ID[root_slot];
                ++L;
            }
        }
        __syncwarp();
        c.n_finish += 1;
        c.solution_found = 1;
        if (prm->mode == RRTFND_MODE_STAR) {
            const int len = __shfl_sync(kFull, L, 0);
            c.best_len = len; c.best_cost = __shfl_sync(kFull, cost, 0); c.best_leaf_slot = new_slot;
            if (len > prm->path_cap) { c.status = RRTFND_ST_PATH_CAP; add_counters(w, hops, 0); return kStop; }
        }
    }
    // The reference recomputes every finish node's path each iteration and keeps a strictly cheaper one (:166-170, :247-251).
    // A path cost only changes when a node with a goal-reaching descendant is rewired, and the list only grows through a
    // new finish node, so those are the only iterations in which it can have an effect.
    const int n_finish = c.n_finish;
    if ((reached || c.need_refresh) && n_finish > 0) {
        double bd = CUDART_INF; int bf = kIntMax;
        for (int f = lane; f < n_finish; f += 32) {
            long long h = 0;
            const double cost = path_cost(PARENT, ECOST, FINISH[f], root_slot, h);
            hops += (int)h;
            if (cost < bd) { bd = cost; bf = f; }
        }
        __syncwarp();
        warp_argmin(bd, bf);
        if (bd < c.best_cost) {                                               // strict (:168, :249)
            const int leaf = FINISH[bf];
            int L = 0;
            if (lane == 0) {
                int n = leaf;
                while (n != root_slot && n >= 0) {
                    if (L < prm->path_cap) BPATH[L] = ID[n];
                    ++L; n = PARENT[n]; ++hops;
                }
                if (L < prm->path_cap) BPATH[L] = ID[root_slot];
                ++L;
            }
            __syncwarp();
            const int len = __shfl_sync(kFull, L, 0);
            c.best_len = len; c.best_cost = bd; c.best_leaf_slot = leaf;
            if (len > prm->path_cap) { c.status = RRTFND_ST_PATH_CAP; add_counters(w, hops, 0); return kStop; }
        }
    }
    add_counters(w, hops, 0);
    // ---- end of iteration
    rrtfnd_trace_t* TRACE = bt->trace ? bt->trace + (size_t)w.p * prm->trace_cap : nullptr;
    if (TRACE && c.iter < prm->trace_cap && lane == 0) {
        const bool eval_cp = (prm->flags & (RRTFND_FLAG_EVAL_CHOOSE_PARENT | RRTFND_FLAG_COMMIT_CHOOSE_PARENT)) != 0;
        rrtfnd_trace_t t;
        t.id_new = c.id_new; t.id_near = ID[c.near_slot]; t.n_near = c.k_total; t.n_rewired = c.n_rw;
        t.removed_id = c.removed_id;
        t.cp_parent = c.cp_parent_id; t.cp_cost = eval_cp ? c.cp_cost : 0.0;
        TRACE[c.iter] = t;
    }
    __syncwarp();
    c.iter += 1;                                                              // :173 / :254
    return kNext;
}

// load the planner record into the warp's context (start / resume)
static __device__ __noinline__ void phase_load(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    WarpCtx& c = *w.c;
    WordSrc& rng = *w.rng;
    const int lane = w.lane;
    const rrtfnd_planner_t* PL = bt->planners + w.p;
    c.gx = PL->goal_x; c.gy = PL->goal_y; c.best_cost = PL->best_cost;
    c.attempts = PL->attempts; c.n_edge_checks = PL->n_edge_checks; c.n_edge_checks_ref = PL->n_edge_checks_ref;
    c.n_scan_nodes = PL->n_scan_nodes; c.n_hops = 0; c.n_tests = 0;
    c.n_slots = PL->n_slots; c.n_live = PL->n_live; c.last_id = PL->last_id; c.iter = PL->iter; c.root_slot = PL->root_slot;
    c.n_finish = PL->n_finish; c.best_len = PL->best_len; c.status = RRTFND_ST_RUNNING;
    c.n_rewired = PL->n_rewired; c.n_removed = PL->n_removed; c.solution_found = PL->solution_found;
    c.n_compactions = PL->n_compactions; c.fn_count = PL->fn_count;
    c.removed_id = -1; c.need_refresh = 0;
    rng.mode = prm->stream_mode;
    rng.words = bt->words ? bt->words + (size_t)w.p * prm->words_per_planner : nullptr;
    rng.n_words = prm->words_per_planner;
    rng.seed = PL->seed; rng.stream = PL->stream_id; rng.cursor = PL->cursor;
    rng.have_blk = false; rng.exhausted = false; rng.blk_idx = 0; rng.b0 = rng.b1 = rng.b2 = rng.b3 = 0;
    const int n_slots = PL->n_slots, best_len = PL->best_len;
    // slot of the cached best path's leaf (ids ascend with slots: binary search around tombstones)
    int best_leaf_slot = -1;
    if (best_len > 0) {
        const int leaf_id = w.BPATH[0];
        int lo = 0, hi = n_slots - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            int m2 = mid;
            while (m2 <= hi && w.ID[m2] < 0) ++m2;
            if (m2 > hi) { hi = mid - 1; continue; }
            const int v = w.ID[m2];
            if (v == leaf_id) { best_leaf_slot = m2; break; }
            if (v < leaf_id) lo = m2 + 1; else hi = mid - 1;
        }
    }
    c.best_leaf_slot = best_leaf_slot;
    // childless live nodes (forced_removal draws among them); kept incrementally afterwards
    int n_childless = 0;
    if (prm->mode == RRTFND_MODE_FN) {
        for (int base = 0; base < n_slots; base += 32) {
            const int s = base + lane;
            n_childless += __popc(__ballot_sync(kFull, s < n_slots && w.NCHILD[s] == 0));
        }
    }
    c.n_childless = n_childless;
    __syncwarp();
}

static __device__ __noinline__ void phase_store(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, unsigned char* smem_base) {
    const Warp w = make_warp(prm, bt, smem_base);
    const WarpCtx& c = *w.c;
    __syncwarp();
    if (w.lane == 0) {
        rrtfnd_planner_t* PL = bt->planners + w.p;
        PL->best_cost = c.best_cost; PL->cursor = w.rng->cursor; PL->attempts = c.attempts;
        PL->n_edge_checks = c.n_edge_checks; PL->n_edge_checks_ref = c.n_edge_checks_ref;
        PL->n_circle_tests += c.n_tests; PL->n_scan_nodes = c.n_scan_nodes; PL->n_chain_hops += c.n_hops;
        PL->n_slots = c.n_slots; PL->n_live = c.n_live; PL->last_id = c.last_id; PL->iter = c.iter;
        PL->root_slot = c.root_slot; PL->n_finish = c.n_finish; PL->best_len = c.best_len; PL->status = c.status;
        PL->n_rewired = c.n_rewired; PL->n_removed = c.n_removed; PL->solution_found = c.solution_found;
        PL->n_compactions = c.n_compactions; PL->fn_count = c.fn_count;
    }
}

// WPB warps (= planners) per CTA; MINW = resident warps per SM the register allocation must allow (16 -> 128 registers,
// 24 -> 80, 32 -> 64); NG = node-grid variant.
template <int WPB, int MINW, bool NG>
__global__ void __launch_bounds__(32 * WPB, MINW / WPB) plan_kernel(const __grid_constant__ rrtfnd_params_t prm,
                                                                    const __grid_constant__ rrtfnd_batch_t bt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (blockIdx.x * WPB + wib >= bt.n_planners) return;      // whole warp leaves; warps never synchronise with each other
    {
        const int n_stages = prm.reserved;
        unsigned char* ws = smem + (size_t)wib * warp_smem_bytes(prm, n_stages);
        unsigned char* bars = ws + (size_t)n_stages * kTileBytes;
        WarpCtx* c = reinterpret_cast<WarpCtx*>(bars + align_up((size_t)n_stages * 8, 16));
        XYStream xs;
        stream_init(xs, ws, bars, n_stages, lane);
        c->xs = xs;
        __syncwarp();
    }
    phase_load(&prm, &bt, smem);
    for (;;) {
        int r = phase_attempt<NG>(&prm, &bt, smem);
        if (r == kRetry) continue;
        if (r == kStop) break;
        r = phase_insert<NG>(&prm, &bt, smem);
        if (r == kRetry) continue;
        if (r == kStop) break;
        if (phase_near<NG>(&prm, &bt, smem) == kStop) break;
        WarpCtx* c = reinterpret_cast<WarpCtx*>(smem + (size_t)wib * warp_smem_bytes(prm, prm.reserved) +
                                                (size_t)prm.reserved * kTileBytes + align_up((size_t)prm.reserved * 8, 16));
        c->removed_id = -1;
        if (prm.mode == RRTFND_MODE_FN && c->fn_count > prm.max_nodes)           // n_of_nodes > max_nodes (:233)
            if (phase_fn_remove(&prm, &bt, smem) == kStop) break;
        if (phase_goal(&prm, &bt, smem) == kStop) break;
    }
    phase_store(&prm, &bt, smem);
}

// Graph.__init__ (src/graph.py:25-32): a tree holding only the start node, id 0.
__global__ void init_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const size_t o = (size_t)p * prm.capacity;
    rrtfnd_planner_t* PL = bt.planners + p;
    bt.xy[2 * o] = PL->start_x; bt.xy[2 * o + 1] = PL->start_y; bt.ecost[o] = 0.0;
    bt.parent[o] = -1; bt.nchild[o] = 0; bt.id[o] = 0; bt.nflags[o] = 0;
    PL->best_cost = CUDART_INF;
    PL->cursor = 0; PL->attempts = 0; PL->n_edge_checks = 0; PL->n_edge_checks_ref = 0; PL->n_circle_tests = 0;
    PL->n_scan_nodes = 0; PL->n_chain_hops = 0;
    PL->n_slots = 1; PL->n_live = 1; PL->last_id = 0; PL->iter = 0; PL->root_slot = 0; PL->n_finish = 0;
    PL->best_len = 0; PL->status = RRTFND_ST_RUNNING; PL->n_rewired = 0; PL->n_removed = 0;
    PL->solution_found = 0; PL->n_compactions = 0; PL->fn_count = 1; PL->reserved = 0;
    if ((prm.flags & RRTFND_FLAG_NODE_GRID) && bt.node_grid.head) {               // heads were reset to -1 by the host wrapper
        const rrtfnd_node_grid_t& g = bt.node_grid;
        const int c = grid_cell(PL->start_y, g.y0, g.inv_cell, g.ny) * g.nx + grid_cell(PL->start_x, g.x0, g.inv_cell, g.nx);
        g.head[(size_t)p * g.nx * g.ny + c] = 0;
        g.next[o] = -1;
    }
}

constexpr size_t kMaxSmem = 227 * 1024;

template <int WPB, int MINW, bool NG = false>
static int launch_plan(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, cudaStream_t st) {
    const size_t bytes = warp_smem_bytes(*prm, prm->reserved) * WPB;
    if (bytes > kMaxSmem) return RRTFND_E_SMEM;
    auto kern = plan_kernel<WPB, MINW, NG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    kern<<<(bt->n_planners + WPB - 1) / WPB, 32 * WPB, bytes, st>>>(*prm, *bt);
    return (int)cudaGetLastError();
}

// Launch geometry. prm->reserved = warps (= planners) per CTA (1, 2, 4) + 16 * xy-ring stages (2 or 4) + 256 *
// resident warps per SM to allocate registers for (16, 24, 32); a zero field = automatic: one planner per CTA spreads
// a small batch evenly over the SMs (an SM holds at most 32 CTAs), larger batches need more warps per CTA to reach
// the SM's warp limit; few resident warps get more registers and the deeper ring.
struct PlanCfg { int wpb, stages, minw; };
static PlanCfg choose_cfg(const rrtfnd_params_t& prm, int n_planners) {
    PlanCfg c{prm.reserved & 15, (prm.reserved >> 4) & 15, (prm.reserved >> 8) & 255};
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (c.wpb != 1 && c.wpb != 2 && c.wpb != 4) c.wpb = n_planners <= 32 * sms ? 1 : 2;
    // measured on B200 (profiles/): trading registers for residency (80 / 64 registers) spills and halves throughput,
    // and a ring deeper than 2 tiles does not pay once a dozen warps share an SM
    if (c.minw != 16 && c.minw != 24 && c.minw != 32) c.minw = 16;
    if (c.stages == 1) c.stages = 0;            // 1 = no ring: the scans use plain loads
    else if (c.stages != 2 && c.stages != 4) c.stages = 2;
    return c;
}

template <int WPB>
static int launch_wpb(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, int minw, cudaStream_t st) {
    switch (minw) {
        case 16: return launch_plan<WPB, 16>(prm, bt, st);
        case 24: return launch_plan<WPB, 24>(prm, bt, st);
        default: return launch_plan<WPB, 32>(prm, bt, st);
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
    if (!bt->planners || !bt->xy || !bt->ecost || !bt->parent || !bt->nchild || !bt->id || !bt->nflags || !bt->finish ||
        !bt->best_path || (prm->n_obst > 0 && !bt->obst)) return RRTFND_E_BADARG;
    if (bt->grid.nx > 0 && (!bt->grid.seg_start || !bt->grid.seg_items || !bt->grid.pt_start || !bt->grid.pt_items ||
                            bt->grid.ny <= 0)) return RRTFND_E_BADARG;
    return 0;
}

extern "C" int64_t rrtfnd_plan_smem_bytes(const rrtfnd_params_t* prm) {
    if (!prm) return RRTFND_E_BADARG;
    const size_t b = warp_smem_bytes(*prm, 4);
    return b <= kMaxSmem ? (int64_t)b : (int64_t)RRTFND_E_SMEM;
}

extern "C" int rrtfnd_plan_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, void* stream) {
    int rc = check_params(prm, bt);
    if (rc) return rc;
    rrtfnd_params_t q = *prm;
    if (!(q.flags & RRTFND_FLAG_TRACE)) q.trace_cap = 0;
    rrtfnd_batch_t b = *bt;
    if (!(q.flags & RRTFND_FLAG_TRACE)) b.trace = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    const PlanCfg cfg = choose_cfg(q, b.n_planners);
    q.reserved = cfg.stages;                    // the kernel reads its ring depth here
    if (q.flags & RRTFND_FLAG_NODE_GRID) {      // large single trees: node-grid kernel, one planner per CTA
        const rrtfnd_node_grid_t& g = b.node_grid;
        if (q.mode == RRTFND_MODE_FN || g.nx <= 0 || g.ny <= 0 || !g.head || !g.next || !g.cand || g.cand_cap < 32 || !(g.inv_cell > 0.0))
            return RRTFND_E_BADARG;
        return launch_plan<1, 16, true>(&q, &b, (cudaStream_t)stream);
    }
    switch (cfg.wpb) {
        case 1: return launch_wpb<1>(&q, &b, cfg.minw, st);
        case 2: return launch_wpb<2>(&q, &b, cfg.minw, st);
        case 4: return launch_wpb<4>(&q, &b, cfg.minw, st);
        default: return RRTFND_E_BADARG;
    }
}

extern "C" int rrtfnd_init_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, void* stream) {
    if (!prm || !bt || bt->n_planners <= 0 || !bt->planners || !bt->xy || !bt->nflags) return RRTFND_E_BADARG;
    const int nt = 128;
    if ((prm->flags & RRTFND_FLAG_NODE_GRID) && bt->node_grid.head) {
        const rrtfnd_node_grid_t& g = bt->node_grid;
        if (g.nx <= 0 || g.ny <= 0 || !g.next) return RRTFND_E_BADARG;
        cudaError_t e = cudaMemsetAsync(g.head, 0xff, (size_t)bt->n_planners * g.nx * g.ny * sizeof(int32_t), (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    init_kernel<<<(bt->n_planners + nt - 1) / nt, nt, 0, (cudaStream_t)stream>>>(*prm, *bt);
    return (int)cudaGetLastError();
}
