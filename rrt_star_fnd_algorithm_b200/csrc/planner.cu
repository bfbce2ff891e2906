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
#include <stdlib.h>

#include <atomic>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "cull.cuh"
#include "grid.cuh"
#include "rng.cuh"
#include "stream.cuh"

// annulus schedule of the nearest-visible search: squared-distance multiplier of the first and of the later annuli.
// Measured on B200 (profiles/r01_phase_clocks.md): 2 / 2 is the best of 2/2, 8/4, 24/8, 64/8 - wider annuli save scans but
// stage more, longer segments, and the segment tests are what the search costs.
#ifndef RRTFND_ANN_FIRST
#define RRTFND_ANN_FIRST 2
#endif
#ifndef RRTFND_ANN_GROW
#define RRTFND_ANN_GROW 2
#endif

// Measured on B200 in round 2 and removed (profiles/r02_variants.md): the winner's rank counted inside the near-set scan instead
// of a separate pass (-9 %: six more instructions in the hottest loop cost more than the pass they save), candidates split over
// idle lanes (+-0), a first annulus bounded by the lanes' own minima (-4 %), persistent CTAs with the obstacle tables in shared
// memory (hung), the segment test and the chain walk fused into one loop (-26 %), and an annulus search that stages a whole
// annulus, narrows it when it holds more than 96 nodes and tests it in key order 32 at a time with an early exit (-13 % on
// 16,384 planners, -32 % on 2,048: the circle tests it saves are few - segments stop at their first blocking circle - and the
// extra scans and the threshold bisection cost more; profiles/r02_s8_raw.txt).
// Graph.get_cost walks of the candidate chunks: over the parent[] / ecost[] arrays (two loads per hop, one hop per memory
// round trip), or - kernels instantiated with WREC, the node-grid kernel of large single trees - over the 16-byte walk
// records with the self-healing grandparent hint (walk_records: one load per hop, two hops per round trip while the hints
// hold). Measured on B200 (profiles/r02_variants.md): +1 % on c3's 5,000-node trees, -8 % on c2's L1-resident 1,000-node
// trees, so the batch kernels keep the arrays; parent chains of hundreds of hops (config 4) are where the records pay.
// choose-parent's re-summed chain (src/algorithm.py:849) is prepared by all lanes at once (1: +4.7 % on c3, measured) or
// re-summed at each acceptance (0, round 1's form).
#ifndef RRTFND_CP_PARALLEL
#define RRTFND_CP_PARALLEL 1
#endif
// resident warps per SM the lean instances' register allocation admits: 16 -> 128 registers, 20 -> 96, 24 -> 80
#ifndef RRTFND_LEAN_MINW
#define RRTFND_LEAN_MINW 16
#endif

// Diagnostic build (-DRRTFND_PHASE_CLOCKS): lane 0 of every warp accumulates clock64() deltas per phase of the iteration,
// summed over the grid into g_phase_clocks and read back by rrtfnd_debug_phase_clocks(). Not compiled into the product.
#ifdef RRTFND_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[16];
#define PH_MARK(i) do { __syncwarp(); const long long ph_now = clock64(); if (lane == 0) ph_acc[(i)] += ph_now - ph_last; ph_last = ph_now; } while (0)
#else
#define PH_MARK(i) do { } while (0)
#endif

namespace rrtfnd {

constexpr int kReserveWords = 64;   // must equal RESERVE_WORDS in _structs.py / ORC_RESERVE_WORDS
constexpr int kStage = 160;         // slots staged per warp: < 32 left over + one 128-node block of the scan
constexpr int kIntMax = 0x7fffffff;
// edge costs of the new node's parent chain cached per warp (choose-parent re-sums them): 96 hops cover the batch workloads'
// trees; a single large tree (node grid, config 4) has chains of several hundred hops - without the cache every acceptance of
// choose-parent re-walked them at memory latency (44 % of an extension at 10^5 nodes, profiles/r02_c4.md)
constexpr int kChainCapBatch = 96, kChainCapLarge = 1024;
__host__ __device__ inline int chain_cap(const rrtfnd_params_t& prm) { return (prm.flags & RRTFND_FLAG_NODE_GRID) ? kChainCapLarge : kChainCapBatch; }

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int bitmap_words(const rrtfnd_params_t& prm) { return prm.capacity / 32 + 1; }
// Compaction keeps a live bitmap + prefix counts in shared memory: always for FN; for the other modes (tombstones only
// after FND surgery) while the capacity is moderate. Without it a full slot array is a capacity error.
__host__ __device__ inline bool can_compact(const rrtfnd_params_t& prm) { return prm.mode == RRTFND_MODE_FN || prm.capacity <= 32768; }

// prepared attempts of one planner (see "speculative attempts" below)
struct SampleBatch {
    double qx[32], qy[32];
    unsigned long long c0;     // word cursor of entry 0
    uint32_t free_mask;        // entries whose sample is not occupied
    int pos, len;              // next entry to consume, number of valid entries
    int last_goal;             // the last valid entry is a goal-bias hit (2 words instead of 6)
};

// shared memory of one warp: xy ring (n_stages tiles) | attempt batch | chain[chain_cap] | stage[kStage] | rwlist[near_cap]
// (re-parented slots of one iteration; overflow goes to nflags) | mbarriers | word source | [live bitmap + prefix counts].
// The compaction's live bitmap and prefix counts are only alive inside compact_tree(), at an attempt boundary, when the
// ring, the attempt batch and the per-iteration scratch hold nothing: they ALIAS that region (scratch_bytes) when they fit
// and get their own space behind the word source otherwise. Every kilobyte not allocated here is L1 for the tree gathers.
__host__ __device__ inline size_t scratch_bytes(const rrtfnd_params_t& prm, int n_stages) {
    return (size_t)n_stages * kTileBytes + align_up(sizeof(SampleBatch), 16) + (size_t)chain_cap(prm) * 8 + kStage * 4 +
           align_up((size_t)prm.near_cap * 4, 16);
}
__host__ __device__ inline size_t compact_bytes(const rrtfnd_params_t& prm) {
    return can_compact(prm) ? 2 * align_up((size_t)bitmap_words(prm) * 4, 16) : 0;
}
// RRT*-FN keeps one bit per slot "live and childless" in shared memory: forced_removal's random.choice(childless_nodes)
// (src/algorithm.py:812-813) needs the r-th childless node in id order, which is a popcount prefix over these words instead of
// a scan of the slot array per removal (16 % of a config-2 extension, profiles/r02_phase_clocks.md).
__host__ __device__ inline size_t childless_bytes(const rrtfnd_params_t& prm) {
    return prm.mode == RRTFND_MODE_FN ? align_up((size_t)bitmap_words(prm) * 4, 16) : 0;
}
__host__ __device__ inline size_t warp_smem_bytes(const rrtfnd_params_t& prm, int n_stages) {
    size_t o = scratch_bytes(prm, n_stages) + align_up((size_t)n_stages * 8, 16) + align_up(sizeof(WordSrc), 16) + childless_bytes(prm);
    if (compact_bytes(prm) > scratch_bytes(prm, n_stages)) o += compact_bytes(prm);
    return align_up(o, 128);
}

// ------------------------------------------------------------------------------------ speculative attempts
// The attempts of one planner are a serial chain in the reference (draw, test occupancy, on rejection draw again), and
// on dense maps most of them end at the occupancy test (c3: 1.55 of 2.55 attempts per extension). The word stream is
// addressable (Philox is counter based, the explicit buffer is an array), so lane j prepares the attempt that starts
// 6 j words further on - its three random() values, its sample and that sample's occupancy - and the warp then consumes
// the prepared attempts in order. Entry j is what the serial chain produces iff every earlier attempt consumed six
// words and its own header checks pass, so the batch ends after the first goal-bias hit (that attempt consumes two
// words) and before the first attempt whose word-reserve or attempt-cap check would fail; any other draw from the stream
// (forced_removal) moves the cursor away from the batch, which is noticed when the next entry is about to be used.

__device__ __forceinline__ uint32_t word_at(const WordSrc& r, unsigned long long i, unsigned long long& cached, uint4& blk) {
    if (r.mode == RRTFND_STREAM_WORDS) return r.words[i];
    const unsigned long long b = i >> 2;
    if (b != cached) { blk = philox4x32_10(b, r.stream, r.seed); cached = b; }
    const unsigned k = (unsigned)i & 3u;
    return k == 0 ? blk.x : k == 1 ? blk.y : k == 2 ? blk.z : blk.w;
}
__device__ __forceinline__ double random_from(uint32_t w0, uint32_t w1) {              // CPython random(), as next_random()
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}

// Map.is_occupied_c for a point of this lane alone (lanes hold different points): bit 0 = occupied, the rest = circles tested
template <bool RECT>
static __device__ __noinline__ int point_occupied_lane(const double2* tab, const rrtfnd_obst_grid_t* g, int M, int full_cells,
                                                       double px, double py, double node_radius) {
    const bool cull = full_cells > 0;
    int j = 0, e = M;
    if (cull) {
        const int c = grid_cell(py, g->y0, g->inv_cell, g->ny) * g->nx + grid_cell(px, g->x0, g->inv_cell, g->nx);
        j = __ldg(g->pt_start + c);
        e = __ldg(g->pt_start + c + 1);
    }
    int tests = 0;
    for (; j < e; ++j) {
        double ox, oy, r, K;
        load_obst(tab, cull ? __ldg(g->pt_items + j) : j, ox, oy, r, K);
        ++tests;
        if (point_in_inflated<RECT>(px, py, ox, oy, r, node_radius, K)) return 2 * tests + 1;
    }
    return 2 * tests;
}

// returns the number of circles this lane tested
template <bool RECT>
static __device__ __noinline__ int fill_batch(SampleBatch& sb, const WordSrc& rng, const rrtfnd_params_t& prm, const Obst& ob,
                                              double gx, double gy, long long attempts, int lane) {
    const unsigned long long c0 = rng.cursor, cj = c0 + 6ull * (unsigned)lane;
    bool valid = true;                                    // the attempt header's checks, as they would read for attempt j
    if (rng.mode == RRTFND_STREAM_WORDS) valid = (long long)cj + kReserveWords <= rng.n_words;
    if (prm.max_attempts > 0) valid = valid && attempts + lane < prm.max_attempts;
    double qx = gx, qy = gy;
    bool hit = false, occ = false;
    if (valid) {
        unsigned long long cached = ~0ull;
        uint4 blk = make_uint4(0, 0, 0, 0);
        const uint32_t w0 = word_at(rng, cj, cached, blk), w1 = word_at(rng, cj + 1, cached, blk);
        hit = random_from(w0, w1) < prm.bias;             // Graph.random_node (src/graph.py:96-98)
        if (!hit) {
            const uint32_t w2 = word_at(rng, cj + 2, cached, blk), w3 = word_at(rng, cj + 3, cached, blk);
            const uint32_t w4 = word_at(rng, cj + 4, cached, blk), w5 = word_at(rng, cj + 5, cached, blk);
            qx = dadd(prm.x_lo, dmul(dsub(prm.x_hi, prm.x_lo), random_from(w2, w3)));
            qy = dadd(prm.y_lo, dmul(dsub(prm.y_hi, prm.y_lo), random_from(w4, w5)));
        }
    }
    const unsigned bad = __ballot_sync(kFull, !valid), hits = __ballot_sync(kFull, valid && hit);
    int len = bad ? __ffs(bad) - 1 : 32;
    int last_goal = 0;
    if (hits && __ffs(hits) - 1 < len) { len = __ffs(hits); last_goal = 1; }
    int tests = 0;
    if (lane < len) {
        const int r = point_occupied_lane<RECT>(ob.tab, ob.g, ob.M, ob.full_cells, qx, qy, prm.node_radius);
        occ = r & 1; tests = r >> 1;
    }
    const unsigned free_mask = __ballot_sync(kFull, lane < len && !occ);
    sb.qx[lane] = qx; sb.qy[lane] = qy;
    sb.c0 = c0; sb.free_mask = free_mask; sb.pos = 0; sb.len = len; sb.last_goal = last_goal;   // every lane: same values
    __syncwarp();
    return tests;
}

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

// ---- walk records (rrtfnd_batch_t.walk): one 16-byte record per node = (ecost, parent, grandparent guess) --------------
// Graph.get_cost (src/graph.py:41-46) is a pointer chase: ~40 dependent hops per candidate, each a round trip to L2 or
// HBM. The record packs what one hop needs into ONE 16-byte load, and its third field lets a walk request the parent's
// and the grandparent's records together: when the parent's record confirms the guess, two hops cost one round trip.
// The guess is only a hint - re-parenting a node leaves its children's guesses stale; a walk that finds one wrong takes
// the slow step and repairs it - so results never depend on it. ecost[] / parent[] stay the truth (ABI, FND operators).
struct WRec { double ecost; int parent; int gp; };
__device__ __forceinline__ WRec wrec_load(const rrtfnd_walk_t* W, int s) {
    const double2 v = *reinterpret_cast<const double2*>(W + s);
    WRec r; r.ecost = v.x;
    const long long b = __double_as_longlong(v.y);
    r.parent = (int)(unsigned)(b & 0xffffffffll); r.gp = (int)(b >> 32);
    return r;
}
__device__ __forceinline__ void wrec_store(rrtfnd_walk_t* W, int s, double ecost, int parent, int gp) {
    const long long b = ((long long)gp << 32) | (unsigned)parent;
    *reinterpret_cast<double2*>(W + s) = make_double2(ecost, __longlong_as_double(b));
}
// records of slots [0, n_slots) from the truth arrays (start of a call, after a compaction)
static __device__ __noinline__ void wrec_rebuild(rrtfnd_walk_t* W, const int32_t* PARENT, const double* ECOST, int n_slots, int lane) {
    for (int s = lane; s < n_slots; s += 32) {
        const int p = PARENT[s];
        wrec_store(W, s, ECOST[s], p, p >= 0 ? PARENT[p] : -1);
    }
    __syncwarp();
}

// Graph.get_cost over the walk records: leaf -> root, `acc` the running sum to start from; `rec_chain` (one lane per chunk at
// most) receives the addends in order. Returns the hops walked. One round = the parent's and the guessed grandparent's
// records requested together; a confirmed guess advances two hops for one memory round trip.
__device__ __forceinline__ int walk_records(rrtfnd_walk_t* W, int from, double* rec_chain, double& acc, int kChainCap) {
    int depth = 0, v = from;
    WRec rv = wrec_load(W, v);
    while (rv.parent >= 0) {
        const WRec rp = wrec_load(W, rv.parent);
        WRec rg = rp;
        if (rv.gp >= 0) rg = wrec_load(W, rv.gp);
        if (rec_chain && depth < kChainCap) rec_chain[depth] = rv.ecost;
        acc = dadd(acc, rv.ecost); ++depth;
        if (rp.parent >= 0 && rv.gp == rp.parent) {          // the guess holds: the parent's edge as well, on to the grandparent
            if (rec_chain && depth < kChainCap) rec_chain[depth] = rp.ecost;
            acc = dadd(acc, rp.ecost); ++depth;
            v = rv.gp; rv = rg;
        } else {
            if (rp.parent >= 0) wrec_store(W, v, rv.ecost, rv.parent, rp.parent);   // repair the hint (a benign same-value race)
            v = rv.parent; rv = rp;
        }
    }
    return depth;
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
template <bool NG, bool RING>
__device__ __forceinline__ void fetch_tile(XYStream& xs, const double2* XY, const int32_t* list, int n_items, int t, int lane,
                                           double2 (&v)[4], int (&sl)[4]) {
    const int base = t * kTileNodes;
    const double2 nan2 = make_double2(CUDART_NAN, CUDART_NAN);
    if (NG && list) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = base + 32 * u + lane; sl[u] = i < n_items ? list[i] : -1; }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = sl[u] >= 0 ? XY[sl[u]] : nan2;
    } else if (!RING) {                 // plain coalesced 16-byte loads, four in flight per lane
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int s = base + 32 * u + lane; sl[u] = s; v[u] = s < n_items ? XY[s] : nan2; }
    } else {
        const double2* tile = stream_wait(xs, t);
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int s = base + 32 * u + lane; sl[u] = s; v[u] = s < n_items ? tile[32 * u + lane] : nan2; }
        stream_release(xs, XY, n_items, t, lane);
    }
}


// WPB warps (= planners) per CTA; MINW = resident warps per SM the register allocation must allow (16 -> 128
// registers, 24 -> 80, 32 -> 64): small batches get registers, large ones get occupancy.
// MODE >= 0 is a LEAN instance: that planner mode with choose-parent evaluated but not committed and no trace — the
// configuration batches run in — so the other modes' code, the trace writes and the second near-set pass are not even
// compiled in (the kernel's instruction footprint is what limits it, profiles/). MODE = -1 takes everything at run time.
template <int WPB, int MINW, bool NG, int MODE, bool RING, bool WREC = NG, bool RECT = (MODE < 0), bool XORD = true>
__global__ void __launch_bounds__(32 * WPB, MINW / WPB) plan_kernel(const rrtfnd_params_t prm, const __grid_constant__ rrtfnd_batch_t bt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* const warp_base = smem;
    const double* const obst_tab = bt.obst;
    const rrtfnd_obst_grid_t* const obst_grid = &bt.grid;
    const int p = blockIdx.x * WPB + wib;
    if (p >= bt.n_planners) return;                       // whole warp leaves; warps never synchronise with each other
    // ---- sliced launches (launch_sliced below): this launch grows every planner up to prm.iter_num, one slice of the run; the
    // launch of the next slice is allowed onto the SMs as soon as every CTA of this one has STARTED (programmatic dependent
    // launch), so the slots that the last planners of a slice leave idle are taken by the next slice instead of waiting for
    // the slowest planner. A planner's slices are ordered through a ticket in its own record, not through the stream.
    const int slice = prm.reserved >> 8, n_slices = (prm.reserved >> 4) & 15;
    if (n_slices > 1) {
        int* ticket = &bt.planners[p].reserved;
        if (lane == 0) {
            if (slice == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ticket), "r"(bt.reserved) : "memory");
            else {
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ticket) : "memory");
                    if (seen - (bt.reserved + slice) < 0) __nanosleep(2000);
                } while (seen - (bt.reserved + slice) < 0);
            }
        }
        __syncwarp();
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    {
    const int C = prm.capacity;
    const unsigned lt_mask = (1u << lane) - 1u;

    // per-planner global arrays
    double2* XY = reinterpret_cast<double2*>(bt.xy) + (size_t)p * C;
    double* ECOST = bt.ecost + (size_t)p * C;
    int32_t* PARENT = bt.parent + (size_t)p * C;
    int32_t* NCHILD = bt.nchild + (size_t)p * C;
    int32_t* ID = bt.id + (size_t)p * C;
    uint8_t* NFLAG = bt.nflags + (size_t)p * C;
    int32_t* FINISH = bt.finish + (size_t)p * prm.finish_cap;
    int32_t* BPATH = bt.best_path + (size_t)p * prm.path_cap;
    rrtfnd_planner_t* PL = bt.planners + p;
    rrtfnd_walk_t* WR = WREC ? bt.walk + (size_t)p * C : nullptr;
    constexpr bool kLean = MODE >= 0;
    // Prepared attempt batches pay where rejected attempts are frequent and nothing else draws from the stream between
    // attempts; RRT*-FN draws in forced_removal after every extension once the cap is reached, which discards the batch
    // (measured: c3 +10 %, c2 -4 %), so its lean instance keeps the one-at-a-time path.
    constexpr bool kBatchAttempts = MODE != RRTFND_MODE_FN;
    const int mode = kLean ? MODE : prm.mode;
    rrtfnd_trace_t* TRACE = (!kLean && bt.trace) ? bt.trace + (size_t)p * prm.trace_cap : nullptr;

    // per-warp shared memory
    const int n_stages = prm.reserved & 15;            // resolved by the host wrapper (2 or 4)
    unsigned char* ws = warp_base + (size_t)wib * warp_smem_bytes(prm, n_stages);
    unsigned char* wp = ws + (size_t)n_stages * kTileBytes;
    SampleBatch& sb = *reinterpret_cast<SampleBatch*>(wp);   // prepared attempts; identical scalars written by every lane
    wp += align_up(sizeof(SampleBatch), 16);
    double* chain = reinterpret_cast<double*>(wp);     // edge costs along the new node's parent chain, leaf side first
    constexpr int kChainCap = NG ? kChainCapLarge : kChainCapBatch;
    wp += kChainCap * 8;
    int* stage = reinterpret_cast<int*>(wp);
    int* rwlist = stage + kStage;
    wp = reinterpret_cast<unsigned char*>(rwlist + align_up((size_t)prm.near_cap * 4, 16) / 4);
    XYStream xs;
    stream_init(xs, ws, wp, n_stages, lane);
    wp += align_up((size_t)n_stages * 8, 16);
    WordSrc& rng = *reinterpret_cast<WordSrc*>(wp);    // identical values written by every lane
    wp += align_up(sizeof(WordSrc), 16);
    uint32_t* cb = reinterpret_cast<uint32_t*>(wp);    // RRT*-FN: bit s = slot s is live and childless (lane 0 maintains it)
    wp += childless_bytes(prm);
    // compaction scratch: over the (then idle) ring / batch / scratch region when it fits, else its own space
    uint32_t* bm = reinterpret_cast<uint32_t*>(compact_bytes(prm) > scratch_bytes(prm, n_stages) ? wp : ws);
    uint32_t* pre = bm + align_up((size_t)bitmap_words(prm) * 4, 16) / 4;

    const Obst ob = make_obst(obst_tab, prm.n_obst, obst_grid);
    // node grid (NG kernels only)
    const rrtfnd_node_grid_t& ng = bt.node_grid;
    int32_t* HEAD = NG ? ng.head + (size_t)p * ng.nx * ng.ny : nullptr;
    int32_t* NEXT = NG ? ng.next + (size_t)p * C : nullptr;
    int32_t* CAND = NG ? ng.cand + (size_t)p * 2 * ng.cand_cap : nullptr;
    // all nodes within rho cells' width of a point lie in the (2 rho + 1)^2 cells around the point's cell
    auto ng_window = [&](double x, double y, int rho, int& cx0, int& cx1, int& cy0, int& cy1) {
        const int cx = grid_cell(x, ng.x0, ng.inv_cell, ng.nx), cy = grid_cell(y, ng.y0, ng.inv_cell, ng.ny);
        cx0 = max(cx - rho, 0); cx1 = min(cx + rho, ng.nx - 1); cy0 = max(cy - rho, 0); cy1 = min(cy + rho, ng.ny - 1);
        return cx0 == 0 && cy0 == 0 && cx1 == ng.nx - 1 && cy1 == ng.ny - 1;
    };
    auto ng_complete_d2 = [&](int rho) { const double w = rho * ng.cell; return w * w * (1.0 - 1e-9); };

    // ---- planner state: warp-uniform, replicated in the registers of every lane
    const double gx = PL->goal_x, gy = PL->goal_y;
    double best_cost = PL->best_cost;
    long long attempts = PL->attempts, n_edge_checks = PL->n_edge_checks, n_edge_checks_ref = PL->n_edge_checks_ref;
    long long n_scan_nodes = PL->n_scan_nodes;
    int n_slots = PL->n_slots, n_live = PL->n_live, last_id = PL->last_id, iter = PL->iter, root_slot = PL->root_slot;
    int n_finish = PL->n_finish, best_len = PL->best_len, status = RRTFND_ST_RUNNING;
    int n_rewired = PL->n_rewired, n_removed = PL->n_removed, solution_found = PL->solution_found;
    int n_compactions = PL->n_compactions;
    int fn_count = PL->fn_count;             // n_of_nodes of RRT_star_FN (:209): counts this call's nodes, not the live ones
    long long my_hops = 0, my_tests = 0;     // per-lane counters, summed at the end

    rng.mode = prm.stream_mode;
    rng.words = bt.words ? bt.words + (size_t)p * prm.words_per_planner : nullptr;
    rng.n_words = prm.words_per_planner;
    rng.seed = PL->seed; rng.stream = PL->stream_id; rng.cursor = PL->cursor;
    rng.have_blk = false; rng.exhausted = false; rng.blk_idx = 0; rng.b0 = rng.b1 = rng.b2 = rng.b3 = 0;
    sb.pos = 0; sb.len = 0;
    __syncwarp();
    if (WREC) wrec_rebuild(WR, PARENT, ECOST, PL->n_slots, lane);

    // slot of the cached best path's leaf (ids ascend with slots: binary search around tombstones)
    int best_leaf_slot = -1;
    if (best_len > 0) {
        const int leaf_id = BPATH[0];
        int lo = 0, hi = n_slots - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            int m2 = mid;
            while (m2 <= hi && ID[m2] < 0) ++m2;
            if (m2 > hi) { hi = mid - 1; continue; }
            const int v = ID[m2];
            if (v == leaf_id) { best_leaf_slot = m2; break; }
            if (v < leaf_id) lo = m2 + 1; else hi = mid - 1;
        }
    }
    // childless live nodes (forced_removal draws among them); kept incrementally afterwards
    int n_childless = 0;
    auto childless_rebuild = [&](int slots) {
        int n = 0;
        for (int w = 0; w < bitmap_words(prm); ++w) {
            const int s = w * 32 + lane;
            const unsigned m = __ballot_sync(kFull, s < slots && NCHILD[s] == 0);
            if (lane == 0) cb[w] = m;
            n += __popc(m);
        }
        __syncwarp();
        return n;
    };
    if (mode == RRTFND_MODE_FN) n_childless = childless_rebuild(n_slots);

    const double r2 = dmul(prm.radius, prm.radius);
    const double goal_thresh = dmul(2.0, prm.goal_node_radius);
    const bool eval_cp = kLean || (prm.flags & (RRTFND_FLAG_EVAL_CHOOSE_PARENT | RRTFND_FLAG_COMMIT_CHOOSE_PARENT)) != 0;
    const bool commit_cp = !kLean && (prm.flags & RRTFND_FLAG_COMMIT_CHOOSE_PARENT) != 0;
    const double2 nan2 = make_double2(CUDART_NAN, CUDART_NAN);
#ifdef RRTFND_PHASE_CLOCKS
    __shared__ long long ph_smem[WPB][12];
    long long* ph_acc = ph_smem[wib];
    if (lane == 0) for (int i = 0; i < 12; ++i) ph_acc[i] = 0;
    long long ph_last = clock64();
    __syncwarp();
#endif

    for (;;) {
        PH_MARK(9);                                                               // tail of the previous iteration / rejected attempt
        // ================================================================ A. attempt header
        if (iter >= prm.iter_num) { status = RRTFND_ST_DONE; break; }
        if (prm.max_attempts > 0 && attempts >= prm.max_attempts) { status = RRTFND_ST_ATTEMPT_CAP; break; }
        if (prm.stream_mode == RRTFND_STREAM_WORDS && (long long)rng.cursor + kReserveWords > prm.words_per_planner) {
            status = RRTFND_ST_NEED_WORDS; break;
        }
        if (n_slots >= C) {
            if (n_live >= n_slots || !can_compact(prm)) { status = RRTFND_ST_CAPACITY; break; }
            // ---- A'. stable in-place compaction (FN only; rare)
            compact_tree(XY, ECOST, PARENT, NCHILD, ID, NFLAG, FINISH, n_finish, n_slots, bm, pre, &root_slot, &best_leaf_slot, lane);
            n_slots = n_live;
            ++n_compactions;
            sb.pos = 0; sb.len = 0;       // the bitmap may have lain over the prepared attempts: prepare them again
            __syncwarp();
            if (WREC) wrec_rebuild(WR, PARENT, ECOST, n_slots, lane);
            if (mode == RRTFND_MODE_FN) n_childless = childless_rebuild(n_slots);
        }
        // ================================================================ A/B. Graph.random_node (src/graph.py:96-98) and
        // Map.is_occupied_c (src/map.py:38-47; rejection at src/algorithm.py:39-40)
        double qx, qy;
        bool free_sample;
        if (kBatchAttempts) {                             // through the prepared attempts
            int pos = sb.pos, len = sb.len;
            const bool stale = pos < len && rng.cursor != sb.c0 + 6ull * (unsigned)pos;   // something else drew from the stream
            __syncwarp();
            if (pos >= len || stale) { my_tests += fill_batch<RECT>(sb, rng, prm, ob, gx, gy, attempts, lane); pos = 0; len = sb.len; }
            const unsigned rest = sb.free_mask >> pos;
            free_sample = rest != 0;
            const int e = free_sample ? pos + __ffs(rest) - 1 : len - 1;      // occupied entries before it are rejected attempts
            const unsigned long long c_end = sb.c0 + 6ull * (unsigned)e + ((e == len - 1 && sb.last_goal) ? 2ull : 6ull);
            qx = sb.qx[e]; qy = sb.qy[e];
            __syncwarp();
            attempts += e - pos + 1;
            rng.cursor = c_end;
            sb.pos = e + 1;
            __syncwarp();
        } else {                                          // one attempt at a time
            ++attempts;
            if (next_random(rng) < prm.bias) { qx = gx; qy = gy; }
            else {
                qx = dadd(prm.x_lo, dmul(dsub(prm.x_hi, prm.x_lo), next_random(rng)));
                qy = dadd(prm.y_lo, dmul(dsub(prm.y_hi, prm.y_lo), next_random(rng)));
            }
            free_sample = !point_occupied_warp<RECT>(ob, qx, qy, prm.node_radius, lane, my_tests);
        }
        PH_MARK(0);                                                               // sample + occupancy
        if (!free_sample) continue;                                               // src/algorithm.py:39-40

        // ================================================================ C. nearest visible node (:666-699)
        // The reference asks the k-d tree for the 1st, 2nd, 3rd ... nearest node until one is visible. Round 0 here is
        // that first query (plain argmin). If it is blocked, the remaining order is consumed in annuli of doubling
        // squared radius: every node of an annulus is tested (one lane per segment) and the lowest (d2, slot) key among
        // the visible ones is the node the reference would have stopped at, because everything in earlier annuli was
        // blocked. A boxed-in sample therefore costs O(log N) scans instead of O(N).
        int near_slot = -1;
        {
            double bd = CUDART_INF; int bs = kIntMax;
            // what the scans below run over: the whole tree (TMA ring), or - node grid - the nodes of a window of cells
            const int32_t* list = nullptr;
            int n_items = n_slots, rho = 1;
            bool covers_all = true;
            const bool use_ng = NG && n_live >= ng.min_nodes;
            for (;;) {      // (node grid: widen the window until the nearest node is certain; one pass when scanning)
                if (use_ng) {
                    int cx0, cx1, cy0, cy1;
                    covers_all = ng_window(qx, qy, rho, cx0, cx1, cy0, cy1);
                    const int n = ng_gather(HEAD, NEXT, ng.nx, cx0, cx1, cy0, cy1, CAND, ng.cand_cap, lane);
                    if (n > ng.cand_cap) { list = nullptr; n_items = n_slots; covers_all = true; }
                    else { list = CAND; n_items = n; }
                }
                bd = CUDART_INF; bs = kIntMax;
                const int n_tiles = (n_items + kTileNodes - 1) / kTileNodes;
                if (RING && !list) stream_begin(xs, XY, n_slots, lane);
                for (int t = 0; t < n_tiles; ++t) {
                    double2 v[4]; int sl[4];
                    fetch_tile<NG, RING>(xs, XY, list, n_items, t, lane, v, sl);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double d = dist2(v[u].x, v[u].y, qx, qy);                 // NaN for tombstones / padding
                        if (d < bd || (NG && d == bd && sl[u] < bs)) { bd = d; bs = sl[u]; }
                    }
                }
                warp_argmin(bd, bs);
                PH_MARK(1);                                                       // nearest scan
                n_scan_nodes += list ? n_items : n_live;
                if (!list || covers_all || (bs != kIntMax && bd <= ng_complete_d2(rho))) break;
                rho *= 2;
            }
            if (bs == kIntMax) continue;                                          // empty tree (cannot happen)
            if (bd == 0.0) {                                                      // G.id_vertex[vertex] hit (:676-678)
                const double2 e = XY[bs];
                bool exact = (e.x == qx && e.y == qy);
                if (!exact) exact = any_exact_match(XY, n_slots, qx, qy, lane);   // d2 underflowed to 0 without equality
                if (exact) continue;                                              // rejected at :45
            }
            // through_obstacle(Line(q_rand, node)) — p1 = sample, p2 = node (:690-692); the warp shares the one segment
            int rounds = 1;
            {
                const double2 nb = XY[bs];
                const int r = seg_blocked<RECT, XORD>(ob.tab, ob.g, ob.M, ob.full_cells, qx, qy, nb.x, nb.y, lane, 32);
                my_tests += r >> 1;
                if (!__any_sync(kFull, r & 1)) near_slot = bs;
            }
            PH_MARK(2);                                                           // segment test of the nearest node
            if (near_slot < 0 && n_live > 1) {
                double lo_d = bd; int lo_s = bs;                                  // keys <= (lo_d, lo_s) are known to be blocked
                double T = bd;
                const double t_floor = dmul(dmul(prm.step_length, prm.step_length), 0.0625);
                rounds = n_live;                                                  // if nothing is visible every node was tried
                bool first_annulus = true;
                for (;;) {
                    // any schedule of thresholds is exact; it trades full scans (one per annulus) against segment tests
                    T = dmul(T, first_annulus ? (double)RRTFND_ANN_FIRST : (double)RRTFND_ANN_GROW);
                    first_annulus = false;
                    if (T < t_floor) T = t_floor;
                    if (NG && list && !covers_all && T > ng_complete_d2(rho)) {    // the window must hold every node within sqrt(T)
                        rho = max(2 * rho, (int)(sqrt(T) * ng.inv_cell) + 1);
                        int cx0, cx1, cy0, cy1;
                        covers_all = ng_window(qx, qy, rho, cx0, cx1, cy0, cy1);
                        const int n = ng_gather(HEAD, NEXT, ng.nx, cx0, cx1, cy0, cy1, CAND, ng.cand_cap, lane);
                        if (n > ng.cand_cap) { list = nullptr; n_items = n_slots; covers_all = true; }
                        else n_items = n;
                    }
                    double vd = CUDART_INF; int vs = kIntMax;                      // this lane's best visible key
                    bool beyond = NG && list && !covers_all;
                    int count = 0;
                    const int n_tiles = (n_items + kTileNodes - 1) / kTileNodes;
                    if (RING && !list) stream_begin(xs, XY, n_slots, lane);
                    for (int t = 0; t < n_tiles; ++t) {
                        double2 v[4]; int sl[4];
                        fetch_tile<NG, RING>(xs, XY, list, n_items, t, lane, v, sl);
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
                        while (count - head >= 32 || (last && count > head)) {     // one lane per staged segment
                            const int n = min(count - head, 32);
                            __syncwarp();
                            if (lane < n) {
                                const int s = stage[head + lane];
                                const double2 nb = XY[s];
                                const int r = seg_blocked<RECT, XORD>(ob.tab, ob.g, ob.M, ob.full_cells, qx, qy, nb.x, nb.y, 0, 1);
                                my_tests += r >> 1;
                                if (!(r & 1)) {
                                    const double d = dist2(nb.x, nb.y, qx, qy);
                                    if (d < vd || (d == vd && s < vs)) { vd = d; vs = s; }
                                }
                            }
                            __syncwarp();
                            head += n;
                        }
                        if (head > 0) {                                           // keep the (< 32) leftover at the front
                            const int rem = count - head;
                            const int t2 = lane < rem ? stage[head + lane] : 0;
                            __syncwarp();
                            if (lane < rem) stage[lane] = t2;
                            count = rem;
                            __syncwarp();
                        }
                    }
                    n_scan_nodes += list ? n_items : n_live;
                    warp_argmin(vd, vs);
                    if (vs != kIntMax) {
                        near_slot = vs;
                        rounds = (NG && list) ? ng_key_rank(XY, list, n_items, qx, qy, vd, vs, lane)
                                              : key_rank(XY, n_slots, qx, qy, vd, vs, lane);
                        break;
                    }
                    if (!__any_sync(kFull, beyond)) break;                        // every node tried, none visible
                    lo_d = T; lo_s = kIntMax;
                }
            }
            n_edge_checks += rounds; n_edge_checks_ref += rounds;
            PH_MARK(3);                                                           // annulus search
            if (near_slot < 0) continue;                                          // :45-46
        }

        // ================================================================ D. steer + insert (every lane computes, lane 0 stores)
        const double2 nb = XY[near_slot];
        double px = qx, py = qy;
        {
            const double d = dist_fma(qx, qy, nb.x, nb.y);                        // steer (:739)
            const double ux = ddiv(dsub(qx, nb.x), d), uy = ddiv(dsub(qy, nb.y), d);
            const double tx = dadd(nb.x, dmul(ux, prm.step_length)), ty = dadd(nb.y, dmul(uy, prm.step_length));
            if (d > prm.step_length) { px = tx; py = ty; }
        }
        const int new_slot = n_slots++;
        const int id_new = ++last_id;                                            // Graph.add_vertex (src/graph.py:57-61)
        ++n_live; ++fn_count;                                                    // n_of_nodes += 1 (:221)
        const double e0 = dist_fma(px, py, nb.x, nb.y);                          // :49
        const bool reached = dist_fma(px, py, gx, gy) < goal_thresh;             // check_solution (:757-758)
        if (lane == 0) {
            XY[new_slot] = make_double2(px, py);
            ID[new_slot] = id_new; PARENT[new_slot] = near_slot; ECOST[new_slot] = e0;   // :141-142
            NCHILD[new_slot] = 0; NFLAG[new_slot] = 0;
            if (WREC) wrec_store(WR, new_slot, e0, near_slot, PARENT[near_slot]);
            if (NG) {                                                             // push onto its cell's list
                const int c = grid_cell(py, ng.y0, ng.inv_cell, ng.ny) * ng.nx + grid_cell(px, ng.x0, ng.inv_cell, ng.nx);
                NEXT[new_slot] = HEAD[c]; HEAD[c] = new_slot;
            }
        }
        __syncwarp();

        if (mode == RRTFND_MODE_RRT) {
            int L = 0; double cost = 0.0;
            if (lane == 0) {
                NCHILD[near_slot] += 1;                                           // add_edge (:76)
                if (reached) {                                                    // :78-81, break before iter += 1
                    int n = new_slot;
                    while (n != root_slot && n >= 0) {
                        if (L < prm.path_cap) BPATH[L] = ID[n];
                        ++L; cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++my_hops;
                    }
                    if (L < prm.path_cap) BPATH[L] = ID[root_slot];
                    ++L;
                }
                if (TRACE && iter < prm.trace_cap) {
                    rrtfnd_trace_t t; t.id_new = id_new; t.id_near = ID[near_slot]; t.n_near = 0; t.n_rewired = 0;
                    t.removed_id = -1; t.cp_parent = -1; t.cp_cost = 0.0;
                    TRACE[iter] = t;
                }
            }
            __syncwarp();
            if (reached) {
                best_len = __shfl_sync(kFull, L, 0); best_cost = __shfl_sync(kFull, cost, 0);
                best_leaf_slot = new_slot; solution_found = 1;
                status = best_len > prm.path_cap ? RRTFND_ST_PATH_CAP : RRTFND_ST_DONE;
                break;
            }
            ++iter;
            continue;
        }

        PH_MARK(4);                                                               // steer + insert
        // ================================================================ E. near set -> collision, costs, choose-parent, rewire
        // The ball (query_ball_point, :837 / :893) is produced in ascending slot order by the scan and handed to the
        // lanes in chunks of up to 32 candidates: lane i owns candidate i (its own segment test, its own parent-chain
        // walk). In the first chunk lane 31 walks the new node's chain (Graph.get_cost(id_new)). Rewire decisions are
        // taken against the pre-rewire tree and committed after the scan.
        int n_rw = 0, k_total = 0;
        int cp_slot = near_slot; double cp_cost = e0;
        bool dup = false;
        double cost_new = 0.0;
        // node grid: the ball lies inside the window of cells within `radius` of q_new; its nodes, sorted by slot, replace the scan
        const int32_t* nlist = nullptr;
        int n_near_items = n_slots;
        if (NG && n_live >= ng.min_nodes) {
            int cx0, cx1, cy0, cy1;
            ng_window(px, py, (int)(prm.radius * ng.inv_cell) + 1, cx0, cx1, cy0, cy1);
            const int n = ng_gather(HEAD, NEXT, ng.nx, cx0, cx1, cy0, cy1, CAND, ng.cand_cap, lane);
            if (n <= ng.cand_cap && n <= 512) { ng_sort(CAND, CAND + ng.cand_cap, n, lane); nlist = CAND + ng.cand_cap; n_near_items = n; }
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
            if (RING && !nlist) stream_begin(xs, XY, n_slots, lane);
            for (int t = 0; t < n_tiles; ++t) {
                double2 v[4]; int sl[4];
                fetch_tile<NG, RING>(xs, XY, nlist, n_near_items, t, lane, v, sl);
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
                PH_MARK(5);                                                       // near-set scan + staging
                while (count - head >= (first ? 31 : 32) || (last && count > head)) {
                    const int n = min(count - head, first ? 31 : 32);
                    __syncwarp();
                    // ---- one chunk: lane i < n owns candidate i; in the first chunk lane 31 walks the new node's chain
                    int s = -1; double dpl = 0.0, cc = 0.0; bool hit = true;
                    int w_from = -1;
                    if (lane < n) {
                        s = stage[head + lane];
                        const double2 c = XY[s];
                        dpl = dist_plain(c.x, c.y, px, py);                       // :842 / :899
                        if (dpl == 0.0 && c.x == px && c.y == py) dup = true;     // src/graph.py:54-56 corner (B9)
                        // Line(node, q_new): p1 = node (:846 / :903)
                        const int r = seg_blocked<RECT, XORD>(ob.tab, ob.g, ob.M, ob.full_cells, c.x, c.y, px, py, 0, 1);
                        my_tests += r >> 1;
                        hit = r & 1;
                        if (!hit) w_from = s;                                     // Graph.get_cost(id) (:848 / :905)
                    }
                    // parent-chain walks, one loop for all lanes (Graph.get_cost, src/graph.py:41-46): lane 31 of the first
                    // chunk walks the new node's chain, starts from its edge cost and records the addends
                    const bool rec = first && lane == 31;
                    int depth = 0;
                    PH_MARK(6);                                                   // per-candidate segment tests
                    if (rec) { w_from = walk_from; cc = walk_acc; }
                    if (w_from >= 0) {
                        if (WREC) depth = walk_records(WR, w_from, rec ? chain : nullptr, cc, kChainCap);
                        else {
                            int wn = w_from, q = PARENT[wn];
                            while (q >= 0) {
                                const double e = ECOST[wn];
                                if (rec && depth < kChainCap) chain[depth] = e;
                                cc = dadd(cc, e);
                                wn = q; q = PARENT[wn];
                                ++depth;
                            }
                        }
                        my_hops += depth;
                    }
                    __syncwarp();
                    PH_MARK(7);                                                   // parent-chain walks
                    if (first) { cost_new = __shfl_sync(kFull, cc, 31); cp_cn = cost_new; chain_depth = __shfl_sync(kFull, depth, 31); }
                    if (do_cp) {
                        // choose_parent_kdtree's RETURNED edge (:845-852), candidates in ascending id
                        unsigned open = __ballot_sync(kFull, lane < n && !hit);
#if RRTFND_CP_PARALLEL
                        // An accepted candidate sets G.cost[id_new] = its distance with the parent untouched, so get_cost(id_new)
                        // becomes that distance summed along id_near's chain (:849) - a value that depends on the candidate
                        // alone, not on the acceptances before it: every lane prepares its own, then the sequential pass over
                        // the candidates is a compare and a select per candidate.
                        double resum = dpl;
                        if (open && chain_depth <= kChainCap)
                            for (int h = 0; h < chain_depth; ++h) resum = dadd(resum, chain[h]);
                        const double via = dadd(cc, dpl);
                        while (open) {
                            const int i = __ffs(open) - 1;
                            open &= open - 1;
                            if (cp_cn > __shfl_sync(kFull, via, i)) {
                                cp_slot = __shfl_sync(kFull, s, i); cp_cost = __shfl_sync(kFull, dpl, i);
                                if (chain_depth <= kChainCap) cp_cn = __shfl_sync(kFull, resum, i);
                                else {
                                    long long h = 0;
                                    cp_cn = chain_walk(PARENT, ECOST, near_slot, cp_cost, h);   // every lane walks the same chain
                                    if (lane == 0) my_hops += h;
                                }
                            }
                        }
#else
                        while (open) {
                            const int i = __ffs(open) - 1;
                            open &= open - 1;
                            const double ci = __shfl_sync(kFull, cc, i), di = __shfl_sync(kFull, dpl, i);
                            if (cp_cn > dadd(ci, di)) {
                                cp_slot = __shfl_sync(kFull, s, i); cp_cost = di;
                                // G.cost[id_new] = cost, parent untouched: the chain is still id_near's (:849)
                                if (chain_depth <= kChainCap) {                       // re-sum the cached addends
                                    cp_cn = di;
                                    for (int h = 0; h < chain_depth; ++h) cp_cn = dadd(cp_cn, chain[h]);
                                } else {
                                    long long h = 0;
                                    cp_cn = chain_walk(PARENT, ECOST, near_slot, di, h);   // every lane walks the same chain
                                    if (lane == 0) my_hops += h;
                                }
                            }
                        }
#endif
                    }
                    if (do_rw) {
                        const bool rw = lane < n && !hit && cc > dadd(cost_new, dpl);   // :905, strict
                        const unsigned m = __ballot_sync(kFull, rw);
                        if (m) {   // remember the decision: short list in shared memory, overflow as a per-node flag bit
                            const int pos = n_rw + __popc(m & lt_mask);
                            if (rw) { if (pos < prm.near_cap) rwlist[pos] = s; else NFLAG[s] |= 2; }
                            n_rw += __popc(m);
                        }
                    }
                    first = false;
                    __syncwarp();
                    PH_MARK(8);                                                   // choose-parent evaluation + rewire decisions
                    head += n;
                }
                if (head > 0) {                                                   // keep the (< 32) leftover at the front
                    const int rem = count - head;
                    const int t = lane < rem ? stage[head + lane] : 0;
                    __syncwarp();
                    if (lane < rem) stage[lane] = t;
                    count = rem;
                    __syncwarp();
                }
            }

            if (pass == 0 && commit_cp && k_total > 0) {   // EXTENSION (not the shipped behaviour): adopt the returned edge
                if (lane == 0) {
                    PARENT[new_slot] = cp_slot; ECOST[new_slot] = cp_cost;
                    if (WREC) wrec_store(WR, new_slot, cp_cost, cp_slot, PARENT[cp_slot]);
                }
                __syncwarp();
            }
        }
        if (__any_sync(kFull, dup)) { status = RRTFND_ST_DUPLICATE; break; }
        n_scan_nodes += nlist ? n_near_items : n_live;
        n_edge_checks += k_total;
        n_edge_checks_ref += 2 * (long long)k_total + 2;   // each candidate twice + Line(q_new, q_new) twice

        // ---- E4. commit: rewire_kdtree's re-parenting (:906-910) and G.add_edge(*best_edge) (:148), serial on lane 0
        bool need_refresh = false;
        {
            int d_childless = 0, flagged = 0;
            if (lane == 0) {
                const int new_parent = PARENT[new_slot];
                auto reparent = [&](int s) {
                    const int oldp = PARENT[s];
                    if (--NCHILD[oldp] == 0) { ++d_childless; if (mode == RRTFND_MODE_FN) cb[oldp >> 5] |= 1u << (oldp & 31); }
                    const double2 v = XY[s];
                    PARENT[s] = new_slot;
                    const double ec = dist_plain(v.x, v.y, px, py);
                    ECOST[s] = ec;
                    if (WREC) wrec_store(WR, s, ec, new_slot, new_parent);        // its children keep a stale hint until a walk repairs it
                    flagged |= NFLAG[s] & 1;
                };
                const int n_list = n_rw < prm.near_cap ? n_rw : prm.near_cap;
                for (int i = 0; i < n_list; ++i) reparent(rwlist[i]);
                if (n_rw > n_list) {          // more re-parented nodes than the list holds (rare): find the flagged rest
                    for (int s = 0; s < new_slot; ++s)
                        if (NFLAG[s] & 2) { NFLAG[s] &= (uint8_t)~2; reparent(s); }
                }
                NCHILD[new_slot] = n_rw;
                if (n_rw == 0) { ++d_childless; if (mode == RRTFND_MODE_FN) cb[new_slot >> 5] |= 1u << (new_slot & 31); }
                const int par_new = PARENT[new_slot];
                if (NCHILD[par_new]++ == 0) { --d_childless; if (mode == RRTFND_MODE_FN) cb[par_new >> 5] &= ~(1u << (par_new & 31)); }
                if (flagged) {   // the moved subtree holds a goal-reaching node: mark its new ancestors
                    int n = new_slot;
                    while (n >= 0 && !(NFLAG[n] & 1)) { NFLAG[n] |= 1; n = PARENT[n]; }
                }
            }
            __syncwarp();
            n_childless += __shfl_sync(kFull, d_childless, 0);
            need_refresh = __shfl_sync(kFull, flagged, 0) != 0;
            n_rewired += n_rw;
        }

        PH_MARK(10);                                                              // rewire commit
        const int cp_parent_id = eval_cp ? ID[cp_slot] : -1;   // read before forced_removal can delete that node

        // ================================================================ F. forced_removal (:233-237, :800-819)
        int removed_id = -1;
        if (mode == RRTFND_MODE_FN && fn_count > prm.max_nodes) {                 // n_of_nodes > max_nodes (:233)
            const int leaf = best_leaf_slot;
            const int total = n_childless;
            int n_ok = total - (NCHILD[new_slot] == 0 ? 1 : 0);
            if (leaf >= 0 && leaf != new_slot && NCHILD[leaf] == 0) --n_ok;
            if (n_ok <= 0) { status = RRTFND_ST_NO_CHILDLESS; break; }
            int pick = -1;
            for (;;) {
                const uint32_t r = next_randbelow(rng, (uint32_t)total);          // random.choice (:813 / :816)
                if (rng.exhausted) { status = RRTFND_ST_STREAM_EXHAUSTED; break; }
                // r-th childless node in ascending slot order: every lane counts a contiguous run of bitmap words, a warp
                // prefix sum names the lane whose run holds it, that lane picks the bit
                const int nw = (n_slots + 31) >> 5, wpl = (nw + 31) >> 5;
                int cnt = 0;
                for (int i = 0; i < wpl; ++i) { const int w = lane * wpl + i; if (w < nw) cnt += __popc(cb[w]); }
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(kFull, incl, off); if (lane >= off) incl += t; }
                const int excl = incl - cnt;
                const unsigned owner = __ballot_sync(kFull, (int)r >= excl && (int)r < incl);
                int found = -1;
                if (owner >> lane & 1u) {
                    int rem = (int)r - excl;
                    for (int i = 0; i < wpl; ++i) {
                        const int w = lane * wpl + i;
                        const unsigned bits = w < nw ? cb[w] : 0u;
                        const int c = __popc(bits);
                        if (rem < c) { found = w * 32 + (int)__fns(bits, 0, rem + 1); break; }
                        rem -= c;
                    }
                }
                pick = owner ? __shfl_sync(kFull, found, __ffs(owner) - 1) : -1;
                if (pick < 0) { status = RRTFND_ST_NO_CHILDLESS; break; }          // the count and the bitmap disagree (cannot happen)
                if (pick != new_slot && pick != leaf) break;                      // :815
            }
            if (status != RRTFND_ST_RUNNING) break;
            // finish_nodes_of_path.remove(id_removed) (:235-236), order preserving
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
                --n_finish;
            }
            int d_childless = -1;                                                 // the removed node was childless
            removed_id = ID[pick];
            __syncwarp();
            if (lane == 0) {                                                      // Graph.remove_vertex (src/graph.py:64-77)
                const int par = PARENT[pick];
                if (par >= 0 && --NCHILD[par] == 0) { ++d_childless; cb[par >> 5] |= 1u << (par & 31); }
                cb[pick >> 5] &= ~(1u << (pick & 31));
                ID[pick] = -1; NCHILD[pick] = -1; PARENT[pick] = -1;
                XY[pick] = nan2;
            }
            __syncwarp();
            n_childless += __shfl_sync(kFull, d_childless, 0);
            --n_live; --fn_count; ++n_removed;
        }

        PH_MARK(11);                                                              // forced removal
        // ================================================================ G. goal test and best path (:155-170, :240-251)
        if (reached) {
            if (n_finish >= prm.finish_cap) { status = RRTFND_ST_FINISH_CAP; break; }
            int L = 0; double cost = 0.0;
            if (lane == 0) {
                FINISH[n_finish] = new_slot;
                int n = new_slot;                                                 // mark the ancestors of the new finish node
                while (n >= 0 && !(NFLAG[n] & 1)) { NFLAG[n] |= 1; n = PARENT[n]; }
                if (mode == RRTFND_MODE_STAR) {                               // unconditional overwrite (:160-161)
                    n = new_slot;
                    while (n != root_slot && n >= 0) {
                        if (L < prm.path_cap) BPATH[L] = ID[n];
                        ++L; cost = dadd(cost, ECOST[n]); n = PARENT[n]; ++my_hops;
                    }
                    if (L < prm.path_cap) BPATH[L] = ID[root_slot];
                    ++L;
                }
            }
            __syncwarp();
            ++n_finish;
            solution_found = 1;
            if (mode == RRTFND_MODE_STAR) {
                best_len = __shfl_sync(kFull, L, 0); best_cost = __shfl_sync(kFull, cost, 0); best_leaf_slot = new_slot;
                if (best_len > prm.path_cap) { status = RRTFND_ST_PATH_CAP; break; }
            }
        }
        // The reference recomputes every finish node's path each iteration and keeps a strictly cheaper one
        // (:166-170, :247-251). A path cost only changes when a node with a goal-reaching descendant is rewired, and
        // the list only grows through a new finish node, so those are the only iterations in which it can have an effect.
        if ((reached || need_refresh) && n_finish > 0) {
            double bd = CUDART_INF; int bf = kIntMax;
            for (int f = lane; f < n_finish; f += 32) {
                const double cost = path_cost(PARENT, ECOST, FINISH[f], root_slot, my_hops);
                if (cost < bd) { bd = cost; bf = f; }
            }
            __syncwarp();
            warp_argmin(bd, bf);
            if (bd < best_cost) {                                                 // strict (:168, :249)
                const int leaf = FINISH[bf];
                int L = 0;
                if (lane == 0) {
                    int n = leaf;
                    while (n != root_slot && n >= 0) {
                        if (L < prm.path_cap) BPATH[L] = ID[n];
                        ++L; n = PARENT[n]; ++my_hops;
                    }
                    if (L < prm.path_cap) BPATH[L] = ID[root_slot];
                    ++L;
                }
                __syncwarp();
                best_len = __shfl_sync(kFull, L, 0); best_cost = bd; best_leaf_slot = leaf;
                if (best_len > prm.path_cap) { status = RRTFND_ST_PATH_CAP; break; }
            }
        }

        // ================================================================ H. end of iteration
        if (TRACE && iter < prm.trace_cap && lane == 0) {
            rrtfnd_trace_t t;
            t.id_new = id_new; t.id_near = ID[near_slot]; t.n_near = k_total; t.n_rewired = n_rw;
            t.removed_id = removed_id;
            t.cp_parent = cp_parent_id; t.cp_cost = eval_cp ? cp_cost : 0.0;
            TRACE[iter] = t;
        }
        ++iter;                                                                   // :173 / :254
    }

    // ---- write back
    __syncwarp();
#ifdef RRTFND_PHASE_CLOCKS
    if (lane == 0) for (int i = 0; i < 12; ++i) atomicAdd(&g_phase_clocks[i], (unsigned long long)ph_acc[i]);
#endif
    const long long hops = warp_sum(my_hops), tests = warp_sum(my_tests);
    if (lane == 0) {
        PL->best_cost = best_cost; PL->cursor = rng.cursor; PL->attempts = attempts;
        PL->n_edge_checks = n_edge_checks; PL->n_edge_checks_ref = n_edge_checks_ref;
        PL->n_circle_tests += tests; PL->n_scan_nodes = n_scan_nodes; PL->n_chain_hops += hops;
        PL->n_slots = n_slots; PL->n_live = n_live; PL->last_id = last_id; PL->iter = iter;
        PL->root_slot = root_slot; PL->n_finish = n_finish; PL->best_len = best_len; PL->status = status;
        PL->n_rewired = n_rewired; PL->n_removed = n_removed; PL->solution_found = solution_found;
        PL->n_compactions = n_compactions; PL->fn_count = fn_count;
        if (n_slices > 1) {                               // this planner's state is complete: its next slice may start
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(&PL->reserved), "r"(bt.reserved + slice + 1) : "memory");
        }
    }
    }
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

template <int WPB, int MINW, bool NG = false, int MODE = -1, bool RING = true, bool RECT = (MODE < 0), bool XORD = true>
static int launch_plan(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, cudaStream_t st) {
    size_t bytes = warp_smem_bytes(*prm, prm->reserved & 15) * WPB;
    auto kern = plan_kernel<WPB, MINW, NG, MODE, RING, NG, RECT, XORD>;
    int blocks = (bt->n_planners + WPB - 1) / WPB;
    if (bytes > kMaxSmem) return RRTFND_E_SMEM;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    if (const char* cv = getenv("RRTFND_CARVEOUT")) {   // experiment: shared-memory carve-out in percent (the rest of the 256 KB is L1)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
        if (e != cudaSuccess) return (int)e;
    }
    const int n_slices = (prm->reserved >> 4) & 15;
    if (n_slices <= 1) {
        kern<<<blocks, 32 * WPB, bytes, st>>>(*prm, *bt);
        return (int)cudaGetLastError();
    }
    // Sliced run. Slice k grows every planner to iter_num * f(k / n_slices) (f below). Every
    // launch but the first carries the programmatic-stream-serialization attribute: it may start once all CTAs of its
    // predecessor have started, and its warps wait for their own planner's ticket (see the kernel). Tickets are serial * 16 +
    // slice, strictly increasing over the calls of this process, so a stale ticket in a re-used record is never ahead.
    static std::atomic<int> g_serial{0};
    rrtfnd_batch_t b = *bt;
    b.reserved = (g_serial.fetch_add(1) + 1) * 16;
    for (int k = 0; k < n_slices; ++k) {
        rrtfnd_params_t q = *prm;
        // slice k ends at iter_num * f(x), x = (k + 1) / n_slices. RRT*: f = 1 - (1 - x)^3 - long first slices, ever shorter
        // last ones, because what is exposed at the end of the call is half of the LAST slice (measured on c3: 35.4 M ext/s
        // with f = sqrt(x), 34.7 M linear, 35.8 M with 1 - (1 - x)^2, 36.0 M with the cube; profiles/r02_s7_raw.txt).
        // RRT*-FN: linear (the cap makes every extension cost the same; the shrinking schedule measured -7.5 % on c2).
        const double x = (double)(k + 1) / n_slices;
        const double f = MODE == RRTFND_MODE_STAR ? 1.0 - (1.0 - x) * (1.0 - x) * (1.0 - x) : x;
        q.iter_num = k + 1 == n_slices ? prm->iter_num : (int)(prm->iter_num * f);
        q.reserved = (prm->reserved & 255) | (k << 8);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(32 * WPB); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = k ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, kern, q, b);
        if (e != cudaSuccess) return (int)e;
    }
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
    // one planner per CTA: a finished planner frees its slot at once (an SM holds 32 CTAs, more than the 16 warps the
    // register allocation admits), measured ~2 % faster than two per CTA on a 16,384-planner batch
    if (c.wpb != 1 && c.wpb != 2 && c.wpb != 4) c.wpb = 1;
    // measured on B200 (profiles/): trading registers for residency (80 / 64 registers) spills and halves throughput,
    // and a ring deeper than 2 tiles does not pay once a dozen warps share an SM
    if (c.minw != 16 && c.minw != 24 && c.minw != 32) c.minw = 16;
    if (c.stages != 2 && c.stages != 4) c.stages = 2;
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
    if ((prm->flags & RRTFND_FLAG_NODE_GRID) && !bt->walk) return RRTFND_E_BADARG;     // the node-grid kernel walks the records
    if (bt->grid.nx > 0 && (!bt->grid.seg_start || !bt->grid.seg_items || !bt->grid.pt_start || !bt->grid.pt_items ||
                            bt->grid.ny <= 0)) return RRTFND_E_BADARG;
    return 0;
}

#ifdef RRTFND_PHASE_CLOCKS
// diagnostic builds only: copy (and optionally clear) the per-phase clock sums
extern "C" int rrtfnd_debug_phase_clocks(unsigned long long* out_host, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out_host, g_phase_clocks, sizeof(unsigned long long) * 16);
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_phase_clocks, z, sizeof z);
    }
    return (int)e;
}
#endif

static bool is_lean(const rrtfnd_params_t& q, const PlanCfg& cfg) {
    return (q.flags & (RRTFND_FLAG_TRACE | RRTFND_FLAG_COMMIT_CHOOSE_PARENT | RRTFND_FLAG_NODE_GRID)) == 0 &&
           (q.flags & RRTFND_FLAG_EVAL_CHOOSE_PARENT) && q.mode != RRTFND_MODE_RRT && cfg.wpb == 1 && cfg.minw == 16;
}
// Batches larger than the resident warp slots are grown in slices whose launches overlap (launch_plan). A one-shot launch runs
// planners to completion in index order: its last wave is partly empty and its end waits for the slowest planner while the
// other slots idle. Sliced, every planner advances through the same short stages and only the last eighth of a run is exposed.
// Measured on B200 (profiles/r02_s3_slices_raw.txt): c3 16,384 planners +7.0 % (8 slices; 4: +6.3 %, 15: +4.3 %), c3 4,096
// planners +15 %, c2 4,096 planners +11.8 %. RRTFND_SLICES=n overrides (1 = off).
static int plan_slices(const rrtfnd_params_t& q, int n_planners) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int n = n_planners > sms * RRTFND_LEAN_MINW ? 8 : 1;
    if (const char* sv = getenv("RRTFND_SLICES")) n = atoi(sv);
    return n < 1 ? 1 : n > 15 ? 15 : n;
}

extern "C" int32_t rrtfnd_plan_launches(const rrtfnd_params_t* prm, int32_t n_planners) {
    if (!prm || n_planners <= 0) return RRTFND_E_BADARG;
    return is_lean(*prm, choose_cfg(*prm, n_planners)) ? plan_slices(*prm, n_planners) : 1;
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
    q.reserved = cfg.stages;                    // the kernel reads its ring depth here (+ 16 * slices, set below for big batches)
    if (q.flags & RRTFND_FLAG_NODE_GRID) {      // large single trees: node-grid kernel, one planner per CTA
        const rrtfnd_node_grid_t& g = b.node_grid;
        if (q.mode == RRTFND_MODE_FN || g.nx <= 0 || g.ny <= 0 || !g.head || !g.next || !g.cand || g.cand_cap < 32 || !(g.inv_cell > 0.0))
            return RRTFND_E_BADARG;
        return launch_plan<1, 16, true>(&q, &b, (cudaStream_t)stream);
    }
    // the configuration batches run in gets a lean kernel instance (one planner per CTA, 128 registers)
    const bool lean = is_lean(q, cfg);
    // Scans of large trees stream from L2 / HBM and are faster through the TMA ring (c3, 5,002 slots: +6 %); trees of a
    // thousand nodes stay L1 / L2 resident and are faster with plain loads (c2, 1,508 slots: +7 %). The `stages` field
    // of the launch word overrides: 1 = plain loads, 2 / 4 = ring.
    const int stages_req = (prm->reserved >> 4) & 15;
    const bool ring = stages_req == 1 ? false : (stages_req == 2 || stages_req == 4) ? true : prm->capacity > 2048;
    if (lean) q.reserved = cfg.stages | (plan_slices(q, b.n_planners) << 4);
    // circle-only tables (no RRTFND_FLAG_RECTANGLES) get the instances without the rectangle dispatch
    const bool rect = (q.flags & RRTFND_FLAG_RECTANGLES) != 0;
#define RRTFND_LEAN(M_, R_, X_, O_) launch_plan<1, RRTFND_LEAN_MINW, false, M_, R_, X_, O_>(&q, &b, st)
    // XORD (cull.cuh: the circles of a grid row are tested from p1's end) pays where a launch waits for its slowest planner -
    // a batch that fits the warp slots (+9.5 % at 2,048 planners) - and on RRT*-FN (+6.6 % on c2); on a multi-wave RRT* batch,
    // bound by aggregate throughput, its extra index arithmetic costs 1.8 %: those launches take the instance without it.
    if (lean && q.mode == RRTFND_MODE_STAR) {
        const bool xord = plan_slices(q, b.n_planners) <= 1;
        if (xord) return rect ? (ring ? RRTFND_LEAN(RRTFND_MODE_STAR, true, true, true) : RRTFND_LEAN(RRTFND_MODE_STAR, false, true, true))
                              : (ring ? RRTFND_LEAN(RRTFND_MODE_STAR, true, false, true) : RRTFND_LEAN(RRTFND_MODE_STAR, false, false, true));
        return rect ? (ring ? RRTFND_LEAN(RRTFND_MODE_STAR, true, true, false) : RRTFND_LEAN(RRTFND_MODE_STAR, false, true, false))
                    : (ring ? RRTFND_LEAN(RRTFND_MODE_STAR, true, false, false) : RRTFND_LEAN(RRTFND_MODE_STAR, false, false, false));
    }
    if (lean && q.mode == RRTFND_MODE_FN)
        return rect ? (ring ? RRTFND_LEAN(RRTFND_MODE_FN, true, true, true) : RRTFND_LEAN(RRTFND_MODE_FN, false, true, true))
                    : (ring ? RRTFND_LEAN(RRTFND_MODE_FN, true, false, true) : RRTFND_LEAN(RRTFND_MODE_FN, false, false, true));
#undef RRTFND_LEAN
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
