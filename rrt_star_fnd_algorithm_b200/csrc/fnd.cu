// fnd.cu — RRT*-FND tree surgery on the device-resident batch.
//
//   rrtfnd_fnd_select_branch   select_branch + remove_children     src/algorithm.py:294-341
//   rrtfnd_fnd_valid_path      valid_path                          src/algorithm.py:344-367
//   rrtfnd_fnd_reconnect       reconnect + get_near_nodes + get_distance_dict + check_for_tree_associativity
//                                                                  src/algorithm.py:370-434, :938-951
//
// One warp per planner; node ids at the interface (the reference's arguments are ids), slots inside. The reference
// deletes subtrees recursively through its children lists; here a node decides for itself by walking its parent chain
// (the tree only stores parents), decisions are written to nflags first and applied afterwards so that every walk sees
// the tree as it was. `path` is the planner's stored best path (ids, goal side first), as in the reference's usage
// example (test_select_branch, :437-479). Obstacles are read un-culled from batch.obst (they have usually just changed).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "cull.cuh"
#include "rng.cuh"

namespace rrtfnd {

constexpr int kFndWarps = 4;
constexpr uint8_t kKeep = 4, kOnPath = 8, kOccupied = 16, kDelete = 32, kSeparate = 64;
constexpr int kReserveWordsFnd = 64;   // as in planner.cu

struct PlannerView {
    double2* XY; double* ECOST; int32_t* PARENT; int32_t* NCHILD; int32_t* ID; uint8_t* NFLAG; int32_t* FINISH; int32_t* BPATH;
    rrtfnd_planner_t* PL;
};

__device__ __forceinline__ PlannerView view_of(const rrtfnd_params_t& prm, const rrtfnd_batch_t& bt, int p) {
    const size_t o = (size_t)p * prm.capacity;
    PlannerView v;
    v.XY = reinterpret_cast<double2*>(bt.xy) + o; v.ECOST = bt.ecost + o; v.PARENT = bt.parent + o; v.NCHILD = bt.nchild + o;
    v.ID = bt.id + o; v.NFLAG = bt.nflags + o; v.FINISH = bt.finish + (size_t)p * prm.finish_cap;
    v.BPATH = bt.best_path + (size_t)p * prm.path_cap; v.PL = bt.planners + p;
    return v;
}

// slot holding node `id`, -1 if it is not alive (ids are unique; warp-cooperative scan)
__device__ int find_slot(const int32_t* ID, int n_slots, int id, int lane) {
    int found = -1;
    for (int base = 0; base < n_slots && found < 0; base += 32) {
        const int s = base + lane;
        const unsigned m = __ballot_sync(kFull, s < n_slots && id >= 0 && ID[s] == id);
        if (m) found = base + __ffs(m) - 1;
    }
    return found;
}

__device__ void tombstone(const PlannerView& v, int s) {
    v.ID[s] = -1; v.NCHILD[s] = -1; v.PARENT[s] = -1;
    v.XY[s] = make_double2(CUDART_NAN, CUDART_NAN);
}

// planner-local finish list: keep the entries that are still alive, in order (lane 0)
__device__ int filter_finish(const PlannerView& v, int n_finish) {
    int k = 0;
    for (int f = 0; f < n_finish; ++f) { const int s = v.FINISH[f]; if (v.ID[s] >= 0) v.FINISH[k++] = s; }
    return k;
}

// find_path(G, leaf, root) (src/algorithm.py:778-797): ids into BPATH, cost summed leaf -> root (lane 0)
__device__ void rewrite_best_path(const rrtfnd_params_t& prm, const PlannerView& v, int leaf, int root, int* len, double* cost) {
    int n = leaf, L = 0; double c = 0.0;
    while (n != root && n >= 0) {
        if (L < prm.path_cap) v.BPATH[L] = v.ID[n];
        ++L; c = dadd(c, v.ECOST[n]); n = v.PARENT[n];
    }
    if (n == root) { if (L < prm.path_cap) v.BPATH[L] = v.ID[root]; ++L; }     // a broken chain leaves the partial path (:794-795)
    *len = L; *cost = c;
}

__global__ void __launch_bounds__(32 * kFndWarps) select_branch_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt,
                                                                       const int32_t* __restrict__ current_id,
                                                                       int32_t* __restrict__ out_status) {
    const int lane = threadIdx.x & 31, p = blockIdx.x * kFndWarps + (threadIdx.x >> 5);
    if (p >= bt.n_planners) return;
    const PlannerView v = view_of(prm, bt, p);
    const int n_slots = v.PL->n_slots, best_len = v.PL->best_len;
    const int cur_id = current_id[p];
    const int cur = find_slot(v.ID, n_slots, cur_id, lane);
    int idx = -1;                                                            // path.index(current_node) (:313)
    for (int base = 0; base < best_len && idx < 0; base += 32) {
        const int i = base + lane;
        const unsigned m = __ballot_sync(kFull, i < best_len && v.BPATH[i] == cur_id);
        if (m) idx = base + __ffs(m) - 1;
    }
    int status = 0;
    if (cur < 0 || idx < 0) status = -1;                                     // ValueError in the reference
    else if (v.PARENT[cur] < 0) status = -2;                                 // G.children[None]: KeyError in the reference
    if (status) { if (lane == 0 && out_status) out_status[p] = status; return; }
    // remove_children(stranded node) for every path node on the root side of `current`, then the stranded nodes themselves
    // (:315-323): a node goes iff its parent chain meets a stranded path node before it meets `current`. Trees that hang
    // from neither (split off by an earlier valid_path and never mended) are not touched by the reference, nor here.
    for (int s = lane; s < n_slots; s += 32) {
        const int id = v.ID[s];
        uint8_t f = v.NFLAG[s] & (uint8_t)~(kKeep | kOnPath);
        if (id >= 0)
            for (int i = idx + 1; i < best_len; ++i) if (v.BPATH[i] == id) { f |= kOnPath; break; }
        v.NFLAG[s] = f;
    }
    __syncwarp();
    for (int s = lane; s < n_slots; s += 32) {
        if (v.ID[s] < 0) continue;
        int n = s;
        while (n >= 0 && n != cur && !(v.NFLAG[n] & kOnPath)) n = v.PARENT[n];
        if (n < 0 || n == cur) v.NFLAG[s] |= kKeep;
    }
    __syncwarp();
    int n_del = 0;
    for (int s = lane; s < n_slots; s += 32)
        if (v.ID[s] >= 0 && !(v.NFLAG[s] & kKeep)) { tombstone(v, s); ++n_del; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_del += __shfl_xor_sync(kFull, n_del, off);
    const int leaf = find_slot(v.ID, n_slots, v.BPATH[0], lane);
    __syncwarp();
    if (lane == 0) {
        v.PARENT[cur] = -1;                                                  // :305
        const double2 c = v.XY[cur];
        v.PL->start_x = c.x; v.PL->start_y = c.y;                            // G.start = G.vertices[current_node] (:303)
        v.PL->root_slot = cur;
        v.PL->n_live -= n_del;
        double cost = 0.0;                                                   // new_path = path[:idx + 1]: the stored ids stay,
        for (int n = leaf; n != cur && n >= 0; n = v.PARENT[n]) cost = dadd(cost, v.ECOST[n]);   // its cost is re-summed to the new root
        v.PL->best_len = idx + 1; v.PL->best_cost = cost;
        v.PL->n_finish = filter_finish(v, v.PL->n_finish);
        if (out_status) out_status[p] = 0;
    }
}

__global__ void __launch_bounds__(32 * kFndWarps) valid_path_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt,
                                                                    const int32_t* __restrict__ previous_root_id,
                                                                    int32_t* __restrict__ out_sep_id,
                                                                    int32_t* __restrict__ out_index_error) {
    extern __shared__ int32_t fnd_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, p = blockIdx.x * kFndWarps + wib;
    if (p >= bt.n_planners) return;
    if (previous_root_id[p] < 0) {                                           // negative id = leave this planner alone
        if (lane == 0) { out_sep_id[p] = -1; if (out_index_error) out_index_error[p] = 0; }
        return;
    }
    int32_t* pslot = fnd_smem + (size_t)wib * prm.path_cap;                  // slot of path node i, -1 if it is gone
    const PlannerView v = view_of(prm, bt, p);
    const int n_slots = v.PL->n_slots, L = min(v.PL->best_len, prm.path_cap), M = prm.n_obst;
    const double2* tab = reinterpret_cast<const double2*>(bt.obst);
    for (int i = lane; i < L; i += 32) pslot[i] = -1;
    __syncwarp();
    for (int s = lane; s < n_slots; s += 32) {
        const int id = v.ID[s];
        uint8_t f = v.NFLAG[s] & (uint8_t)~(kOnPath | kOccupied | kDelete);
        if (id >= 0)
            for (int i = 0; i < L; ++i) if (v.BPATH[i] == id) { pslot[i] = s; f |= kOnPath; break; }
        v.NFLAG[s] = f;
    }
    __syncwarp();
    for (int i = lane; i < L; i += 32) {                                      // map.is_occupied_c(pos_node) (:356)
        const int s = pslot[i];
        if (s < 0) continue;
        const double2 c = v.XY[s];
        bool occ = false;
        for (int j = 0; j < M && !occ; ++j) {
            double ox, oy, r, K;
            load_obst(tab, j, ox, oy, r, K);
            occ = point_in_inflated(c.x, c.y, ox, oy, r, prm.node_radius, K);
        }
        if (occ) v.NFLAG[s] |= kOccupied;
    }
    __syncwarp();
    const int leaf = L > 0 ? pslot[0] : -1;
    // a node goes if the first path node on its parent chain (itself included) is occupied: that is the occupied node
    // and the off-path subtrees hanging from it (remove_children, :357). An occupied goal-side leaf stays: the
    // reference raises IndexError at :358 after removing its children.
    for (int s = lane; s < n_slots; s += 32) {
        if (v.ID[s] < 0) continue;
        int n = s;
        while (n >= 0 && !(v.NFLAG[n] & kOnPath)) n = v.PARENT[n];
        if (n >= 0 && (v.NFLAG[n] & kOccupied) && s != leaf) v.NFLAG[s] |= kDelete;
    }
    __syncwarp();
    int sep = -1, err = 0;
    if (lane == 0) {
        for (int i = L - 1; i >= 1; --i) {                                    // root side first (:355)
            const int s = pslot[i];
            if (s < 0 || !(v.NFLAG[s] & kOccupied)) continue;
            const int pv = v.PARENT[s];
            if (pv >= 0 && !(v.NFLAG[pv] & kDelete)) v.NCHILD[pv] -= 1;       // Graph.remove_vertex (:359)
            const int c = pslot[i - 1];                                       // the remaining child is the one on the path (:358)
            if (c >= 0) { v.PARENT[c] = -1; sep = c; }                        // :360-361
        }
        if (leaf >= 0 && (v.NFLAG[leaf] & kOccupied)) { err = 1; v.NCHILD[leaf] = 0; }   // its children were removed before the raise
    }
    sep = __shfl_sync(kFull, sep, 0); err = __shfl_sync(kFull, err, 0);
    __syncwarp();
    int n_del = 0;
    for (int s = lane; s < n_slots; s += 32)
        if (v.ID[s] >= 0 && (v.NFLAG[s] & kDelete)) { tombstone(v, s); ++n_del; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_del += __shfl_xor_sync(kFull, n_del, off);
    __syncwarp();
    if (lane == 0) {
        v.PL->n_live -= n_del;
        v.PL->n_finish = filter_finish(v, v.PL->n_finish);
        if (out_index_error) out_index_error[p] = err;
        out_sep_id[p] = err ? -1 : (sep >= 0 ? v.ID[sep] : previous_root_id[p]);
    }
}

__global__ void __launch_bounds__(32 * kFndWarps) reconnect_kernel(const rrtfnd_params_t prm, const __grid_constant__ rrtfnd_batch_t bt,
                                                                   const int32_t* __restrict__ separate_root_id, double radius,
                                                                   int32_t* __restrict__ out_linked_id) {
    const int lane = threadIdx.x & 31, p = blockIdx.x * kFndWarps + (threadIdx.x >> 5);
    if (p >= bt.n_planners) return;
    const PlannerView v = view_of(prm, bt, p);
    const int n_slots = v.PL->n_slots, root = v.PL->root_slot;
    const int sep = find_slot(v.ID, n_slots, separate_root_id[p], lane);
    int link = -1;
    if (sep >= 0) {
        const double2 q = v.XY[sep];
        const double y2 = dadd(dmul(q.x, q.x), dmul(q.y, q.y));
        const double2* tab = reinterpret_cast<const double2*>(bt.obst);
        for (int base = 0; base < n_slots && link < 0; base += 32) {         // dict order = ascending id: first hit wins (:386-392)
            const int s = base + lane;
            bool ok = false;
            if (s < n_slots && v.ID[s] >= 0) {
                const double2 c = v.XY[s];
                // get_distance_dict (:938-951): sqrt(|p|^2 - 2 p.q + |q|^2), the BLAS dot being fma(px, qx, fl(py*qy))
                const double x2 = dadd(dmul(c.x, c.x), dmul(c.y, c.y));
                const double dot = __fma_rn(c.x, q.x, dmul(c.y, q.y));
                const double d = dsqrt(dadd(dsub(x2, dmul(2.0, dot)), y2));
                if (d <= radius) {                                            // NaN (cancellation) is not <= radius (:432)
                    int n = s;
                    while (v.PARENT[n] >= 0) n = v.PARENT[n];                 // check_for_tree_associativity (:407-420)
                    if (n == root)                                            // Line(node, separate root) (:387-388)
                        ok = !(seg_blocked(tab, &bt.grid, prm.n_obst, 0, c.x, c.y, q.x, q.y, 0, 1) & 1);
                }
            }
            const unsigned m = __ballot_sync(kFull, ok);
            if (m) link = base + __ffs(m) - 1;
        }
    }
    const int leaf = link >= 0 ? find_slot(v.ID, n_slots, v.BPATH[0], lane) : -1;
    __syncwarp();
    if (lane == 0) {
        if (link >= 0) {
            const double2 q = v.XY[sep], c = v.XY[link];
            v.PARENT[sep] = link; v.NCHILD[link] += 1;                        // G.add_edge(id_separate_root, node, cost) (:389-390)
            v.ECOST[sep] = dist_fma(c.x, c.y, q.x, q.y);
            if (v.NFLAG[sep] & 1)                                             // keep "has a goal-reaching descendant" closed upwards
                for (int n = link; n >= 0 && !(v.NFLAG[n] & 1); n = v.PARENT[n]) v.NFLAG[n] |= 1;
            int L = 0; double cost = 0.0;                                     // find_path(G, last_node, id_root) (:400)
            rewrite_best_path(prm, v, leaf, root, &L, &cost);
            v.PL->best_len = L; v.PL->best_cost = cost;
        }
        out_linked_id[p] = link >= 0 ? v.ID[link] : -1;
    }
}


// regrow (src/algorithm.py:482-528): extend the root tree with plain extensions — brute-force nearest_node (:702-728)
// that skips the separate tree and tests Line(node, sample); choose_parent's return value is dropped at :505 — until a
// new node sees a separate-tree node closer than step_length; that node becomes the root of the separate tree
// (reconnect_ancestors, :531-543: the parent pointers on its chain are reversed, edge costs untouched) and is hung under
// the new node. At most max_iter extensions (500 in the reference).
__global__ void __launch_bounds__(32 * kFndWarps) regrow_kernel(const __grid_constant__ rrtfnd_params_t prm,
                                                                const __grid_constant__ rrtfnd_batch_t bt,
                                                                const int32_t* __restrict__ separate_root_id, int max_iter,
                                                                int32_t* __restrict__ out_reconnected, int32_t* __restrict__ out_iters) {
    __shared__ WordSrc rngs[kFndWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, p = blockIdx.x * kFndWarps + wib;
    if (p >= bt.n_planners) return;
    const PlannerView v = view_of(prm, bt, p);
    WordSrc& rng = rngs[wib];
    int n_slots = v.PL->n_slots, n_live = v.PL->n_live, last_id = v.PL->last_id;
    long long attempts = v.PL->attempts;
    const int root = v.PL->root_slot;
    const double gx = v.PL->goal_x, gy = v.PL->goal_y;
    const int sep = find_slot(v.ID, n_slots, separate_root_id[p], lane);
    int iter = 0, result = 0;
    if (sep < 0) { if (lane == 0) { out_reconnected[p] = RRTFND_E_BADARG; if (out_iters) out_iters[p] = 0; } return; }
    for (int s = lane; s < n_slots; s += 32) {                                // separate_tree (:483)
        if (v.ID[s] < 0) continue;
        int n = s;
        while (v.PARENT[n] >= 0) n = v.PARENT[n];
        v.NFLAG[s] = (uint8_t)((v.NFLAG[s] & ~kSeparate) | (n == sep ? kSeparate : 0));
    }
    rng.mode = prm.stream_mode;
    rng.words = bt.words ? bt.words + (size_t)p * prm.words_per_planner : nullptr;
    rng.n_words = prm.words_per_planner;
    rng.seed = v.PL->seed; rng.stream = v.PL->stream_id; rng.cursor = v.PL->cursor;
    rng.have_blk = false; rng.exhausted = false; rng.blk_idx = 0; rng.b0 = rng.b1 = rng.b2 = rng.b3 = 0;
    __syncwarp();
    const Obst ob = make_obst(bt.obst, prm.n_obst, &bt.grid);
    long long tests = 0;
    bool reconnected = false;
    const long long attempt_cap = attempts + (long long)max_iter * RRTFND_REGROW_ATTEMPTS_PER_ITER;
    while (iter < max_iter && !reconnected) {
        if (attempts >= attempt_cap) { result = RRTFND_ST_ATTEMPT_CAP; break; }   // the reference would spin forever here
        if (prm.stream_mode == RRTFND_STREAM_WORDS && (long long)rng.cursor + kReserveWordsFnd > prm.words_per_planner) { result = RRTFND_ST_REGROW_NEED_WORDS; break; }
        if (n_slots >= prm.capacity) { result = RRTFND_ST_CAPACITY; break; }
        ++attempts;
        double qx, qy;                                                        // Graph.random_node (src/graph.py:96-98)
        if (next_random(rng) < prm.bias) { qx = gx; qy = gy; }
        else {
            qx = dadd(prm.x_lo, dmul(dsub(prm.x_hi, prm.x_lo), next_random(rng)));
            qy = dadd(prm.y_lo, dmul(dsub(prm.y_hi, prm.y_lo), next_random(rng)));
        }
        if (point_occupied_warp(ob, qx, qy, prm.node_radius, lane, tests)) continue;          // :493
        // nearest_node (:702-728): an identical vertex is returned as is and rejected at :495; otherwise the closest
        // (calc_distance, lowest id on ties) among the nodes outside the separate tree whose Line(node, sample) is free.
        // The reference tests the visibility of EVERY node; the same (distance, id) minimum is found here by testing nodes
        // in ascending key order, 32 at a time: each lane proposes the nearest untested node of its own stride that still
        // beats the best visible node found so far, all lanes test their proposals at once, repeat until none is left.
        bool exact = false;
        for (int s = lane; s < n_slots; s += 32) {
            if (v.ID[s] < 0) continue;
            const double2 c = v.XY[s];
            exact |= (c.x == qx && c.y == qy);
        }
        if (__any_sync(kFull, exact)) continue;
        double bd = CUDART_INF; int bs = 0x7fffffff;                          // best visible so far (warp-uniform)
        double ld = -1.0; int ls = -1;                                        // this lane's last tested key
        for (;;) {
            double cd = CUDART_INF; int cs = 0x7fffffff;
            for (int s = lane; s < n_slots; s += 32) {
                if (v.ID[s] < 0 || (v.NFLAG[s] & kSeparate)) continue;
                const double2 c = v.XY[s];
                const double d = dist_fma(c.x, c.y, qx, qy);
                if (key_less(ld, ls, d, s) && key_less(d, s, cd, cs)) { cd = d; cs = s; }
            }
            const bool have = cs != 0x7fffffff && key_less(cd, cs, bd, bs);
            if (!__any_sync(kFull, have)) break;
            double pd = CUDART_INF; int ps = 0x7fffffff;
            if (have) {
                const double2 c = v.XY[cs];
                ld = cd; ls = cs;
                if (!(seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, c.x, c.y, qx, qy, 0, 1) & 1)) { pd = cd; ps = cs; }
            }
            __syncwarp();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(kFull, pd, off); const int os = __shfl_xor_sync(kFull, ps, off);
                if (key_less(od, os, pd, ps)) { pd = od; ps = os; }
            }
            if (key_less(pd, ps, bd, bs)) { bd = pd; bs = ps; }
        }
        if (bs == 0x7fffffff) continue;
        const double2 nb = v.XY[bs];
        double px = qx, py = qy;
        {
            const double d = dist_fma(qx, qy, nb.x, nb.y);                    // steer (:496)
            const double ux = ddiv(dsub(qx, nb.x), d), uy = ddiv(dsub(qy, nb.y), d);
            const double tx = dadd(nb.x, dmul(ux, prm.step_length)), ty = dadd(nb.y, dmul(uy, prm.step_length));
            if (d > prm.step_length) { px = tx; py = ty; }
        }
        if (point_occupied_warp(ob, px, py, prm.node_radius, lane, tests)) continue;          // :497
        bool dup = false;
        for (int s = lane; s < n_slots; s += 32) { const double2 c = v.XY[s]; dup |= (c.x == px && c.y == py); }
        if (__any_sync(kFull, dup)) { result = RRTFND_ST_DUPLICATE; break; }
        const int new_slot = n_slots++;
        ++last_id; ++n_live; ++iter;
        if (lane == 0) {
            v.XY[new_slot] = make_double2(px, py);
            v.ID[new_slot] = last_id; v.PARENT[new_slot] = bs; v.ECOST[new_slot] = dist_fma(px, py, nb.x, nb.y);   // :500-506
            v.NCHILD[new_slot] = 0; v.NFLAG[new_slot] = 0;
            v.NCHILD[bs] += 1;
        }
        __syncwarp();
        int link = -1;                                                        // :510-518, separate tree in ascending id
        for (int base = 0; base < new_slot && link < 0; base += 32) {
            const int s = base + lane;
            bool ok = false;
            if (s < new_slot && v.ID[s] >= 0 && (v.NFLAG[s] & kSeparate)) {
                // the reference tests the segment first and the distance second (:512-514); the first node in id order
                // that passes both is the same either way, and the distance leaves a handful of segments to test
                const double2 c = v.XY[s];
                if (dist_fma(px, py, c.x, c.y) < prm.step_length)
                    ok = !(seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, px, py, c.x, c.y, 0, 1) & 1);   // Line(q_new, node)
            }
            const unsigned m = __ballot_sync(kFull, ok);
            if (m) link = base + __ffs(m) - 1;
        }
        if (link >= 0) {
            if (lane == 0) {
                int node = link, parent = v.PARENT[node];                     // reconnect_ancestors (:531-543), counts kept consistent
                const uint8_t had = v.NFLAG[sep] & 1;
                if (parent >= 0) {
                    while (v.PARENT[parent] >= 0) {
                        const int pp = v.PARENT[parent];
                        v.PARENT[parent] = node; v.NCHILD[node] += 1; v.NCHILD[parent] -= 1;
                        node = parent; parent = pp;
                    }
                    v.PARENT[parent] = node; v.NCHILD[node] += 1; v.NCHILD[parent] -= 1;
                }
                const double2 c = v.XY[link];
                v.PARENT[link] = new_slot; v.NCHILD[new_slot] += 1;            // G.add_edge(node, id_new, distance) (:516)
                v.ECOST[link] = dist_fma(px, py, c.x, c.y);
                if (had) {                                                    // keep "has a goal-reaching descendant" closed upwards
                    v.NFLAG[link] |= 1;
                    for (int n = new_slot; n >= 0 && !(v.NFLAG[n] & 1); n = v.PARENT[n]) v.NFLAG[n] |= 1;
                }
            }
            __syncwarp();
            reconnected = true;
        }
    }
    const int leaf = (reconnected && v.PL->best_len > 0) ? find_slot(v.ID, n_slots, v.BPATH[0], lane) : -1;
    __syncwarp();
    if (lane == 0) {
        v.PL->n_slots = n_slots; v.PL->n_live = n_live; v.PL->last_id = last_id; v.PL->attempts = attempts; v.PL->cursor = rng.cursor;
        if (leaf >= 0) {
            int L = 0; double cost = 0.0;
            rewrite_best_path(prm, v, leaf, root, &L, &cost);
            v.PL->best_len = L; v.PL->best_cost = cost;
        }
        out_reconnected[p] = result ? result : (reconnected ? 1 : 0);
        if (out_iters) out_iters[p] = iter;
    }
}


// ---- EXTENSION: invalidation of path EDGES (DetectCollision of the RRT*FND paper; the reference's valid_path only tests the
// path NODES, src/algorithm.py:356, and its loop stops at the init_movement() stub, :283-291). Run after valid_path on the
// stored path (ids, goal side first): every consecutive pair (child = path[i], parent = path[i+1]) that is still linked
// and whose segment through_obstacle(Line(parent, child)) (:581-591; p1 = the node nearer the robot) reports blocked is
// cut - the child becomes the root of a separate tree, nothing is deleted - and the goal-side-most separate root (the
// root of the leaf's tree) is returned, which is what reconnect / regrow then mend. PARITY UNPINNED BY THE REFERENCE;
// pinned on oracle/rrt_oracle.c orc_fnd_cut_blocked_edges, composed from the reference's own through_obstacle.
__global__ void __launch_bounds__(32 * kFndWarps) cut_edges_kernel(const rrtfnd_params_t prm, const __grid_constant__ rrtfnd_batch_t bt,
                                                                   const int32_t* __restrict__ sep_in, int32_t* __restrict__ sep_out,
                                                                   int32_t* __restrict__ n_cut_out) {
    extern __shared__ int32_t fnd_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, p = blockIdx.x * kFndWarps + wib;
    if (p >= bt.n_planners) return;
    if (sep_in[p] < 0) {                                                     // negative id = leave this planner alone
        if (lane == 0) { sep_out[p] = sep_in[p]; if (n_cut_out) n_cut_out[p] = 0; }
        return;
    }
    int32_t* pslot = fnd_smem + (size_t)wib * prm.path_cap;                  // slot of path node i, -1 if it is gone
    const PlannerView v = view_of(prm, bt, p);
    const int n_slots = v.PL->n_slots, L = min(v.PL->best_len, prm.path_cap);
    const double2* tab = reinterpret_cast<const double2*>(bt.obst);
    for (int i = lane; i < L; i += 32) pslot[i] = -1;
    __syncwarp();
    for (int s = lane; s < n_slots; s += 32) {
        const int id = v.ID[s];
        if (id >= 0)
            for (int i = 0; i < L; ++i) if (v.BPATH[i] == id) { pslot[i] = s; break; }
    }
    __syncwarp();
    int n_cut = 0;
    for (int base = 0; base + 1 < L; base += 32) {                            // one lane per path edge
        const int i = base + lane;
        bool cut = false;
        if (i + 1 < L) {
            const int c = pslot[i], q = pslot[i + 1];
            if (c >= 0 && q >= 0 && v.PARENT[c] == q) {
                const double2 a = v.XY[q], b = v.XY[c];
                cut = seg_blocked(tab, &bt.grid, prm.n_obst, 0, a.x, a.y, b.x, b.y, 0, 1) & 1;   // every circle (the map has just changed)
            }
        }
        __syncwarp();
        if (cut) { v.NCHILD[pslot[i + 1]] -= 1; v.PARENT[pslot[i]] = -1; }   // distinct parents and children: no two lanes meet
        n_cut += __popc(__ballot_sync(kFull, cut));
    }
    __syncwarp();
    if (lane == 0) {
        int out = sep_in[p];
        if (n_cut > 0 && L > 0 && pslot[0] >= 0) {                            // root of the tree that holds the goal-side leaf
            int n = pslot[0];
            while (v.PARENT[n] >= 0) n = v.PARENT[n];
            if (n != v.PL->root_slot) out = v.ID[n];
        }
        sep_out[p] = out;
        if (n_cut_out) n_cut_out[p] = n_cut;
    }
}

// ---- the replanning tick's bookkeeping, on the device (rrtfnd_fnd_tick): what fnd_loop.FNDReplanner.tick() computes on the
// host from downloaded planner records, one thread per planner, so that a tick is a sequence of launches without a single
// device -> host copy. Phases: 0 MOVING, 1 ARRIVED, 2 LOST.
enum { kMoving = 0, kArrived = 1, kLost = 2 };
enum { kStMoves = 0, kStSplits, kStReconnected, kStRegrown, kStRegrowExt, kStLost, kStArrived, kStTicks, kStEdgeCuts };

__device__ __forceinline__ void stat_add(const rrtfnd_fnd_loop_t& lp, int k, long long v) {
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(lp.stats) + k, (unsigned long long)v);
}

// the robot steps to the next node of its path: current = path[-2] (fnd_loop.py step 1)
__global__ void tick_begin_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const rrtfnd_planner_t* PL = bt.planners + p;
    const int L = PL->best_len;
    const int cur = (lp.phase[p] == kMoving && L >= 2) ? bt.best_path[(size_t)p * prm.path_cap + (L - 2)] : -1;
    lp.current[p] = cur;
    for (int r = 0; r < lp.max_rounds; ++r) {
        lp.linked[(size_t)p * lp.max_rounds + r] = -1; lp.regrow[(size_t)p * lp.max_rounds + r] = 0;
        lp.regrow_iters[(size_t)p * lp.max_rounds + r] = 0;
    }
    lp.todo[p] = -1;
    stat_add(lp, kStMoves, cur >= 0);
}
// valid_path only for planners whose select_branch succeeded
__global__ void tick_after_select_kernel(const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    lp.prev_root[p] = (lp.current[p] >= 0 && lp.status[p] == 0) ? lp.current[p] : -1;
}
// who is lost, who has a split-off tree to mend
__global__ void tick_after_valid_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const rrtfnd_planner_t* PL = bt.planners + p;
    const int cur = lp.current[p];
    const bool moving = cur >= 0;
    const int root_id = bt.id[(size_t)p * prm.capacity + PL->root_slot];     // is the robot's own node still there?
    const bool bad = moving && (lp.status[p] != 0 || lp.index_error[p] != 0 || root_id < 0);
    if (bad) lp.phase[p] = kLost;
    const int sep = lp.separate[p];
    lp.todo[p] = (moving && !bad && sep != cur) ? sep : -1;
    if (lp.edge_cuts) stat_add(lp, kStEdgeCuts, lp.edge_cuts[p]);
}
// regrow only where reconnect found no link
__global__ void tick_after_reconnect_kernel(const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp, int rnd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const int todo = lp.todo[p];
    const int linked = lp.link_tmp[p];
    lp.linked[(size_t)p * lp.max_rounds + rnd] = todo >= 0 ? linked : -1;
    lp.regrow_arg[p] = (todo >= 0 && linked < 0) ? todo : -1;
}
// outcome of the round; the path now ends at the root of the next separate tree, if there is one
__global__ void tick_after_regrow_kernel(const rrtfnd_params_t prm, const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp, int rnd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    const rrtfnd_planner_t* PL = bt.planners + p;
    const bool split = lp.todo[p] >= 0, need = lp.regrow_arg[p] >= 0;
    const int res = need ? lp.regrow_tmp[p] : 0, its = need ? lp.iters_tmp[p] : 0;
    lp.regrow[(size_t)p * lp.max_rounds + rnd] = res;
    lp.regrow_iters[(size_t)p * lp.max_rounds + rnd] = its;
    const bool failed = need && res != 1;
    if (failed) lp.phase[p] = kLost;
    stat_add(lp, kStSplits, split); stat_add(lp, kStReconnected, split && !need);
    stat_add(lp, kStRegrown, need && !failed); stat_add(lp, kStRegrowExt, its);
    const int L = PL->best_len;
    const int tail = L >= 1 ? bt.best_path[(size_t)p * prm.path_cap + (L - 1)] : -1;
    const int root_id = bt.id[(size_t)p * prm.capacity + PL->root_slot];
    lp.todo[p] = (split && !failed && tail != root_id) ? tail : -1;
}
__global__ void tick_end_kernel(const rrtfnd_batch_t bt, const rrtfnd_fnd_loop_t lp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bt.n_planners) return;
    if (lp.todo[p] >= 0) lp.phase[p] = kLost;                                // still broken after max_rounds
    if (lp.phase[p] == kMoving && lp.current[p] >= 0 && bt.planners[p].best_len <= 1) lp.phase[p] = kArrived;
    stat_add(lp, kStLost, lp.phase[p] == kLost); stat_add(lp, kStArrived, lp.phase[p] == kArrived);
    if (p == 0) stat_add(lp, kStTicks, 1);
}

// ---- culling grid built on the device (rrtfnd_obst_grid_build_device): the same boxes and the same cell function as the
// host builder (host_api.cu), for obstacle tables that change every tick. One CTA; lists in ascending obstacle index.
__global__ void __launch_bounds__(256) grid_build_kernel(const double* __restrict__ obst, int M, double node_radius, rrtfnd_obst_grid_t g,
                                                         int32_t* seg_start, int32_t* seg_items, int32_t* pt_start, int32_t* pt_items,
                                                         int seg_cap, int pt_cap, int32_t* overflow) {
    const int ncells = g.nx * g.ny, tid = threadIdx.x;
    __shared__ int s_tot[2];
    auto box = [&](int j, double infl, int& x0, int& x1, int& y0, int& y1) {
        const double ox = obst[4 * j], oy = obst[4 * j + 1], r = obst[4 * j + 2];
        const double ex = fabs(r), ey = r < 0.0 ? fabs(obst[4 * j + 3]) : fabs(r);
        const double bx = __dmul_rn(__dadd_rn(ex, infl), 1.000000001), by = __dmul_rn(__dadd_rn(ey, infl), 1.000000001);
        x0 = grid_cell(__dsub_rn(ox, bx), g.x0, g.inv_cell, g.nx); x1 = grid_cell(__dadd_rn(ox, bx), g.x0, g.inv_cell, g.nx);
        y0 = grid_cell(__dsub_rn(oy, by), g.y0, g.inv_cell, g.ny); y1 = grid_cell(__dadd_rn(oy, by), g.y0, g.inv_cell, g.ny);
    };
    // per cell: count, then (after the scan) fill, scanning the obstacles in ascending index
    for (int pass = 0; pass < 2; ++pass) {
        for (int c = tid; c < ncells; c += blockDim.x) {
            const int cx = c % g.nx, cy = c / g.nx;
            int ns = 0, np = 0;
            const int bs = pass ? seg_start[c] : 0, bp = pass ? pt_start[c] : 0;
            for (int j = 0; j < M; ++j) {
                int x0, x1, y0, y1;
                box(j, g.margin, x0, x1, y0, y1);
                if (cx >= x0 && cx <= x1 && cy >= y0 && cy <= y1) { if (pass && bs + ns < seg_cap) seg_items[bs + ns] = j; ++ns; }
                box(j, fabs(node_radius), x0, x1, y0, y1);
                if (cx >= x0 && cx <= x1 && cy >= y0 && cy <= y1) { if (pass && bp + np < pt_cap) pt_items[bp + np] = j; ++np; }
            }
            if (!pass) { seg_start[c + 1] = ns; pt_start[c + 1] = np; }
        }
        __syncthreads();
        if (!pass) {
            if (tid < 2) {                                   // exclusive scans (thread 0: seg, thread 1: pt)
                int32_t* a = tid ? pt_start : seg_start;
                int run = 0;
                a[0] = 0;
                for (int c = 1; c <= ncells; ++c) { run += a[c]; a[c] = run; }
                s_tot[tid] = run;
            }
            __syncthreads();
            if (tid == 0 && (s_tot[0] > seg_cap || s_tot[1] > pt_cap)) *overflow = 1;      // sticky: the caller zeroes it
        }
    }
}

}  // namespace rrtfnd

using namespace rrtfnd;

static int fnd_check(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt) {
    if (!prm || !bt || bt->n_planners <= 0 || prm->capacity < 2) return RRTFND_E_BADARG;
    if (!bt->planners || !bt->xy || !bt->ecost || !bt->parent || !bt->nchild || !bt->id || !bt->nflags || !bt->finish || !bt->best_path)
        return RRTFND_E_BADARG;
    return 0;
}

extern "C" int rrtfnd_fnd_select_branch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const int32_t* current_id,
                                        int32_t* out_status, void* stream) {
    if (fnd_check(prm, bt) || !current_id) return RRTFND_E_BADARG;
    const int blocks = (bt->n_planners + kFndWarps - 1) / kFndWarps;
    select_branch_kernel<<<blocks, 32 * kFndWarps, 0, (cudaStream_t)stream>>>(*prm, *bt, current_id, out_status);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_fnd_valid_path(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const int32_t* previous_root_id,
                                     int32_t* out_separate_root_id, int32_t* out_index_error, void* stream) {
    if (fnd_check(prm, bt) || !previous_root_id || !out_separate_root_id || (prm->n_obst > 0 && !bt->obst)) return RRTFND_E_BADARG;
    const size_t smem = (size_t)kFndWarps * prm->path_cap * sizeof(int32_t);
    if (smem > 227 * 1024) return RRTFND_E_SMEM;
    cudaError_t e = cudaFuncSetAttribute(valid_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int blocks = (bt->n_planners + kFndWarps - 1) / kFndWarps;
    valid_path_kernel<<<blocks, 32 * kFndWarps, smem, (cudaStream_t)stream>>>(*prm, *bt, previous_root_id, out_separate_root_id,
                                                                              out_index_error);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_fnd_reconnect(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const int32_t* separate_root_id,
                                    double radius, int32_t* out_linked_id, void* stream) {
    if (fnd_check(prm, bt) || !separate_root_id || !out_linked_id || (prm->n_obst > 0 && !bt->obst)) return RRTFND_E_BADARG;
    const int blocks = (bt->n_planners + kFndWarps - 1) / kFndWarps;
    reconnect_kernel<<<blocks, 32 * kFndWarps, 0, (cudaStream_t)stream>>>(*prm, *bt, separate_root_id, radius, out_linked_id);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_fnd_regrow(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const int32_t* separate_root_id,
                                 int32_t max_iter, int32_t* out_reconnected, int32_t* out_iters, void* stream) {
    if (fnd_check(prm, bt) || !separate_root_id || !out_reconnected || max_iter < 0 || (prm->n_obst > 0 && !bt->obst)) return RRTFND_E_BADARG;
    if (prm->stream_mode == RRTFND_STREAM_WORDS && !bt->words) return RRTFND_E_BADARG;
    const int blocks = (bt->n_planners + kFndWarps - 1) / kFndWarps;
    regrow_kernel<<<blocks, 32 * kFndWarps, 0, (cudaStream_t)stream>>>(*prm, *bt, separate_root_id, max_iter, out_reconnected, out_iters);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_fnd_cut_blocked_edges(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const int32_t* separate_root_id,
                                            int32_t* out_separate_root_id, int32_t* out_n_cut, void* stream) {
    if (fnd_check(prm, bt) || !separate_root_id || !out_separate_root_id || (prm->n_obst > 0 && !bt->obst)) return RRTFND_E_BADARG;
    const size_t smem = (size_t)kFndWarps * prm->path_cap * sizeof(int32_t);
    if (smem > 227 * 1024) return RRTFND_E_SMEM;
    cudaError_t e = cudaFuncSetAttribute(cut_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int blocks = (bt->n_planners + kFndWarps - 1) / kFndWarps;
    cut_edges_kernel<<<blocks, 32 * kFndWarps, smem, (cudaStream_t)stream>>>(*prm, *bt, separate_root_id, out_separate_root_id, out_n_cut);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_fnd_tick(const rrtfnd_params_t* prm, const rrtfnd_batch_t* bt, const rrtfnd_fnd_loop_t* lp,
                               double reconnect_radius, int32_t regrow_iters, int32_t edge_invalidation, void* stream) {
    if (fnd_check(prm, bt) || !lp || lp->max_rounds < 1 || regrow_iters < 0) return RRTFND_E_BADARG;
    if (!lp->phase || !lp->current || !lp->status || !lp->prev_root || !lp->separate || !lp->index_error || !lp->todo ||
        !lp->regrow_arg || !lp->link_tmp || !lp->regrow_tmp || !lp->iters_tmp || !lp->linked || !lp->regrow || !lp->regrow_iters ||
        !lp->stats) return RRTFND_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = bt->n_planners, nt = 128, nb = (P + nt - 1) / nt;
    int rc;
    tick_begin_kernel<<<nb, nt, 0, st>>>(*prm, *bt, *lp);
    if ((rc = rrtfnd_fnd_select_branch(prm, bt, lp->current, lp->status, stream))) return rc;
    tick_after_select_kernel<<<nb, nt, 0, st>>>(*bt, *lp);
    if ((rc = rrtfnd_fnd_valid_path(prm, bt, lp->prev_root, lp->separate, lp->index_error, stream))) return rc;
    if (edge_invalidation) {
        if (!lp->edge_cuts) return RRTFND_E_BADARG;
        if ((rc = rrtfnd_fnd_cut_blocked_edges(prm, bt, lp->separate, lp->separate, lp->edge_cuts, stream))) return rc;
    }
    tick_after_valid_kernel<<<nb, nt, 0, st>>>(*prm, *bt, *lp);
    for (int rnd = 0; rnd < lp->max_rounds; ++rnd) {      // one separate tree per round, goal side first
        if ((rc = rrtfnd_fnd_reconnect(prm, bt, lp->todo, reconnect_radius, lp->link_tmp, stream))) return rc;
        tick_after_reconnect_kernel<<<nb, nt, 0, st>>>(*bt, *lp, rnd);
        if ((rc = rrtfnd_fnd_regrow(prm, bt, lp->regrow_arg, regrow_iters, lp->regrow_tmp, lp->iters_tmp, stream))) return rc;
        tick_after_regrow_kernel<<<nb, nt, 0, st>>>(*prm, *bt, *lp, rnd);
    }
    tick_end_kernel<<<nb, nt, 0, st>>>(*bt, *lp);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_obst_grid_build_device(const double* obst, int32_t n_obst, double node_radius, double lo_x, double hi_x,
                                             double lo_y, double hi_y, double cell, double coord_max, rrtfnd_obst_grid_t* grid,
                                             int32_t seg_cap, int32_t pt_cap, int32_t cells_cap, int32_t* overflow, void* stream) {
    if (!grid || n_obst < 0 || (n_obst > 0 && !obst) || !(hi_x > lo_x) || !(hi_y > lo_y) || !overflow || !(coord_max > 0.0) ||
        !(coord_max < 1e100)) return RRTFND_E_BADARG;
    if (!grid->seg_start || !grid->seg_items || !grid->pt_start || !grid->pt_items) return RRTFND_E_BADARG;
    int32_t* ss = const_cast<int32_t*>(grid->seg_start); int32_t* si = const_cast<int32_t*>(grid->seg_items);
    int32_t* ps = const_cast<int32_t*>(grid->pt_start); int32_t* pi = const_cast<int32_t*>(grid->pt_items);
    const double slope_max = rrtfnd::grid_slope_max();
    const double w = hi_x - lo_x, h = hi_y - lo_y;
    if (n_obst == 0) { grid->nx = grid->ny = 0; return 0; }
    if (!(cell > 0.0)) cell = sqrt(w * h / (double)n_obst);
    cell = fmax(cell, fmax(w, h) / 512.0);
    const int nx = (int)fmin(512.0, fmax(1.0, ceil(w / cell))), ny = (int)fmin(512.0, fmax(1.0, ceil(h / cell)));
    if ((long long)nx * ny + 1 > cells_cap) return RRTFND_E_BADARG;
    grid->x0 = lo_x; grid->y0 = lo_y; grid->inv_cell = 1.0 / cell;
    grid->slope_max = slope_max; grid->coord_max = coord_max; grid->margin = rrtfnd::grid_seg_margin(coord_max, slope_max);
    grid->nx = nx; grid->ny = ny;
    grid_build_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(obst, n_obst, node_radius, *grid, ss, si, ps, pi, seg_cap, pt_cap, overflow);
    return (int)cudaGetLastError();
}
