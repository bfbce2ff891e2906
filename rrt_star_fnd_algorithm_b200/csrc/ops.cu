// ops.cu — stateless batched operators behind the C ABI (include/rrtfnd.h).
//
// These are the reference's per-iteration primitives as stand-alone kernels, for callers that drive
// the loop themselves (and for operator-level parity tests):
//   edges_vs_circles  through_obstacle / intersection_circle   src/algorithm.py:555-635, src/line.py:5-16
//   points_occupied   Map.is_occupied_c                        src/map.py:38-47
//   nearest_visible   nearest_node_kdtree                      src/algorithm.py:666-699
//   near_set          cKDTree.query_ball_point                 src/algorithm.py:837, :893
//   chain_cost        Graph.get_cost                           src/graph.py:34-46
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "cull.cuh"

namespace rrtfnd {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kObstTile = 1024;  // circles staged per shared-memory tile (32 KB as SoA of 4 doubles)

__device__ __forceinline__ void stage_obstacles(const double* __restrict__ obst, int j0, int n, double* ox,
                                                double* oy, double* orr, double* oK) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double4 v = reinterpret_cast<const double4*>(obst)[j0 + j];   // 32 B row: ox, oy, r, K
        ox[j] = v.x; oy[j] = v.y; orr[j] = v.z; oK[j] = v.w;
    }
}

// One warp per edge; lanes stride over the obstacle tile; __any_sync vote with early exit per tile row.
__global__ void __launch_bounds__(256) edges_vs_circles_kernel(const double* __restrict__ p1x, const double* __restrict__ p1y,
                                                               const double* __restrict__ p2x, const double* __restrict__ p2y,
                                                               long long n_edges, const double* __restrict__ obst, int M,
                                                               uint8_t* __restrict__ out_hit) {
    __shared__ __align__(16) double ox[kObstTile], oy[kObstTile], orr[kObstTile], oK[kObstTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long n_rounds = (n_edges + (long long)gridDim.x * wpb - 1) / ((long long)gridDim.x * wpb);
    for (long long r = 0; r < n_rounds; ++r) {
        const long long e = (r * gridDim.x + blockIdx.x) * wpb + warp;
        const bool live = e < n_edges;
        Seg sg;
        if (live) sg = make_seg(p1x[e], p1y[e], p2x[e], p2y[e]);
        bool hit = false;
        for (int j0 = 0; j0 < M; j0 += kObstTile) {                     // tiles (one for M <= 1024)
            const int n = min(kObstTile, M - j0);
            if (M > kObstTile || r == 0) {
                __syncthreads();
                stage_obstacles(obst, j0, n, ox, oy, orr, oK);
                __syncthreads();
            }
            if (live && !hit) {                                         // warp-uniform condition
                for (int jj = 0; jj < n; jj += 32) {
                    const int j = jj + lane;
                    bool h = false;
                    if (j < n) h = seg_hits_circle(sg, ox[j], oy[j], orr[j], oK[j]);
                    if (__any_sync(kFullMask, h)) { hit = true; break; }
                }
            }
        }
        if (live && lane == 0) out_hit[e] = hit ? 1 : 0;
    }
}

// One warp per point.
__global__ void __launch_bounds__(256) points_occupied_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                              long long n_points, const double* __restrict__ obst, int M,
                                                              double node_radius, uint8_t* __restrict__ out) {
    __shared__ __align__(16) double ox[kObstTile], oy[kObstTile], orr[kObstTile], oK[kObstTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long n_rounds = (n_points + (long long)gridDim.x * wpb - 1) / ((long long)gridDim.x * wpb);
    for (long long r = 0; r < n_rounds; ++r) {
        const long long e = (r * gridDim.x + blockIdx.x) * wpb + warp;
        const bool live = e < n_points;
        double x = 0, y = 0;
        if (live) { x = px[e]; y = py[e]; }
        bool occ = false;
        for (int j0 = 0; j0 < M; j0 += kObstTile) {
            const int n = min(kObstTile, M - j0);
            if (M > kObstTile || r == 0) {
                __syncthreads();
                stage_obstacles(obst, j0, n, ox, oy, orr, oK);
                __syncthreads();
            }
            if (live && !occ) {
                bool h = false;
                for (int j = lane; j < n; j += 32) h |= point_in_inflated(x, y, ox[j], oy[j], orr[j], node_radius, oK[j]);
                occ = __any_sync(kFullMask, h);
            }
        }
        if (live && lane == 0) out[e] = occ ? 1 : 0;
    }
}

// The culled forms used inside the fused planner, as stand-alone kernels: one LANE per edge / one warp per point.
__global__ void __launch_bounds__(128) edges_vs_circles_grid_kernel(const double* __restrict__ p1x, const double* __restrict__ p1y,
                                                                    const double* __restrict__ p2x, const double* __restrict__ p2y,
                                                                    long long n_edges, const double* __restrict__ obst, int M,
                                                                    const __grid_constant__ rrtfnd_obst_grid_t grid,
                                                                    uint8_t* __restrict__ out_hit) {
    const Obst ob = make_obst(obst, M, &grid);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x)
        out_hit[e] = seg_blocked(ob.tab, ob.g, ob.M, ob.full_cells, p1x[e], p1y[e], p2x[e], p2y[e], 0, 1) & 1;
}

__global__ void __launch_bounds__(128) points_occupied_grid_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                                   long long n_points, const double* __restrict__ obst, int M,
                                                                   const __grid_constant__ rrtfnd_obst_grid_t grid,
                                                                   double node_radius, uint8_t* __restrict__ out) {
    const Obst ob = make_obst(obst, M, &grid);
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long tests = 0;
    for (long long e = warp0; e < n_points; e += nwarps) {
        const bool occ = point_occupied_warp(ob, px[e], py[e], node_radius, lane, tests);
        if (lane == 0) out[e] = occ ? 1 : 0;
    }
}

template <int NT>
__device__ __forceinline__ void block_argmin(double& bd, int& bs, double* red_d, int* red_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(kFullMask, bd, off);
        const int os = __shfl_xor_sync(kFullMask, bs, off);
        if (key_less(od, os, bd, bs)) { bd = od; bs = os; }
    }
    __syncthreads();
    if (lane == 0) { red_d[warp] = bd; red_s[warp] = bs; }
    __syncthreads();
    bd = red_d[0]; bs = red_s[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w)
        if (key_less(red_d[w], red_s[w], bd, bs)) { bd = red_d[w]; bs = red_s[w]; }
}

// One CTA per planner: repeated (d2, slot) argmin passes over x[], y[] (double2-vectorised, coalesced)
// until the candidate's segment to q is collision free. Tests every circle (the un-culled cross-check of planner.cu).
template <int NT>
__global__ void __launch_bounds__(NT) nearest_visible_kernel(const rrtfnd_batch_t bt, int C, const double* __restrict__ q,
                                                             int M, int* __restrict__ out_slot, int* __restrict__ out_rounds) {
    __shared__ __align__(16) double ox[kObstTile], oy[kObstTile], orr[kObstTile], oK[kObstTile];
    __shared__ double red_d[32];
    __shared__ int red_s[32];
    const int p = blockIdx.x, tid = threadIdx.x;
    const double2* __restrict__ XY = reinterpret_cast<const double2*>(bt.xy) + (size_t)p * C;
    const int n_slots = bt.planners[p].n_slots, n_live = bt.planners[p].n_live;
    const double qx = q[2 * p], qy = q[2 * p + 1];
    stage_obstacles(bt.obst, 0, min(M, kObstTile), ox, oy, orr, oK);   // M <= kObstTile enforced by the host wrapper
    __syncthreads();
    double prev_d = -1.0; int prev_s = -1, result = -1, rounds = 0;
    for (int round = 0; round < n_live; ++round) {
        double bd = CUDART_INF; int bs = 0x7fffffff; bool exact = false;
        for (int s = tid; s < n_slots; s += NT) {                        // one coalesced 16-byte load per node
            const double2 v = XY[s];
            const double d = dist2(v.x, v.y, qx, qy);
            if (v.x == qx && v.y == qy) exact = true;
            if ((d > prev_d || (d == prev_d && s > prev_s)) && d < bd) { bd = d; bs = s; }
        }
        if (round == 0 && __syncthreads_or(exact)) { result = -2; break; }
        block_argmin<NT>(bd, bs, red_d, red_s);
        if (bs == 0x7fffffff) break;
        ++rounds;
        const double2 nb = XY[bs];
        const Seg sg = make_seg(qx, qy, nb.x, nb.y);
        bool hit = false;
        for (int j = tid; j < M; j += NT) hit |= seg_hits_circle(sg, ox[j], oy[j], orr[j], oK[j]);
        if (!__syncthreads_or(hit)) { result = bs; break; }
        prev_d = bd; prev_s = bs;
    }
    if (tid == 0) { out_slot[p] = result; if (out_rounds) out_rounds[p] = rounds; }
}

// One CTA per planner: ball predicate d2 <= r*r, ballot + popc prefix -> ascending compacted slot list.
template <int NT>
__global__ void __launch_bounds__(NT) near_set_kernel(const rrtfnd_batch_t bt, int C, const double* __restrict__ q, double radius,
                                                      int near_cap, int* __restrict__ out_slots, int* __restrict__ out_count) {
    __shared__ int warp_cnt[NT / 32];
    __shared__ int running;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2* __restrict__ XY = reinterpret_cast<const double2*>(bt.xy) + (size_t)p * C;
    const int n_slots = bt.planners[p].n_slots;
    const double qx = q[2 * p], qy = q[2 * p + 1];
    const double r2 = dmul(radius, radius);
    int* out = out_slots + (size_t)p * near_cap;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n_slots; base += NT) {
        const int s = base + tid;
        bool in = false;
        if (s < n_slots) { const double2 v = XY[s]; in = dist2(v.x, v.y, qx, qy) <= r2; }
        const unsigned b = __ballot_sync(kFullMask, in);
        if (lane == 0) warp_cnt[warp] = __popc(b);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        if (in) {
            const int pos = before + __popc(b & ((1u << lane) - 1u));
            if (pos < near_cap) out[pos] = s;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < NT / 32; ++w) t += warp_cnt[w]; running += t; }
        __syncthreads();
    }
    if (tid == 0) out_count[p] = running;
}

__global__ void chain_cost_kernel(const rrtfnd_batch_t bt, int C, const int* __restrict__ planner_idx,
                                  const int* __restrict__ slot, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = planner_idx[i];
    const int32_t* parent = bt.parent + (size_t)p * C;
    const double* ecost = bt.ecost + (size_t)p * C;
    double acc = 0.0;
    int s = slot[i];
    int par = parent[s];
    while (par >= 0) { acc = dadd(acc, ecost[s]); s = par; par = parent[s]; }
    out[i] = acc;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace rrtfnd

using namespace rrtfnd;

extern "C" int rrtfnd_abi_version(void) { return RRTFND_ABI_VERSION; }

extern "C" const char* rrtfnd_status_string(int st) {
    switch (st) {
        case RRTFND_ST_RUNNING: return "running";
        case RRTFND_ST_DONE: return "done";
        case RRTFND_ST_NEED_WORDS: return "need-words";
        case RRTFND_ST_ATTEMPT_CAP: return "attempt-cap";
        case RRTFND_ST_CAPACITY: return "node capacity exhausted";
        case RRTFND_ST_NEAR_CAP: return "more nodes rewired in one iteration than near_cap";
        case RRTFND_ST_FINISH_CAP: return "finish list full";
        case RRTFND_ST_PATH_CAP: return "best path longer than path_cap";
        case RRTFND_ST_DUPLICATE: return "steered point coincides with an existing vertex";
        case RRTFND_ST_NO_CHILDLESS: return "forced_removal has no removable childless node";
        case RRTFND_ST_STREAM_EXHAUSTED: return "random word buffer exhausted inside forced_removal";
        case RRTFND_ST_REGROW_NEED_WORDS: return "regrow: random word buffer exhausted, refill and call again";
        case RRTFND_E_BADARG: return "bad argument";
        case RRTFND_E_NOGPU: return "no CUDA device";
        case RRTFND_E_SMEM: return "problem does not fit in shared memory";
        default: return "unknown";
    }
}

extern "C" int rrtfnd_edges_vs_circles(const double* p1x, const double* p1y, const double* p2x, const double* p2y,
                                       int64_t n_edges, const double* obst, int32_t n_obst, uint8_t* out_hit,
                                       void* stream) {
    if (n_edges < 0 || n_obst < 0 || (n_edges > 0 && (!p1x || !p1y || !p2x || !p2y || !out_hit))) return RRTFND_E_BADARG;
    if (n_edges == 0) return 0;
    const int wpb = 8;
    long long blocks = (n_edges + wpb - 1) / wpb;
    const long long cap = (long long)sm_count() * 8;     // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    edges_vs_circles_kernel<<<(int)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(p1x, p1y, p2x, p2y, n_edges, obst, n_obst, out_hit);
    return (int)cudaGetLastError();
}

static int check_grid(const rrtfnd_obst_grid_t* g) {
    if (!g) return RRTFND_E_BADARG;
    if (g->nx > 0 && (g->ny <= 0 || !g->seg_start || !g->seg_items || !g->pt_start || !g->pt_items)) return RRTFND_E_BADARG;
    return 0;
}

extern "C" int rrtfnd_edges_vs_circles_grid(const double* p1x, const double* p1y, const double* p2x, const double* p2y,
                                            int64_t n_edges, const double* obst, int32_t n_obst,
                                            const rrtfnd_obst_grid_t* grid, uint8_t* out_hit, void* stream) {
    if (n_edges < 0 || n_obst < 0 || check_grid(grid) || (n_edges > 0 && (!p1x || !p1y || !p2x || !p2y || !out_hit))) return RRTFND_E_BADARG;
    if (n_edges == 0) return 0;
    long long blocks = (n_edges + 127) / 128;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    edges_vs_circles_grid_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(p1x, p1y, p2x, p2y, n_edges, obst, n_obst, *grid, out_hit);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_points_occupied_grid(const double* px, const double* py, int64_t n_points, const double* obst,
                                           int32_t n_obst, const rrtfnd_obst_grid_t* grid, double node_radius,
                                           uint8_t* out_occ, void* stream) {
    if (n_points < 0 || n_obst < 0 || check_grid(grid) || (n_points > 0 && (!px || !py || !out_occ))) return RRTFND_E_BADARG;
    if (n_points == 0) return 0;
    long long blocks = (n_points + 3) / 4;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    points_occupied_grid_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(px, py, n_points, obst, n_obst, *grid, node_radius, out_occ);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_points_occupied(const double* px, const double* py, int64_t n_points, const double* obst,
                                      int32_t n_obst, double node_radius, uint8_t* out_occ, void* stream) {
    if (n_points < 0 || n_obst < 0 || (n_points > 0 && (!px || !py || !out_occ))) return RRTFND_E_BADARG;
    if (n_points == 0) return 0;
    const int wpb = 8;
    long long blocks = (n_points + wpb - 1) / wpb;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    points_occupied_kernel<<<(int)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(px, py, n_points, obst, n_obst, node_radius, out_occ);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_nearest_visible(const rrtfnd_batch_t* bt, int32_t capacity, const double* q, int32_t n_obst,
                                      int32_t* out_slot, int32_t* out_rounds, void* stream) {
    if (!bt || bt->n_planners <= 0 || !q || !out_slot || n_obst < 0 || n_obst > kObstTile) return RRTFND_E_BADARG;
    nearest_visible_kernel<256><<<bt->n_planners, 256, 0, (cudaStream_t)stream>>>(*bt, capacity, q, n_obst, out_slot, out_rounds);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_near_set(const rrtfnd_batch_t* bt, int32_t capacity, const double* q, double radius,
                               int32_t near_cap, int32_t* out_slots, int32_t* out_count, void* stream) {
    if (!bt || bt->n_planners <= 0 || !q || !out_slots || !out_count || near_cap < 1) return RRTFND_E_BADARG;
    near_set_kernel<256><<<bt->n_planners, 256, 0, (cudaStream_t)stream>>>(*bt, capacity, q, radius, near_cap, out_slots, out_count);
    return (int)cudaGetLastError();
}

extern "C" int rrtfnd_chain_cost(const rrtfnd_batch_t* bt, int32_t capacity, const int32_t* planner_idx,
                                 const int32_t* slot, int64_t n, double* out_cost, void* stream) {
    if (!bt || n < 0 || (n > 0 && (!planner_idx || !slot || !out_cost))) return RRTFND_E_BADARG;
    if (n == 0) return 0;
    chain_cost_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*bt, capacity, planner_idx, slot, n, out_cost);
    return (int)cudaGetLastError();
}

// ---- `**` as the reference's interpreter evaluates it (libm pow of glibc 2.39, csrc/libm_pow.h) --------------------------
#include "libm_pow.h"

namespace rrtfnd {
__global__ void pow_libm_kernel(const double* __restrict__ x, long long n, int which, double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = which ? rrtfnd_libm_pow_half(x[i]) : rrtfnd_libm_pow2(x[i]);
}
}  // namespace rrtfnd

extern "C" int rrtfnd_pow_libm(const double* x, int64_t n, int32_t which, double* out, void* stream) {
    if (n < 0 || (n > 0 && (!x || !out)) || (which != 0 && which != 1)) return RRTFND_E_BADARG;
    if (n == 0) return 0;
    const int nt = 256;
    const long long want = (n + nt - 1) / nt;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);             // grid-stride over a multiple of the SM count
    rrtfnd::pow_libm_kernel<<<blocks, nt, 0, (cudaStream_t)stream>>>(x, (long long)n, which, out);
    return (int)cudaGetLastError();
}
