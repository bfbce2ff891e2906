// host_api.cu — host-buffer entry point of the C ABI: rrtfnd_plan_host.
//
// This is what the Python drop-in (RRT / RRT_star / RRT_star_FN with reference signatures) and the
// end-to-end leg of bench.py call: inputs and outputs are HOST buffers, the library owns the device
// memory for the duration of the call, and the host<->device copies are part of the call.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <stdio.h>
#include <chrono>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/rrtfnd.h"
#include "grid.cuh"

namespace {

// Device memory of rrtfnd_plan_host: one grow-only arena per process, kept between calls (cudaMalloc / cudaFree of a few
// hundred megabytes cost tens of milliseconds, a quarter of a config-2 call) and released by rrtfnd_host_release(). Calls are
// serialised by the arena's mutex; a call carves its buffers out of the arena with 256-byte alignment.
struct Arena {
    std::mutex lock;
    void* base = nullptr;
    size_t cap = 0, used = 0;
    int device = -1;
    cudaError_t reserve(int dev, size_t bytes) {
        used = 0;
        if (base && device == dev && cap >= bytes) return cudaSuccess;
        if (base) { cudaSetDevice(device); cudaFree(base); base = nullptr; cap = 0; cudaSetDevice(dev); }
        cudaError_t e = cudaMalloc(&base, bytes);
        if (e == cudaSuccess) { cap = bytes; device = dev; } else base = nullptr;
        return e;
    }
    void* take(size_t bytes) { void* p = (char*)base + used; used += (bytes + 255) / 256 * 256; return p; }
    void release() { if (base) { cudaSetDevice(device); cudaFree(base); } base = nullptr; cap = used = 0; device = -1; }
};
Arena g_arena;

#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return (int)e_; } while (0)

}  // namespace

// ---- obstacle grid builder (pure host code) --------------------------------------------------------------
static int grid_build(const double* obst_host, int stride, int32_t M, double node_radius, double lo_x,
                      double hi_x, double lo_y, double hi_y, double cell,
                      rrtfnd_obst_grid_t* g, int32_t* seg_start_host, int32_t* seg_items_host,
                      int32_t* pt_start_host, int32_t* pt_items_host, int64_t* n_items_out) {
    if (!g || M < 0 || (M > 0 && !obst_host) || !(hi_x > lo_x) || !(hi_y > lo_y) || !n_items_out) return RRTFND_E_BADARG;
    memset(g, 0, sizeof *g);
    n_items_out[0] = n_items_out[1] = 0;
    if (M == 0) return 0;                                       // nx == 0: culling disabled
    double S = fmax(fmax(fabs(lo_x), fabs(hi_x)), fmax(fabs(lo_y), fabs(hi_y)));
    for (int j = 0; j < M; ++j) {
        for (int c = 0; c < 3; ++c) S = fmax(S, fabs(obst_host[stride * j + c]));
        if (stride == 4 && obst_host[4 * j + 2] < 0.0) S = fmax(S, fabs(obst_host[4 * j + 3]));   // a rectangle's half height
    }
    if (!(S < 1e100)) return RRTFND_E_BADARG;
    const double slope_max = rrtfnd::grid_slope_max();
    const double margin = rrtfnd::grid_seg_margin(S, slope_max);
    const double w = hi_x - lo_x, h = hi_y - lo_y;
    if (!(cell > 0.0)) cell = sqrt(w * h / (double)M);
    cell = fmax(cell, fmax(w, h) / 512.0);
    const int nx = (int)fmin(512.0, fmax(1.0, ceil(w / cell))), ny = (int)fmin(512.0, fmax(1.0, ceil(h / cell)));
    g->x0 = lo_x; g->y0 = lo_y; g->inv_cell = 1.0 / cell;
    g->slope_max = slope_max; g->coord_max = S; g->margin = margin; g->nx = nx; g->ny = ny;
    const size_t ncells = (size_t)nx * ny;
    std::vector<int32_t> cnt_s(ncells + 1, 0), cnt_p(ncells + 1, 0);
    struct Box { int x0, x1, y0, y1; };
    std::vector<Box> bs((size_t)M), bp((size_t)M);
    for (int j = 0; j < M; ++j) {
        const double ox = obst_host[stride * j], oy = obst_host[stride * j + 1], r = obst_host[stride * j + 2];
        // half extents of the obstacle's own box: (r, r) for a circle, (hx, hy) for a rectangle row (cx, cy, -hx, hy)
        const bool rect = stride == 4 && r < 0.0;
        const double ex = fabs(r), ey = rect ? fabs(obst_host[4 * j + 3]) : fabs(r);
        // segments: inflate by the margin; points: by node_radius (Map.is_occupied_c); both with relative slack
        const double Sx = (ex + margin) * 1.000000001, Sy = (ey + margin) * 1.000000001;
        const double Px = (ex + fabs(node_radius)) * 1.000000001, Py = (ey + fabs(node_radius)) * 1.000000001;
        bs[j] = Box{rrtfnd::grid_cell(ox - Sx, g->x0, g->inv_cell, nx), rrtfnd::grid_cell(ox + Sx, g->x0, g->inv_cell, nx),
                    rrtfnd::grid_cell(oy - Sy, g->y0, g->inv_cell, ny), rrtfnd::grid_cell(oy + Sy, g->y0, g->inv_cell, ny)};
        bp[j] = Box{rrtfnd::grid_cell(ox - Px, g->x0, g->inv_cell, nx), rrtfnd::grid_cell(ox + Px, g->x0, g->inv_cell, nx),
                    rrtfnd::grid_cell(oy - Py, g->y0, g->inv_cell, ny), rrtfnd::grid_cell(oy + Py, g->y0, g->inv_cell, ny)};
        for (int cy = bs[j].y0; cy <= bs[j].y1; ++cy) for (int cx = bs[j].x0; cx <= bs[j].x1; ++cx) ++cnt_s[(size_t)cy * nx + cx + 1];
        for (int cy = bp[j].y0; cy <= bp[j].y1; ++cy) for (int cx = bp[j].x0; cx <= bp[j].x1; ++cx) ++cnt_p[(size_t)cy * nx + cx + 1];
    }
    for (size_t c = 0; c < ncells; ++c) { cnt_s[c + 1] += cnt_s[c]; cnt_p[c + 1] += cnt_p[c]; }
    n_items_out[0] = cnt_s[ncells]; n_items_out[1] = cnt_p[ncells];
    if (!seg_items_host) return 0;                              // sizing call
    if (!seg_start_host || !pt_start_host || !pt_items_host) return RRTFND_E_BADARG;
    memcpy(seg_start_host, cnt_s.data(), (ncells + 1) * sizeof(int32_t));
    memcpy(pt_start_host, cnt_p.data(), (ncells + 1) * sizeof(int32_t));
    std::vector<int32_t> fs(cnt_s.begin(), cnt_s.end() - 1), fp(cnt_p.begin(), cnt_p.end() - 1);
    for (int j = 0; j < M; ++j) {                               // ascending obstacle index inside every cell
        for (int cy = bs[j].y0; cy <= bs[j].y1; ++cy) for (int cx = bs[j].x0; cx <= bs[j].x1; ++cx) seg_items_host[fs[(size_t)cy * nx + cx]++] = j;
        for (int cy = bp[j].y0; cy <= bp[j].y1; ++cy) for (int cx = bp[j].x0; cx <= bp[j].x1; ++cx) pt_items_host[fp[(size_t)cy * nx + cx]++] = j;
    }
    return 0;
}

extern "C" int rrtfnd_obst_grid_build_host(const double* obst_host, int32_t M, double node_radius, double lo_x, double hi_x,
                                           double lo_y, double hi_y, double cell, rrtfnd_obst_grid_t* g, int32_t* seg_start_host,
                                           int32_t* seg_items_host, int32_t* pt_start_host, int32_t* pt_items_host,
                                           int64_t* n_items_out) {
    return grid_build(obst_host, 3, M, node_radius, lo_x, hi_x, lo_y, hi_y, cell, g, seg_start_host, seg_items_host, pt_start_host,
                      pt_items_host, n_items_out);
}

extern "C" int rrtfnd_obst_grid_build_host4(const double* obst4_host, int32_t M, double node_radius, double lo_x, double hi_x,
                                            double lo_y, double hi_y, double cell, rrtfnd_obst_grid_t* g, int32_t* seg_start_host,
                                            int32_t* seg_items_host, int32_t* pt_start_host, int32_t* pt_items_host,
                                            int64_t* n_items_out) {
    return grid_build(obst4_host, 4, M, node_radius, lo_x, hi_x, lo_y, hi_y, cell, g, seg_start_host, seg_items_host, pt_start_host,
                      pt_items_host, n_items_out);
}

extern "C" int rrtfnd_plan_host(const rrtfnd_params_t* prm, int32_t P, const double* starts_goals_host,
                                const uint64_t* seeds_host, const double* obst_host, const uint32_t* words_host,
                                rrtfnd_planner_t* planners_host, double* xy_host, double* ecost_host,
                                int32_t* parent_host, int32_t* nchild_host, int32_t* id_host, int32_t* best_path_host,
                                int32_t device) {
    if (!prm || P <= 0 || !starts_goals_host || !seeds_host || (prm->n_obst > 0 && !obst_host)) return RRTFND_E_BADARG;
    const bool timing = getenv("RRTFND_HOST_TIMING") != nullptr;     // phase times on stderr (diagnostics)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    if (prm->stream_mode == RRTFND_STREAM_WORDS && !words_host) return RRTFND_E_BADARG;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return RRTFND_E_NOGPU;
    CK(cudaSetDevice(device));
    const size_t C = (size_t)prm->capacity, M = (size_t)prm->n_obst, n = (size_t)P * C;

    std::vector<rrtfnd_planner_t> pl((size_t)P);
    memset(pl.data(), 0, pl.size() * sizeof(rrtfnd_planner_t));
    for (int p = 0; p < P; ++p) {
        pl[p].start_x = starts_goals_host[4 * p + 0]; pl[p].start_y = starts_goals_host[4 * p + 1];
        pl[p].goal_x = starts_goals_host[4 * p + 2];  pl[p].goal_y = starts_goals_host[4 * p + 3];
        pl[p].seed = seeds_host[2 * p]; pl[p].stream_id = seeds_host[2 * p + 1];
    }
    // obstacle rows ox, oy, r, K. K is the circle-only part of calc_delta's `c`, `-r ** 2 + x0 ** 2 + y0 ** 2`
    // (src/algorithm.py:619). The reference's `**` on floats is the C library's pow(), which is NOT always fl(x * x), so K
    // is evaluated with pow() here as well - the same call, on the same host libm, in the same left-to-right order as the
    // interpreter makes (PlannerBatch evaluates the Python expression itself; tests/test_headline_paths.py requires both
    // tables to be bit-equal). The exponent is volatile so that the host compiler cannot fold pow(x, 2.0) into x * x.
    // A caller whose obstacles are not plain floats (Python ints are squared exactly) passes its own K:
    // RRTFND_FLAG_OBST_HAS_K makes obst_host [M][4] = (ox, oy, r, K).
    const bool has_k = (prm->flags & RRTFND_FLAG_OBST_HAS_K) != 0;
    const size_t ostride = has_k ? 4 : 3;
    std::vector<double> ob(4 * M + 4);
    for (size_t j = 0; j < M; ++j) {
        const double ox = obst_host[ostride * j], oy = obst_host[ostride * j + 1], r = obst_host[ostride * j + 2];
        double K;
        if (!has_k && r < 0.0) return RRTFND_E_BADARG;          // rectangle rows need their fourth column (RRTFND_FLAG_OBST_HAS_K)
        if (has_k) K = obst_host[4 * j + 3];
        else {
            volatile double two = 2.0;
            const volatile double r2 = pow(r, two), x2 = pow(ox, two), y2 = pow(oy, two);
            const volatile double t = -r2 + x2;
            K = t + y2;
        }
        ob[4 * j] = ox; ob[4 * j + 1] = oy; ob[4 * j + 2] = r; ob[4 * j + 3] = K;
    }

    // culling grid over the box that bounds xy_range and every start / goal
    double lo_x = prm->x_lo, hi_x = prm->x_hi, lo_y = prm->y_lo, hi_y = prm->y_hi;
    for (int p = 0; p < P; ++p) {
        lo_x = fmin(lo_x, fmin(pl[p].start_x, pl[p].goal_x)); hi_x = fmax(hi_x, fmax(pl[p].start_x, pl[p].goal_x));
        lo_y = fmin(lo_y, fmin(pl[p].start_y, pl[p].goal_y)); hi_y = fmax(hi_y, fmax(pl[p].start_y, pl[p].goal_y));
    }
    rrtfnd_obst_grid_t grid;
    memset(&grid, 0, sizeof grid);
    std::vector<int32_t> g_ss, g_si, g_ps, g_pi;
    if (M > 0 && hi_x > lo_x && hi_y > lo_y) {
        int64_t cnt[2];
        int grc = rrtfnd_obst_grid_build_host4(ob.data(), (int32_t)M, prm->node_radius, lo_x, hi_x, lo_y, hi_y, 0.0, &grid, 0, 0, 0, 0, cnt);
        if (grc) return grc;
        const size_t ncells = (size_t)grid.nx * grid.ny;
        g_ss.resize(ncells + 1); g_ps.resize(ncells + 1); g_si.resize((size_t)cnt[0] + 1); g_pi.resize((size_t)cnt[1] + 1);
        grc = rrtfnd_obst_grid_build_host4(ob.data(), (int32_t)M, prm->node_radius, lo_x, hi_x, lo_y, hi_y, 0.0, &grid, g_ss.data(),
                                          g_si.data(), g_ps.data(), g_pi.data(), cnt);
        if (grc) return grc;
    }

    std::lock_guard<std::mutex> guard(g_arena.lock);
    const size_t wb = prm->stream_mode == RRTFND_STREAM_WORDS ? (size_t)P * (size_t)prm->words_per_planner * 4 : 0;
    const size_t sizes[] = {(size_t)P * sizeof(rrtfnd_planner_t), n * 16, n * 8, n * 4, n * 4, n * 4, n, (size_t)P * prm->finish_cap * 4,
                            (size_t)P * prm->path_cap * 4, ob.size() * 8, wb, g_ss.size() * 4, g_si.size() * 4, g_ps.size() * 4, g_pi.size() * 4};
    size_t total = 0;
    for (size_t b_ : sizes) total += (b_ + 255) / 256 * 256 + 256;
    CK(g_arena.reserve(device, total));
    struct { void* p; } d_pl{g_arena.take(sizes[0])}, d_xy{g_arena.take(sizes[1])}, d_e{g_arena.take(sizes[2])}, d_par{g_arena.take(sizes[3])},
        d_nc{g_arena.take(sizes[4])}, d_id{g_arena.take(sizes[5])}, d_fl{g_arena.take(sizes[6])}, d_fin{g_arena.take(sizes[7])},
        d_bp{g_arena.take(sizes[8])}, d_ob{g_arena.take(sizes[9])}, d_w{g_arena.take(sizes[10])}, d_gss{g_arena.take(sizes[11])},
        d_gsi{g_arena.take(sizes[12])}, d_gps{g_arena.take(sizes[13])}, d_gpi{g_arena.take(sizes[14])};
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const auto t1 = now();
    auto t2 = t1, t3 = t1;
    int rc = 0;
    do {
        if ((rc = (int)cudaMemcpyAsync(d_pl.p, pl.data(), (size_t)P * sizeof(rrtfnd_planner_t), cudaMemcpyHostToDevice, st))) break;
        if ((rc = (int)cudaMemcpyAsync(d_ob.p, ob.data(), ob.size() * 8, cudaMemcpyHostToDevice, st))) break;
        if (grid.nx > 0) {
            if ((rc = (int)cudaMemcpyAsync(d_gss.p, g_ss.data(), g_ss.size() * 4, cudaMemcpyHostToDevice, st))) break;
            if ((rc = (int)cudaMemcpyAsync(d_gsi.p, g_si.data(), g_si.size() * 4, cudaMemcpyHostToDevice, st))) break;
            if ((rc = (int)cudaMemcpyAsync(d_gps.p, g_ps.data(), g_ps.size() * 4, cudaMemcpyHostToDevice, st))) break;
            if ((rc = (int)cudaMemcpyAsync(d_gpi.p, g_pi.data(), g_pi.size() * 4, cudaMemcpyHostToDevice, st))) break;
            grid.seg_start = (const int32_t*)d_gss.p; grid.seg_items = (const int32_t*)d_gsi.p;
            grid.pt_start = (const int32_t*)d_gps.p; grid.pt_items = (const int32_t*)d_gpi.p;
        }
        if (prm->stream_mode == RRTFND_STREAM_WORDS) {
            if ((rc = (int)cudaMemcpyAsync(d_w.p, words_host, wb, cudaMemcpyHostToDevice, st))) break;
        }
        rrtfnd_batch_t bt;
        memset(&bt, 0, sizeof bt);
        bt.n_planners = P; bt.planners = (rrtfnd_planner_t*)d_pl.p;
        bt.xy = (double*)d_xy.p; bt.ecost = (double*)d_e.p;
        bt.parent = (int32_t*)d_par.p; bt.nchild = (int32_t*)d_nc.p; bt.id = (int32_t*)d_id.p; bt.nflags = (uint8_t*)d_fl.p;
        bt.grid = grid;
        bt.finish = (int32_t*)d_fin.p; bt.best_path = (int32_t*)d_bp.p; bt.obst = (const double*)d_ob.p;
        bt.words = wb ? (const uint32_t*)d_w.p : nullptr; bt.trace = nullptr;
        rrtfnd_params_t q = *prm;
        q.flags &= ~(RRTFND_FLAG_TRACE | RRTFND_FLAG_OBST_HAS_K | RRTFND_FLAG_RECTANGLES);
        for (size_t j = 0; j < M; ++j) if (ob[4 * j + 2] < 0.0) q.flags |= RRTFND_FLAG_RECTANGLES;
        if ((rc = rrtfnd_init_batch(&q, &bt, st))) break;
        if (timing) { cudaStreamSynchronize(st); t2 = now(); }
        if ((rc = rrtfnd_plan_batch(&q, &bt, st))) break;
        if (timing) { cudaStreamSynchronize(st); t3 = now(); }
        if (planners_host && (rc = (int)cudaMemcpyAsync(planners_host, d_pl.p, (size_t)P * sizeof(rrtfnd_planner_t), cudaMemcpyDeviceToHost, st))) break;
        if (xy_host && (rc = (int)cudaMemcpyAsync(xy_host, d_xy.p, n * 16, cudaMemcpyDeviceToHost, st))) break;
        if (ecost_host && (rc = (int)cudaMemcpyAsync(ecost_host, d_e.p, n * 8, cudaMemcpyDeviceToHost, st))) break;
        if (parent_host && (rc = (int)cudaMemcpyAsync(parent_host, d_par.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (nchild_host && (rc = (int)cudaMemcpyAsync(nchild_host, d_nc.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (id_host && (rc = (int)cudaMemcpyAsync(id_host, d_id.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (best_path_host && (rc = (int)cudaMemcpyAsync(best_path_host, d_bp.p, (size_t)P * prm->path_cap * 4, cudaMemcpyDeviceToHost, st))) break;
        rc = (int)cudaStreamSynchronize(st);
    } while (0);
    cudaStreamDestroy(st);
    if (timing) {
        const auto t4 = now();
        fprintf(stderr, "rrtfnd_plan_host: setup+alloc %.1f ms, h2d+init %.1f ms, plan %.1f ms, d2h %.1f ms\n", ms(t0, t1), ms(t1, t2),
                ms(t2, t3), ms(t3, t4));
    }
    return rc;
}

extern "C" int rrtfnd_host_release(void) {
    std::lock_guard<std::mutex> guard(g_arena.lock);
    g_arena.release();
    return 0;
}
