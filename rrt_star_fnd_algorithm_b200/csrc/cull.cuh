// cull.cuh — obstacle tests of the planners with grid culling (see grid.cuh for the exactness argument).
//   seg_blocked_lane / seg_blocked_warp   through_obstacle      src/algorithm.py:581-591
//   point_occupied_warp                   Map.is_occupied_c     src/map.py:38-47
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "grid.cuh"

namespace rrtfnd {

constexpr unsigned kFull = 0xffffffffu;

struct Obst {
    const double2* tab;      // [M][2]: (ox, oy), (r, K)
    rrtfnd_obst_grid_t g;
    int M;
    int full_cells;          // bounding boxes covering more cells than this test the whole table instead
    bool use_grid;
};

__device__ __forceinline__ void load_obst(const double2* __restrict__ tab, int j, double& ox, double& oy, double& r,
                                          double& K) {
    const double2 a = __ldg(tab + 2 * j), b = __ldg(tab + 2 * j + 1);
    ox = a.x; oy = a.y; r = b.x; K = b.y;
}

__device__ __forceinline__ bool seg_can_cull(const Obst& ob, const Seg& sg, double ya, double yb, int& cx0, int& cx1,
                                             int& cy0, int& cy1) {
    if (!ob.use_grid || sg.vertical || !(fabs(sg.m) <= ob.g.slope_max)) return false;
    cx0 = grid_cell(sg.lo, ob.g.x0, ob.g.inv_cell, ob.g.nx);
    cx1 = grid_cell(sg.hi, ob.g.x0, ob.g.inv_cell, ob.g.nx);
    cy0 = grid_cell(ya < yb ? ya : yb, ob.g.y0, ob.g.inv_cell, ob.g.ny);
    cy1 = grid_cell(ya < yb ? yb : ya, ob.g.y0, ob.g.inv_cell, ob.g.ny);
    return (cx1 - cx0 + 1) * (cy1 - cy0 + 1) <= ob.full_cells;
}

// through_obstacle for this LANE's own segment (lanes hold different segments; sequential over its circles)
__device__ __forceinline__ bool seg_blocked_lane(const Obst& ob, const Seg& sg, double ya, double yb, long long& tests) {
    int cx0, cx1, cy0, cy1;
    double ox, oy, r, K;
    if (seg_can_cull(ob, sg, ya, yb, cx0, cx1, cy0, cy1)) {
        for (int cy = cy0; cy <= cy1; ++cy) {
            int j = __ldg(ob.g.seg_start + cy * ob.g.nx + cx0);
            const int e = __ldg(ob.g.seg_start + cy * ob.g.nx + cx1 + 1);
            for (; j < e; ++j) {
                load_obst(ob.tab, __ldg(ob.g.seg_items + j), ox, oy, r, K);
                ++tests;
                if (seg_hits_circle(sg, ox, oy, r, K)) return true;
            }
        }
        return false;
    }
    for (int j = 0; j < ob.M; ++j) {
        load_obst(ob.tab, j, ox, oy, r, K);
        ++tests;
        if (seg_hits_circle(sg, ox, oy, r, K)) return true;
    }
    return false;
}

// through_obstacle for ONE segment held by every lane (lanes stride over its circles)
__device__ __forceinline__ bool seg_blocked_warp(const Obst& ob, const Seg& sg, double ya, double yb, int lane,
                                                 long long& tests) {
    int cx0, cx1, cy0, cy1;
    double ox, oy, r, K;
    bool hit = false;
    if (seg_can_cull(ob, sg, ya, yb, cx0, cx1, cy0, cy1)) {
        for (int cy = cy0; cy <= cy1; ++cy) {
            const int b = __ldg(ob.g.seg_start + cy * ob.g.nx + cx0);
            const int e = __ldg(ob.g.seg_start + cy * ob.g.nx + cx1 + 1);
            for (int j = b + lane; j < e; j += 32) {
                load_obst(ob.tab, __ldg(ob.g.seg_items + j), ox, oy, r, K);
                ++tests;
                hit |= seg_hits_circle(sg, ox, oy, r, K);
            }
        }
    } else {
        for (int j = lane; j < ob.M; j += 32) {
            load_obst(ob.tab, j, ox, oy, r, K);
            ++tests;
            hit |= seg_hits_circle(sg, ox, oy, r, K);
        }
    }
    return __any_sync(kFull, hit);
}

// Map.is_occupied_c for ONE point held by every lane
__device__ __forceinline__ bool point_occupied_warp(const Obst& ob, double px, double py, double node_radius, int lane,
                                                    long long& tests) {
    double ox, oy, r, K;
    bool occ = false;
    if (ob.use_grid) {
        const int c = grid_cell(py, ob.g.y0, ob.g.inv_cell, ob.g.ny) * ob.g.nx + grid_cell(px, ob.g.x0, ob.g.inv_cell, ob.g.nx);
        const int b = __ldg(ob.g.pt_start + c), e = __ldg(ob.g.pt_start + c + 1);
        for (int j = b + lane; j < e; j += 32) {
            load_obst(ob.tab, __ldg(ob.g.pt_items + j), ox, oy, r, K);
            ++tests;
            occ |= point_in_inflated(px, py, ox, oy, r, node_radius);
        }
    } else {
        for (int j = lane; j < ob.M; j += 32) {
            load_obst(ob.tab, j, ox, oy, r, K);
            ++tests;
            occ |= point_in_inflated(px, py, ox, oy, r, node_radius);
        }
    }
    return __any_sync(kFull, occ);
}


// Obst for a kernel: obstacle table + grid; bounding boxes covering more cells than `full_cells` are cheaper to
// test against the whole table (cells overlap, so their lists repeat circles).
__device__ __forceinline__ Obst make_obst(const double* obst, int n_obst, const rrtfnd_obst_grid_t& grid) {
    Obst ob;
    ob.tab = reinterpret_cast<const double2*>(obst);
    ob.g = grid;
    ob.M = n_obst;
    ob.use_grid = grid.nx > 0 && grid.seg_start != nullptr && n_obst > 0;
    ob.full_cells = 0;
    if (ob.use_grid) {
        const int ncells = grid.nx * grid.ny;
        const int n_items = __ldg(grid.seg_start + ncells);
        ob.full_cells = (int)(((long long)n_obst * ncells) / (n_items > 0 ? n_items : 1));
        if (ob.full_cells < 1) ob.full_cells = 1;
    }
    return ob;
}

}  // namespace rrtfnd
