// cull.cuh — obstacle tests of the planners with grid culling (see grid.cuh for the exactness argument).
//   seg_blocked            through_obstacle      src/algorithm.py:581-591
//   point_occupied_warp    Map.is_occupied_c     src/map.py:38-47
//
// Code size matters here: the fused planner keeps dozens of warps in different phases on one SM, so its
// instruction footprint has to stay inside the instruction cache. The segment test (one division for the line; per
// circle the quadratic's discriminant, the root filter and - rarely - a square root and two divisions, arith.cuh)
// therefore exists ONCE, not inlined, and serves both uses through a
// (first, stride) pair: a lane testing its own segment walks its circles with (0, 1), a warp sharing one segment
// splits them with (lane, 32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtfnd.h"
#include "arith.cuh"
#include "grid.cuh"

namespace rrtfnd {

constexpr unsigned kFull = 0xffffffffu;

// What a kernel needs to test obstacles: the table and the grid (pointer to the kernel's __grid_constant__ parameter
// or to any other copy); full_cells > 0 says that the grid is usable.
struct Obst {
    const double2* tab;               // [M][2]: (ox, oy), (r, K)
    const rrtfnd_obst_grid_t* g;
    int M;
    int full_cells;                   // 0 = no grid
};

__device__ __forceinline__ Obst make_obst(const double* obst, int n_obst, const rrtfnd_obst_grid_t* grid) {
    Obst ob;
    ob.tab = reinterpret_cast<const double2*>(obst);
    ob.g = grid;
    ob.M = n_obst;
    ob.full_cells = 0;
    if (grid->nx > 0 && grid->seg_start != nullptr && n_obst > 0) {
        const int ncells = grid->nx * grid->ny;
        const int n_items = __ldg(grid->seg_start + ncells);
        ob.full_cells = (int)(((long long)n_obst * ncells) / (n_items > 0 ? n_items : 1));
        if (ob.full_cells < 1) ob.full_cells = 1;
    }
    return ob;
}

__device__ __forceinline__ void load_obst(const double2* __restrict__ tab, int j, double& ox, double& oy, double& r,
                                          double& K) {
    const double2 a = __ldg(tab + 2 * j), b = __ldg(tab + 2 * j + 1);
    ox = a.x; oy = a.y; r = b.x; K = b.y;
}

// through_obstacle(Line(p1, p2), obstacles) over the circles first, first + stride, ... of the candidate list.
// Returns (circle tests performed << 1) | hit. Stops at the first hit.
// XORD: the circles of a grid row are tested from p1's end as well (a row lists its cells in ascending x). Measured: +9.5 % on a
// single-wave batch, +6.6 % on c2, -1.8 % on c3's 16,384 planners (the extra index arithmetic), and -8.7 % when one kernel holds
// both forms (a second copy of this function does not fit the instruction cache): a kernel instance uses ONE form throughout.
template <bool RECT = true, bool XORD = true>
static __device__ __noinline__ int seg_blocked(const double2* __restrict__ tab, const rrtfnd_obst_grid_t* __restrict__ g, int M,
                                        int full_cells, double p1x, double p1y, double p2x, double p2y, int first,
                                        int stride) {
    const Seg sg = make_seg(p1x, p1y, p2x, p2y);
    int cy0 = 0, cy1 = 0, nx = 0;
    double x0 = 0, y0 = 0, inv = 0, cell = 0, slack = 0, dxdy = 0;
    const double ylo = p1y < p2y ? p1y : p2y, yhi = p1y < p2y ? p2y : p1y;
    bool cull = false, flat = true;
    if (full_cells > 0 && !sg.vertical && fabs(sg.m) <= g->slope_max) {
        x0 = g->x0; y0 = g->y0; inv = g->inv_cell;
        nx = g->nx;
        cy0 = grid_cell(ylo, y0, inv, g->ny);
        cy1 = grid_cell(yhi, y0, inv, g->ny);
        // A circle that can hit lies within r + margin of the SEGMENT (not merely of its bounding box), so its
        // inflated box contains the segment's closest point and the cell of that point lists it: only cells the
        // segment passes through need to be visited. Per row of cells the segment's x-extent is computed in ordinary
        // floating point and widened by `slack`, far more than its rounding error (a few ulps of the coordinates).
        cell = 1.0 / inv;
        slack = g->coord_max * 1e-9;
        dxdy = (p2x - p1x) / (p2y - p1y);
        flat = !(fabs(dxdy) < 1e300);                       // (nearly) horizontal: every row gets the whole x-range
        cull = true;
    }
    if (!cull) { cy0 = 0; cy1 = 0; }                       // one pseudo-row holding the whole table
    const int32_t* __restrict__ start = g->seg_start;
    const int32_t* __restrict__ items = g->seg_items;
    int tests = 0;
    // rows are visited from p1's end of the segment: the nearest-visible search puts the SAMPLE there (Line(q_rand, node)), and a
    // sample that no node can see - most samples while a tree is young - is boxed in by circles next to it, so its segments end
    // at their first rows instead of after half of them (the OR over circles does not care about the order)
    const int n_rows = cy1 - cy0, dcy = p1y <= p2y ? 1 : -1;
    int cy = p1y <= p2y ? cy0 : cy1;
    for (int row = 0; row <= n_rows; ++row, cy += dcy) {
        int j = 0, e = M;
        if (cull) {
            double xa = sg.lo, xb = sg.hi;
            if (cy1 > cy0 && !flat) {                       // the part of the segment inside this row's y-interval
                const double ra = fmax(y0 + cy * cell - slack, ylo), rb = fmin(y0 + (cy + 1) * cell + slack, yhi);
                const double xs = p1x + (ra - p1y) * dxdy, xe = p1x + (rb - p1y) * dxdy;
                xa = fmax(sg.lo, fmin(xs, xe) - slack);
                xb = fmin(sg.hi, fmax(xs, xe) + slack);
            }
            const int cx0 = grid_cell(xa, x0, inv, nx), cx1 = grid_cell(xb, x0, inv, nx);
            j = __ldg(start + cy * nx + cx0);
            e = __ldg(start + cy * nx + cx1 + 1);
        }
        if (XORD) {   // ... and the circles of a row from p1's end as well (the row's cells are listed in ascending x)
            const int n_row = e - j, base = p1x <= p2x ? j : e - 1, dj = p1x <= p2x ? 1 : -1;
            for (int k = first; k < n_row; k += stride) {
                double ox, oy, r, K;
                load_obst(tab, cull ? __ldg(items + base + dj * k) : base + dj * k, ox, oy, r, K);
                ++tests;
                if (seg_hits_circle<RECT>(sg, ox, oy, r, K)) return (tests << 1) | 1;
            }
        } else {
            for (j += first; j < e; j += stride) {
                double ox, oy, r, K;
                load_obst(tab, cull ? __ldg(items + j) : j, ox, oy, r, K);
                ++tests;
                if (seg_hits_circle<RECT>(sg, ox, oy, r, K)) return (tests << 1) | 1;
            }
        }
    }
    return tests << 1;
}

// Map.is_occupied_c for ONE point held by every lane of the warp
template <bool RECT = true>
__device__ __forceinline__ bool point_occupied_warp(const Obst& ob, double px, double py, double node_radius, int lane,
                                                    long long& tests) {
    const bool cull = ob.full_cells > 0;
    int j = 0, e = ob.M;
    if (cull) {
        const rrtfnd_obst_grid_t* g = ob.g;
        const int c = grid_cell(py, g->y0, g->inv_cell, g->ny) * g->nx + grid_cell(px, g->x0, g->inv_cell, g->nx);
        j = __ldg(g->pt_start + c);
        e = __ldg(g->pt_start + c + 1);
    }
    bool occ = false;
    for (j += lane; j < e; j += 32) {
        double ox, oy, r, K;
        load_obst(ob.tab, cull ? __ldg(ob.g->pt_items + j) : j, ox, oy, r, K);
        ++tests;
        occ |= point_in_inflated<RECT>(px, py, ox, oy, r, node_radius, K);
    }
    return __any_sync(kFull, occ);
}

}  // namespace rrtfnd
