// grid.cuh — uniform grid over the circular obstacles: exact culling of circle tests.
//
// The reference tests every segment against every obstacle (through_obstacle, src/algorithm.py:581-591)
// and every sample against every obstacle (Map.is_occupied_c, src/map.py:38-47). Both results are ORs, so
// skipping circles whose test is provably False leaves every boolean unchanged. DESIGN.md ("Obstacle
// grid") derives the bound used here: for a non-vertical segment with |Line.dir| <= slope_max and all
// coordinates / radii bounded by coord_max, intersection_circle can only return True if the circle's
// centre lies within r + margin of the segment's bounding box. A circle is therefore listed in every cell
// its box inflated by r + margin overlaps, and a segment looks only at the cells its own bounding box
// overlaps. Everything else (vertical segments, slopes beyond grid_slope_max(), NaN) takes the full table.
//
// The cell function is the same monotone expression on the host (builder) and the device (queries), which
// is all the overlap argument needs.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rrtfnd.h"

namespace rrtfnd {

__host__ __device__ __forceinline__ int grid_cell(double v, double v0, double inv_cell, int n) {
#ifdef __CUDA_ARCH__
    const double t = floor(__dmul_rn(__dsub_rn(v, v0), inv_cell));
#else
    volatile double d = v - v0;      // one rounding per operation on the host as well (-ffp-contract=off)
    volatile double m = d * inv_cell;
    const double t = floor(m);
#endif
    // NaN and anything below the grid fall into cell 0, anything above into the last cell (monotone clamp)
    if (!(t >= 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)t;
}

// margin(coord_max, slope_max): see DESIGN.md. 2^-26 > sqrt(2^-53); the factor 128 is 4x the derived bound.
__host__ __device__ inline double grid_seg_margin(double coord_max, double slope_max) {
    return 128.0 * (1.0 + slope_max) * coord_max * (1.0 / 67108864.0);
}

// Slope bound of the culled segments (both grid builders). Steeper segments test the WHOLE table, one lane, circle after circle,
// while the other 31 lanes of the chunk wait, so the bound trades that tail (probability ~ 4 atan(1 / slope_max) / 2 pi per
// segment: 0.25 % at 256, 0.016 % at 4096) against the margin, which grows with it and lists every circle in more cells.
// Measured on B200 (profiles/r02_s18_raw.txt; 256 / 1024 / 2048 / 4096): config 3 39.0 / 41.1 / 41.1 / 40.8 M ext/s, its
// 2,048-planner shard 25.1 / 28.0 / 28.1 / 28.4 M, config 2 63.7 / 65.4 / 65.2 / 64.8 M. RRTFND_SLOPE_MAX overrides the default
// for A/B runs; any value keeps the culling exact (the margin follows it).
#ifndef RRTFND_SLOPE_MAX_DEFAULT
#define RRTFND_SLOPE_MAX_DEFAULT 1024.0
#endif
inline double grid_slope_max() {
    static const double v = [] {
        const char* e = getenv("RRTFND_SLOPE_MAX");
        const double x = e ? atof(e) : 0.0;
        return (x >= 1.0 && x <= 1048576.0) ? x : (double)RRTFND_SLOPE_MAX_DEFAULT;
    }();
    return v;
}

}  // namespace rrtfnd
