// grid.cuh — uniform grid over the circular obstacles: exact culling of circle tests.
//
// The reference tests every segment against every obstacle (through_obstacle, src/algorithm.py:581-591)
// and every sample against every obstacle (Map.is_occupied_c, src/map.py:38-47). Both results are ORs, so
// skipping circles whose test is provably False leaves every boolean unchanged. DESIGN.md ("Obstacle
// grid") derives the bound used here: for a non-vertical segment with |Line.dir| <= slope_max and all
// coordinates / radii bounded by coord_max, intersection_circle can only return True if the circle's
// centre lies within r + margin of the segment's bounding box. A circle is therefore listed in every cell
// its box inflated by r + margin overlaps, and a segment looks only at the cells its own bounding box
// overlaps. Everything else (vertical segments, steep slopes, NaN) takes the full table.
//
// The cell function is the same monotone expression on the host (builder) and the device (queries), which
// is all the overlap argument needs.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

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

}  // namespace rrtfnd
