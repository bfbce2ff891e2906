// stream.cuh — per-warp streaming of a planner's node coordinates through shared memory with TMA bulk copies.
//
// Every nearest / near-set scan reads xy[0 .. n_slots) of ONE planner front to back (16 B per node). A warp that
// fetches its tiles with ordinary loads has at most a few hundred bytes per lane in flight and pays the L2 round trip
// (~250 cycles) once per tile. Here lane 0 issues `cp.async.bulk` (the 1-D TMA copy) for whole 2 KB tiles of 128
// nodes into a ring of kStages shared-memory buffers and the warp consumes them behind an mbarrier per stage: up to
// kStages tiles are in flight per warp, no registers are tied up by loads in flight, and the scan body reads
// shared memory.
//
// The ring is private to the warp (warps never synchronise with each other), so the mbarriers have an arrival count
// of 1: lane 0 arrives with the expected byte count, the copy engine completes the transaction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrtfnd {

constexpr int kTileNodes = 128;                 // nodes per tile: 4 per lane
constexpr int kTileBytes = kTileNodes * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct XYStream {
    uint32_t ring;      // shared-window address of the ring (n_stages * kTileBytes, 128-byte aligned)
    uint32_t bars;      // shared-window address of n_stages mbarriers
    uint32_t parity;    // bit s = phase parity the warp waits for next on stage s
    int n_stages;       // power of two
};

__device__ __forceinline__ void stream_init(XYStream& st, void* ring, void* bars, int n_stages, int lane) {
    st.ring = smem_u32(ring); st.bars = smem_u32(bars); st.parity = 0; st.n_stages = n_stages;
    if (lane == 0) {
        for (int s = 0; s < n_stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st.bars + 8 * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}

// lane 0 only: request tile `t` (nodes t*128 ..) of xy[0 .. n_slots) into its stage
__device__ __forceinline__ void stream_issue(const XYStream& st, const double2* xy, int n_slots, int t) {
    const int first = t * kTileNodes;
    const int n = min(kTileNodes, n_slots - first);
    const uint32_t s = (uint32_t)t & (uint32_t)(st.n_stages - 1);
    const uint32_t bar = st.bars + 8 * s, dst = st.ring + s * kTileBytes, bytes = (uint32_t)n * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(xy + first), "r"(bytes), "r"(bar) : "memory");
}

// Start a scan: order earlier generic-proxy stores to xy (inserts, tombstones, compaction) before the copy engine's
// reads, then put the first n_stages tiles in flight.
static __device__ __noinline__ void stream_begin_impl(uint32_t ring, uint32_t bars, int n_stages, const double2* xy, int n_slots) {
    XYStream st; st.ring = ring; st.bars = bars; st.parity = 0; st.n_stages = n_stages;
    asm volatile("fence.proxy.async;" ::: "memory");
    const int n_tiles = (n_slots + kTileNodes - 1) / kTileNodes;
#pragma unroll 1
    for (int t = 0; t < n_stages && t < n_tiles; ++t) stream_issue(st, xy, n_slots, t);
}
__device__ __forceinline__ void stream_begin(const XYStream& st, const double2* xy, int n_slots, int lane) {
    __syncwarp();
    if (lane == 0 && st.n_stages > 0) stream_begin_impl(st.ring, st.bars, st.n_stages, xy, n_slots);
}

// Wait for tile t; returns its shared-memory address (generic pointer to 128 double2, the tail of the last tile is stale)
__device__ __forceinline__ const double2* stream_wait(XYStream& st, int t) {
    const uint32_t s = (uint32_t)t & (uint32_t)(st.n_stages - 1);
    const uint32_t bar = st.bars + 8 * s, par = (st.parity >> s) & 1u;
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(par) : "memory");
    } while (!done);
    st.parity ^= 1u << s;
    return reinterpret_cast<const double2*>(__cvta_shared_to_generic((size_t)(st.ring + s * kTileBytes)));
}

// The warp is done with tile t: refill its stage with tile t + n_stages (if the scan has one)
__device__ __forceinline__ void stream_release(const XYStream& st, const double2* xy, int n_slots, int t, int lane) {
    __syncwarp();                                   // every lane has read the stage
    const int nt = t + st.n_stages;
    if (lane == 0 && nt * kTileNodes < n_slots) stream_issue(st, xy, n_slots, nt);
}

}  // namespace rrtfnd
