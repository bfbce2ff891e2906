// fp64_rate.cu — measured FP64 issue-rate peaks of this GPU: the roofline denominators for the collision arithmetic.
//
// the planners are compiled with --fmad=false (every fp64 operation of the reference is rounded separately),
// so the useful ceiling of the circle tests is the DADD / DMUL issue rate, not the DFMA flop figure; MEASURED_PEAKS.json has no
// fp64 row. This stand-alone program measures, with CUDA events, per-SM-clock and aggregate rates of
//   DADD, DMUL, DFMA   — 8 independent dependent chains per thread (enough ILP to hide the pipe latency), 4 x 148 x 8 warps;
//   __ddiv_rn, __dsqrt_rn — the IEEE sequences the circle test calls twice / once per circle (their cost in DADD-equivalents).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_rate fp64_rate.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + (threadIdx.x + 1) * 1e-3 + k;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) a[k] = __dadd_rn(a[k], c);
            if (OP == 1) a[k] = __dmul_rn(a[k], m);
            if (OP == 2) a[k] = __fma_rn(a[k], m, c);
            if (OP == 3) a[k] = __ddiv_rn(a[k] + 2.0, m);
            if (OP == 4) a[k] = __dsqrt_rn(a[k] + 2.0);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keep the chains alive
}

template <int OP>
static double run(int sms, int iters, double* d_out) {
    const int blocks = sms * 8;                       // 8 CTAs of 8 warps: 64 warps per SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<OP><<<blocks, 256>>>(d_out, iters / 8, 1.0);            // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        rate_kernel<OP><<<blocks, 256>>>(d_out, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return (double)blocks * 256.0 * 8.0 * iters / (best * 1e-3);          // operations per second
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double* d_out; cudaMalloc(&d_out, sizeof(double) * sms * 8 * 256);
    const double add = run<0>(sms, 4096, d_out), mul = run<1>(sms, 4096, d_out), fma_ = run<2>(sms, 4096, d_out);
    const double dv = run<3>(sms, 256, d_out), sq = run<4>(sms, 256, d_out);
    const double clk = khz * 1e3;
    printf("{\"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"dadd_per_s\": %.4e, \"dmul_per_s\": %.4e, \"dfma_per_s\": %.4e, "
           "\"ddiv_rn_per_s\": %.4e, \"dsqrt_rn_per_s\": %.4e, \"dadd_per_clk_per_sm\": %.2f, \"dmul_per_clk_per_sm\": %.2f, "
           "\"dfma_per_clk_per_sm\": %.2f, \"ddiv_in_dadd_equivalents\": %.1f, \"dsqrt_in_dadd_equivalents\": %.1f, "
           "\"note\": \"thread-level operations per second, best of 5, CUDA events; per-clock figures use the maximum SM clock\"}\n",
           sms, clk / 1e6, add, mul, fma_, dv, sq, add / clk / sms, mul / clk / sms, fma_ / clk / sms, add / dv, add / sq);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : 1;
}
