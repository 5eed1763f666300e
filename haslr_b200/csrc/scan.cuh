// Single-block exclusive scan shared by the kernel families (sizes here are reads / contig ends / edge entries / text
// chunks: up to a few million counters).
#pragma once
#include <stdint.h>

static constexpr unsigned FULLM = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------------------
// exclusive scan of n uint32 (single block; sizes here are reads / contig ends / edge entries)
// out[i] = sum in[0..i), out[n] = total
// ---------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(1024) k_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024 * 4) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; s += v[k]; }
        uint32_t incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(FULLM, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) warp_sum[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t t = warp_sum[lane], ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(FULLM, ti, d); if (lane >= d) ti += o; }
            warp_sum[lane] = ti - t;   // exclusive
        }
        __syncthreads();
        uint32_t run = carry_s + warp_sum[w] + incl - s;
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}
