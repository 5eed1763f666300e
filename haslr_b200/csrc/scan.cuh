// Exclusive scan shared by the kernel families (reads / contig ends / edge entries / text chunks / CIGAR runs).
// Small inputs take one block; larger ones three passes (tile sums, scan of the tile sums in 64 bits, tile scans) with
// grids sized by the input, so the 2.5e8-row hit tables of a human dataset do not funnel through one SM. The 64-bit
// total is kept beside the 32-bit offsets: callers refuse inputs whose offsets would wrap instead of corrupting them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

static constexpr unsigned FULLM = 0xFFFFFFFFu;
static constexpr uint32_t SCAN_TILE = 4096;           // elements per block per pass (1024 threads x 4)
static constexpr uint32_t SCAN_ONE_BLOCK_MAX = 32768; // at most this many elements go through the single-block kernel

// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t s, uint32_t* warp_sum /* [32] shared */, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLM, incl, d); if (lane >= d) incl += o; }
    __syncthreads();                                  // warp_sum may still be read by the previous call's consumers
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    uint32_t t = warp_sum[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLM, ti, d); if (lane >= d) ti += o; }
    *total = __shfl_sync(FULLM, ti, 31);
    const uint32_t wex = __shfl_sync(FULLM, ti - t, w);
    return wex + incl - s;
}

// out[i] = sum in[0..i), out[n] = total (low 32 bits), *total64 = total (if not null). Single block.
static __global__ void __launch_bounds__(1024) k_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* total64) {
    __shared__ uint32_t warp_sum[32];
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < n; base += SCAN_TILE) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; s += v[k]; }
        uint32_t tot;
        uint32_t run = (uint32_t)carry + block_excl_scan_1024(s, warp_sum, &tot);
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
        carry += tot;
    }
    if (threadIdx.x == 0) { out[n] = (uint32_t)carry; if (total64) *total64 = carry; }
}

// pass 1: tile_sum[b] = sum of tile b
static __global__ void __launch_bounds__(1024) k_scan_tile_sums(const uint32_t* in, uint32_t n, unsigned long long* tile_sum) {
    __shared__ uint32_t warp_sum[32];
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    uint32_t s = 0;
    if (i0 + 3 < n) { const uint4 v = *reinterpret_cast<const uint4*>(in + i0); s = v.x + v.y + v.z + v.w; }
    else { for (int k = 0; k < 4; ++k) if (i0 + k < n) s += in[i0 + k]; }
    uint32_t tot;
    block_excl_scan_1024(s, warp_sum, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
// pass 2: exclusive scan of the tile sums in 64 bits, in place (one block; n_tiles = n / 4096)
static __global__ void __launch_bounds__(1024) k_scan_tile_offsets(unsigned long long* tile_sum, uint32_t n_tiles, unsigned long long* total64) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sum[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(FULLM, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) ws[w] = incl;
        __syncthreads();
        if (w == 0) {
            const unsigned long long t = ws[lane]; unsigned long long ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(FULLM, ti, d); if (lane >= d) ti += o; }
            ws[lane] = ti - t;
        }
        __syncthreads();
        const unsigned long long ex = carry_s + ws[w] + incl - v;
        if (i < n_tiles) tile_sum[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total64) *total64 = carry_s;
}
// pass 3: scan every tile from its offset
static __global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* in, uint32_t* out, uint32_t n, const unsigned long long* tile_off,
                                                            uint32_t n_tiles) {
    __shared__ uint32_t warp_sum[32];
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    uint32_t v[4] = {0, 0, 0, 0};
    if (i0 + 3 < n) { const uint4 q = *reinterpret_cast<const uint4*>(in + i0); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else { for (int k = 0; k < 4; ++k) if (i0 + k < n) v[k] = in[i0 + k]; }
    uint32_t tot;
    uint32_t run = (uint32_t)tile_off[blockIdx.x] + block_excl_scan_1024(v[0] + v[1] + v[2] + v[3], warp_sum, &tot);
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
    if (blockIdx.x == n_tiles - 1 && threadIdx.x == 1023) out[n] = (uint32_t)tile_off[blockIdx.x] + tot;
}

// Host helper: out[0..n] = exclusive scan of in[0..n) (in and out must be 16-byte aligned, as cudaMalloc gives), 64-bit total
// into *d_total64 (device, may be null). tile_tmp: at least n / 4096 + 1 entries. Returns the number of kernels launched.
static inline int scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* tile_tmp, unsigned long long* d_total64) {
    if (n <= SCAN_ONE_BLOCK_MAX || !tile_tmp) {
        k_exclusive_scan<<<1, 1024, 0, st>>>(in, out, n, d_total64);
        return 1;
    }
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tile_sums<<<n_tiles, 1024, 0, st>>>(in, n, tile_tmp);
    k_scan_tile_offsets<<<1, 1024, 0, st>>>(tile_tmp, n_tiles, d_total64);
    k_scan_tiles<<<n_tiles, 1024, 0, st>>>(in, out, n, tile_tmp, n_tiles);
    return 3;
}
static inline size_t scan_tmp_entries(uint32_t n) { return (size_t)n / SCAN_TILE + 2; }
