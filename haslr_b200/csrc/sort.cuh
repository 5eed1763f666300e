// sort.cuh - one warp sorts n distinct 64-bit keys (optionally with a 32-bit payload) that sit in global memory.
//
// The per-edge lists of K2 (supports of a backbone edge, Backbone_graph.cpp:148-171) and K4 (interval ends, Assemble.cpp:196-226)
// are as long as the coverage: tens. They are rank-sorted (n^2 / 32 comparisons per lane, no scratch, no synchronisation). A list
// longer than SORT_RANK_MAX - a contig end with hundreds or thousands of supports - takes a bitonic network instead, in its
// "flip" form where every comparator is ascending (first step of a merge pairs i with its mirror image in the block, the following
// steps pair i with i + stride): positions >= n then behave like +infinity without existing, a comparator whose upper index is
// >= n is simply skipped, so n needs no padding to a power of two. n log^2 n / 64 compare-exchanges per lane.
#pragma once
#include <cstdint>

namespace hgpu {

#ifndef HGPU_SORT_RANK_MAX
#define HGPU_SORT_RANK_MAX 64
#endif
static constexpr uint32_t SORT_RANK_MAX = HGPU_SORT_RANK_MAX;

template <bool PAYLOAD>
__device__ __forceinline__ void warp_bitonic_u64(unsigned long long* key, uint32_t* val, uint32_t n, int lane) {
    uint32_t m = 1;
    while (m < n) m <<= 1;
    auto cmpx = [&](uint32_t lo, uint32_t hi) {
        if (hi >= n) return;
        const unsigned long long a = key[lo], b = key[hi];
        if (a > b) {
            key[lo] = b; key[hi] = a;
            if (PAYLOAD) { const uint32_t t = val[lo]; val[lo] = val[hi]; val[hi] = t; }
        }
    };
    for (uint32_t size = 1; size < m; size <<= 1) {
        __syncwarp();
        for (uint32_t i = lane; i < m / 2; i += 32) {                 // flip: i-th element of the block's lower half with its mirror image
            const uint32_t blk = i / size, r = i % size;
            cmpx(blk * 2 * size + r, blk * 2 * size + 2 * size - 1 - r);
        }
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (uint32_t i = lane; i < m / 2; i += 32) {
                const uint32_t lo = (i / stride) * 2 * stride + (i % stride);
                cmpx(lo, lo + stride);
            }
        }
    }
    __syncwarp();
}

}  // namespace hgpu
