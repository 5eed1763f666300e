// (iv) edge coordinates: which stretch of every supporting long read spans the gap of a backbone edge — kernel + C ABI.
//
// hgpu_edge_coords replaces asm_calc_single_edge_coordinates and the pthread edge queue of asm_calc_edge_coordinates_MT
// around it (reference src/haslr_assemble/src/Assemble.cpp:24-155,157-363,436-477). One warp per edge:
//   1. lanes load the edge's supports and their head / tail compact-read elements (coalesced over the support list) and
//      lay down four key lists (begin / end on each anchor contig), key = position << 32 | support index;
//   2. each list is rank-sorted by the warp (keys are distinct; supports per edge = coverage, tens);
//   3. one lane runs the two interval sweeps of coords_core.cuh on the sorted lists (bitmask instead of std::set copies);
//   4. lanes intersect the two masks and walk the run-length CIGAR windows of "their" supports (one support per lane).
// Integer, latency-bound work over small per-edge lists: the grid is sized to keep every SM's warp slots full.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "coords_core.cuh"
#include "sort.cuh"

using namespace hgpu;

struct CoordState {
    DevBuf<uint8_t> edge_rev, hit_is_rev;
    DevBuf<uint32_t> supp_off, cl_read_off, read_len, cg_off, cg_ops, mask;
    DevBuf<hgpu_edge_supp> supp;
    DevBuf<hgpu_cl_elem> elems;
    DevBuf<uint64_t> keys;
    DevBuf<hgpu_edge_coord> out_edge;
    DevBuf<hgpu_supp_coord> out_supp;
    DevBuf<unsigned long long> counters;     // [0] CIGAR runs in the windows walked, [1] first invalid support (min), [2] first invalid element (min)
};
void coord_state_destroy(CoordState* s) { delete s; }

static constexpr unsigned FULLM = 0xFFFFFFFFu;

// keys: 8 lists of n_supp entries (4 unsorted, 4 sorted), edge e uses [supp_off[e], supp_off[e+1]) of each;
// mask: 2 bitmasks per edge, words [2 * (supp_off[e] / 32 + e) ...) — ceil(n/32) words each, zeroed by the host
__global__ void __launch_bounds__(128) k4_edge_coords(CoordIn in, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off,
                                                      uint32_t n_supp, uint32_t n_reads, uint32_t n_hits, uint64_t* keys, uint32_t* mask,
                                                      hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp, unsigned long long* counters) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    unsigned long long runs = 0;
    for (uint32_t e = gw; e < n_edges; e += nw) {
        const uint32_t b = supp_off[e];
        if (supp_off[e + 1] < b || supp_off[e + 1] > n_supp) { if (lane == 0) atomicMin(counters + 1, (unsigned long long)b); continue; }
        const uint32_t n = supp_off[e + 1] - b;
        // the reference indexes its containers unchecked; refuse (and skip the edge) instead of reading out of bounds
        bool bad = false;
        for (uint32_t k = lane; k < n; k += 32) {
            const hgpu_edge_supp sp = in.supp[b + k];
            const uint32_t rid = sp.lr_id_strand & 0x7FFFFFFFu;
            bool sbad = rid >= n_reads;
            if (!sbad) {
                const uint32_t cnt = in.cl_read_off[rid + 1] - in.cl_read_off[rid];
                sbad = sp.cmp_head >= cnt || sp.cmp_tail >= cnt;
            }
            if (sbad) { atomicMin(counters + 1, (unsigned long long)(b + k)); bad = true; continue; }
            for (int side = 0; side < 2; ++side) {
                const uint32_t j = in.cl_read_off[rid] + (side ? sp.cmp_tail : sp.cmp_head);
                const hgpu_cl_elem& el = in.elems[j];
                bool ebad = el.hit >= n_hits;
                if (!ebad) {
                    const uint32_t w = in.cg_off[el.hit + 1] - in.cg_off[el.hit];
                    ebad = w && (el.cg_lo > el.cg_hi || el.cg_hi >= w);
                }
                if (ebad) { atomicMin(counters + 2, (unsigned long long)j); bad = true; }
            }
        }
        if (__any_sync(FULLM, bad)) continue;
        const uint32_t rev1 = edge_rev[e] & 1u, rev2 = (edge_rev[e] >> 1) & 1u;
        const hgpu_edge_supp* es = in.supp + b;
        uint64_t* const raw = keys + b;                             // list l: raw + l * n_supp (unsorted), srt + l * n_supp (sorted)
        uint64_t* const srt = keys + (size_t)4 * n_supp + b;
        const uint32_t mw = (n + 31) / 32;
        uint32_t* m1 = mask + 2 * ((size_t)(b / 32) + e);
        uint32_t* m2 = m1 + mw;
        // 1. key lists: begin / end of the head element on contig1, of the tail element on contig2 (Assemble.cpp:196-226)
        for (uint32_t k = lane; k < n; k += 32) {
            const hgpu_cl_elem& h = k4_elem(in, es[k], true);
            const hgpu_cl_elem& t = k4_elem(in, es[k], false);
            raw[k] = k4_key(h.t_start, k); raw[(size_t)n_supp + k] = k4_key(h.t_end, k);
            raw[(size_t)2 * n_supp + k] = k4_key(t.t_start, k); raw[(size_t)3 * n_supp + k] = k4_key(t.t_end, k);
        }
        __syncwarp();
        // 2. rank sort (keys are distinct: the support index is the low word)
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const uint64_t* r = raw + (size_t)l * n_supp;
            if (n <= SORT_RANK_MAX) {
                for (uint32_t k = lane; k < n; k += 32) {
                    const uint64_t key = r[k];
                    uint32_t rank = 0;
                    for (uint32_t q = 0; q < n; ++q) rank += r[q] < key ? 1u : 0u;
                    srt[(size_t)l * n_supp + rank] = key;
                }
            } else {                                              // many supports: bitonic network in place (sort.cuh)
                uint64_t* d = srt + (size_t)l * n_supp;
                for (uint32_t k = lane; k < n; k += 32) d[k] = r[k];
                warp_bitonic_u64<false>(reinterpret_cast<unsigned long long*>(d), nullptr, n, lane);
            }
        }
        __syncwarp();
        // 3. the two sweeps
        uint32_t i1lo = 0, i1hi = 0, i2lo = 0, i2hi = 0;
        if (lane == 0) {
            k4_best_interval(srt, srt + n_supp, n, true, m1, &i1lo, &i1hi);
            k4_best_interval(srt + (size_t)2 * n_supp, srt + (size_t)3 * n_supp, n, false, m2, &i2lo, &i2hi);
        }
        __syncwarp();
        i1lo = __shfl_sync(FULLM, i1lo, 0); i1hi = __shfl_sync(FULLM, i1hi, 0);
        i2lo = __shfl_sync(FULLM, i2lo, 0); i2hi = __shfl_sync(FULLM, i2hi, 0);
        const uint32_t c1 = rev1 == 0 ? i1hi - 1 : i1lo;          // last shared base on the head contig (Assemble.cpp:228-238)
        const uint32_t c2 = rev2 == 0 ? i2lo : i2hi - 1;          // first shared base on the tail contig
        // 4. members of both best sets: one support per lane
        uint32_t n_best = 0, n_cns = 0;
        for (uint32_t k0 = 0; k0 < n; k0 += 32) {
            const uint32_t k = k0 + lane;
            bool best = false, ok = false;
            if (k < n) {
                best = ((m1[k >> 5] & m2[k >> 5]) >> (k & 31u)) & 1u;
                hgpu_supp_coord o;
                o.lr_start = -1; o.lr_end = -1; o.lr_strand = 0; o.in_best = 0;
                if (best) {
                    k4_walk(in, es[k], rev1, rev2, c1, c2, &o); ok = o.lr_start != -1 && o.lr_end != -1;
                    const hgpu_cl_elem& h = k4_elem(in, es[k], true);
                    const hgpu_cl_elem& t = k4_elem(in, es[k], false);
                    runs += (h.cg_hi - h.cg_lo + 1) + (t.cg_hi - t.cg_lo + 1);
                }
                out_supp[b + k] = o;
            }
            n_best += __popc(__ballot_sync(FULLM, best));
            n_cns += __popc(__ballot_sync(FULLM, ok));
        }
        if (lane == 0) {
            hgpu_edge_coord oe;
            oe.int1_lo = i1lo; oe.int1_hi = i1hi; oe.int2_lo = i2lo; oe.int2_hi = i2hi; oe.c1 = c1; oe.c2 = c2; oe.n_best = n_best; oe.n_cns = n_cns;
            out_edge[e] = oe;
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) runs += __shfl_xor_sync(FULLM, runs, d);
    if (lane == 0 && runs) atomicAdd(counters, runs);
}

// K4 on device-resident compact reads and hit table; edges / supports / read lengths come from the host (they are made by
// the host's graph cleaning)
static int edge_coords_run(hgpu_t* ctx, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                           const hgpu_cl_elem* d_elems, const uint32_t* d_cl_read_off, uint32_t n_reads, const uint32_t* read_len,
                           const uint8_t* d_hit_is_rev, const uint32_t* d_cg_off, const uint32_t* d_cg_ops, uint32_t n_hits,
                           hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp) {
    CoordState* S = ctx->coords;
    cudaStream_t st = ctx->stream;
    const uint32_t n_supp = supp_off[n_edges];
    ctx->stage.ms_k4 = 0; ctx->stage.launches_k4 = 0; ctx->stage.k4_edges = n_edges; ctx->stage.k4_supports = n_supp; ctx->stage.k4_runs = 0;
    const size_t mask_words = 2 * ((size_t)n_supp / 32 + n_edges) + 2;
    HGPU_CUDA(ctx, S->edge_rev.ensure(n_edges)); HGPU_CUDA(ctx, S->supp_off.ensure(n_edges + 1)); HGPU_CUDA(ctx, S->supp.ensure(n_supp));
    HGPU_CUDA(ctx, S->read_len.ensure(n_reads));
    HGPU_CUDA(ctx, S->keys.ensure(8 * (size_t)n_supp)); HGPU_CUDA(ctx, S->mask.ensure(mask_words));
    HGPU_CUDA(ctx, S->out_edge.ensure(n_edges)); HGPU_CUDA(ctx, S->out_supp.ensure(n_supp)); HGPU_CUDA(ctx, S->counters.ensure(4));
    HGPU_H2D(ctx, S->edge_rev.p, edge_rev, n_edges);
    HGPU_H2D(ctx, S->supp_off.p, supp_off, (size_t)(n_edges + 1) * 4);
    HGPU_H2D(ctx, S->supp.p, supp, (size_t)n_supp * sizeof(hgpu_edge_supp));
    HGPU_H2D(ctx, S->read_len.p, read_len, (size_t)n_reads * 4);
    HGPU_CUDA(ctx, cudaMemsetAsync(S->mask.p, 0, mask_words * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->out_edge.p, 0, (size_t)n_edges * sizeof(hgpu_edge_coord), st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->out_supp.p, 0, (size_t)n_supp * sizeof(hgpu_supp_coord), st));
    const unsigned long long init[3] = {0ull, ~0ull, ~0ull};
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->counters.p, init, sizeof init, cudaMemcpyHostToDevice, st));

    CoordIn in{S->supp.p, d_elems, d_cl_read_off, S->read_len.p, d_hit_is_rev, d_cg_off, d_cg_ops};
    const uint32_t blocks = std::min<uint32_t>((n_edges + 3) / 4, (uint32_t)ctx->sm_count * 16);     // 16 blocks of 4 warps = every warp slot of an SM
    stage_begin(ctx, ctx->ev_k4);
    k4_edge_coords<<<blocks, 128, 0, st>>>(in, n_edges, S->edge_rev.p, S->supp_off.p, n_supp, n_reads, n_hits, S->keys.p, S->mask.p, S->out_edge.p,
                                           S->out_supp.p, S->counters.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    stage_end(ctx, ctx->ev_k4);
    ctx->launches++; ctx->stage.launches_k4 = 1;
    unsigned long long cnt[3];
    HGPU_CUDA(ctx, cudaMemcpyAsync(cnt, S->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
    HGPU_D2H(ctx, out_edge, S->out_edge.p, (size_t)n_edges * sizeof(hgpu_edge_coord));
    HGPU_D2H(ctx, out_supp, S->out_supp.p, (size_t)n_supp * sizeof(hgpu_supp_coord));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stage.ms_k4 = stage_ms(ctx, ctx->ev_k4);
    ctx->stage.k4_runs = cnt[0];
    if (cnt[1] != ~0ull) HGPU_FAIL(ctx, HGPU_E_INVALID, "support %llu names a read or a compact-read element that does not exist (or supp_off is not monotone there)", cnt[1]);
    if (cnt[2] != ~0ull) HGPU_FAIL(ctx, HGPU_E_INVALID, "compact-read element %llu names a hit or a CIGAR window that does not exist", cnt[2]);
    return HGPU_OK;
}

extern "C" int hgpu_edge_coords(hgpu_t* ctx, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                                const hgpu_cl_elem* elems, const uint32_t* cl_read_off, uint32_t n_reads, const uint32_t* read_len,
                                const uint8_t* hit_is_rev, const uint32_t* cg_off, const uint32_t* cg_ops, uint32_t n_hits,
                                hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp) {
    if (!ctx) return HGPU_E_INVALID;
    if (n_edges == 0) return HGPU_OK;
    if (!edge_rev || !supp_off || !cl_read_off || !read_len || !hit_is_rev || !cg_off || !out_edge) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    const uint32_t n_supp = supp_off[n_edges];
    const uint32_t n_elems = cl_read_off[n_reads];
    if (n_supp && (!supp || !elems || !out_supp)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    const uint32_t n_ops = n_hits ? cg_off[n_hits] : 0;
    if (n_ops && !cg_ops) HGPU_FAIL(ctx, HGPU_E_INVALID, "null cg_ops");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->coords) ctx->coords = new CoordState();
    CoordState* S = ctx->coords;
    HGPU_CUDA(ctx, S->elems.ensure(n_elems)); HGPU_CUDA(ctx, S->cl_read_off.ensure(n_reads + 1));
    HGPU_CUDA(ctx, S->hit_is_rev.ensure(n_hits)); HGPU_CUDA(ctx, S->cg_off.ensure(n_hits + 1)); HGPU_CUDA(ctx, S->cg_ops.ensure(n_ops));
    HGPU_H2D(ctx, S->elems.p, elems, (size_t)n_elems * sizeof(hgpu_cl_elem));
    HGPU_H2D(ctx, S->cl_read_off.p, cl_read_off, (size_t)(n_reads + 1) * 4);
    HGPU_H2D(ctx, S->hit_is_rev.p, hit_is_rev, n_hits);
    HGPU_H2D(ctx, S->cg_off.p, cg_off, (size_t)(n_hits + 1) * 4);
    HGPU_H2D(ctx, S->cg_ops.p, cg_ops, (size_t)n_ops * 4);
    return edge_coords_run(ctx, n_edges, edge_rev, supp_off, supp, S->elems.p, S->cl_read_off.p, n_reads, read_len, S->hit_is_rev.p, S->cg_off.p,
                           S->cg_ops.p, n_hits, out_edge, out_supp);
}

extern "C" int hgpu_edge_coords_dev(hgpu_t* ctx, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                                    const uint32_t* read_len, uint32_t n_reads, hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp) {
    if (!ctx) return HGPU_E_INVALID;
    if (n_edges == 0) return HGPU_OK;
    if (!edge_rev || !supp_off || !read_len || !out_edge) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (!ctx->hits.valid || !ctx->compact.valid || ctx->compact.n_reads != n_reads)
        HGPU_FAIL(ctx, HGPU_E_INVALID, "hgpu_edge_coords_dev needs the hit table and the compact reads of %u reads resident on the device", n_reads);
    if (supp_off[n_edges] && (!supp || !out_supp)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->coords) ctx->coords = new CoordState();
    return edge_coords_run(ctx, n_edges, edge_rev, supp_off, supp, ctx->compact.elems, ctx->compact.read_off, n_reads, read_len, ctx->hits.is_rev,
                           ctx->hits.cg_off, ctx->hits.cg_ops, ctx->hits.n_hits, out_edge, out_supp);
}
