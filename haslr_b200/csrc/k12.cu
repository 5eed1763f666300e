// (i) PAF hits -> compact long reads and (ii) compact long reads -> backbone edge table: kernels + C ABI.
//
// (i)  hgpu_compact_lr replaces load_alignment's filters/sort, process_lr_alignment_group,
//      fix_overlapping_alignments and build_compact_longreads (reference src/haslr_assemble/src/Longread.cpp:182-302,
//      374-624). One warp per long read: the read's hit rows are loaded with coalesced column loads, the load
//      filters F1-F4 run one hit per lane and survivors are compacted with a ballot/popc warp scan; the order-dependent
//      tail (libstdc++-ordered sort, palindrome cut, overlap trimming on run-length CIGARs, weighted interval
//      scheduling) is k1_core.cuh's k1_process_read on one lane — groups are ~5-30 hits.
// (ii) hgpu_backbone_edges replaces bbg_build_graph/bbg_add_edge and the rule of bbg_remove_weak_edges
//      (Backbone_graph.cpp:10-25,148-171,348-375): every adjacent pair upserts its edge key and its twin's into an
//      open-address hash (64-bit keys, atomicCAS claim, match.any warp-aggregated count increments); unique keys are
//      then ordered like the reference's nested std::map iteration (counting sort by `from`, tiny per-bucket sort
//      by `to`), supports are scattered and ordered by (read, element) per entry.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "k1_core.cuh"
#include "scan.cuh"

using namespace hgpu;

struct K12State {
    // K1
    DevBuf<uint32_t> col[8], cg_off, cg_ops, read_off, out_cnt, out_off, idx;
    DevBuf<uint8_t> is_rev, mapq, take;
    DevBuf<double> mean_kmer;
    DevBuf<K1Hit> hit;
    DevBuf<uint32_t> dp, cand; DevBuf<int32_t> prevc;
    DevBuf<ClElem> tmp, out;
    // K2
    DevBuf<uint32_t> cl_tid, cl_read_off, slot_of, h_cnt, from_hist, from_off, bucket_cur, ent_slot, ent_cnt, supp_off, supp_cur, slot_rank;
    DevBuf<uint8_t> cl_rev, keep;
    DevBuf<unsigned long long> h_key, ent_key;
    DevBuf<hgpu_edge_supp> supp, supp_tmp;
    DevBuf<uint32_t> scalars;
};
void k12_state_destroy(K12State* s) { delete s; }
static K12State* k12_state(hgpu_t* ctx) { if (!ctx->k12) ctx->k12 = new K12State(); return ctx->k12; }



// ---------------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------------
struct K1Args {
    HitCols h; const uint32_t* read_off; uint32_t n_reads; const double* mean_kmer; K1Params p;
    uint32_t max_group;                       // scratch stride per warp
    uint32_t* idx; K1Hit* hit; uint32_t* dp; int32_t* prevc; uint32_t* cand; uint8_t* take;
    ClElem* tmp; uint32_t* out_cnt;
};

__global__ void __launch_bounds__(128) k1_compact_lr(K1Args a) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const size_t so = (size_t)gw * a.max_group;
    uint32_t* idx = a.idx + so;
    for (uint32_t r = gw; r < a.n_reads; r += nw) {
        const uint32_t b = a.read_off[r], e = a.read_off[r + 1];
        uint32_t cnt = 0;
        for (uint32_t base = b; base < e; base += 32) {
            const uint32_t i = base + lane;
            const bool ok = i < e && k1_load_filter(a.h, i, a.mean_kmer, a.p);
            const unsigned m = __ballot_sync(FULLM, ok);
            if (ok) idx[cnt + __popc(m & ((1u << lane) - 1))] = i;
            cnt += __popc(m);
        }
        __syncwarp();
        if (lane == 0)
            a.out_cnt[r] = k1_process_read(a.h, a.mean_kmer, a.p, idx, cnt, a.hit + so, a.dp + so, a.prevc + so, a.cand + so,
                                           a.take + so, a.tmp + b);
        __syncwarp();
    }
}

// pack per-read element runs (stored at the read's first hit row) into the output order
__global__ void __launch_bounds__(256) k1_pack(const ClElem* tmp, const uint32_t* read_off, const uint32_t* out_cnt, const uint32_t* out_off,
                                               uint32_t n_reads, ClElem* out) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t r = gw; r < n_reads; r += nw) {
        const uint32_t n = out_cnt[r];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tmp + read_off[r]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + out_off[r]);
        for (uint32_t k = lane; k < n * (sizeof(ClElem) / 4); k += 32) dst[k] = src[k];
    }
}

extern "C" int hgpu_compact_lr(hgpu_t* ctx, const hgpu_hits_t* hits, const uint32_t* read_off, uint32_t n_reads,
                               const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm,
                               hgpu_cl_elem* out_elems, uint32_t* out_read_off, uint64_t* out_n) {
    static_assert(sizeof(ClElem) == sizeof(hgpu_cl_elem), "element layout");
    if (!ctx) return HGPU_E_INVALID;
    if (!hits || !read_off || !prm || !out_read_off || !out_n || (n_contigs && !mean_kmer)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    const uint32_t n_hits = hits->n_hits;
    if (read_off[n_reads] != n_hits) HGPU_FAIL(ctx, HGPU_E_INVALID, "read_off[n_reads] = %u but n_hits = %u", read_off[n_reads], n_hits);
    uint32_t max_group = 1;
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (read_off[r + 1] < read_off[r]) HGPU_FAIL(ctx, HGPU_E_INVALID, "read_off not monotone at read %u", r);
        max_group = std::max(max_group, read_off[r + 1] - read_off[r]);
    }
    for (uint32_t i = 0; i < n_hits; ++i)   // Q1: the reference indexes mean_kmer[t_id] unchecked; refuse instead of reading out of bounds
        if (hits->t_id[i] >= n_contigs) HGPU_FAIL(ctx, HGPU_E_INVALID, "hit %u names contig %u >= n_contigs %u", i, hits->t_id[i], n_contigs);
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    K12State* S = k12_state(ctx);
    cudaStream_t st = ctx->stream;
    *out_n = 0;
    for (uint32_t r = 0; r <= n_reads; ++r) out_read_off[r] = 0;
    if (n_reads == 0) return HGPU_OK;

    const uint32_t* cols[8] = {hits->q_start, hits->q_end, hits->t_id, hits->t_len, hits->t_start, hits->t_end, hits->n_match, hits->n_block};
    for (int c = 0; c < 8; ++c) {
        HGPU_CUDA(ctx, S->col[c].ensure(n_hits + 1));
        if (n_hits) HGPU_CUDA(ctx, cudaMemcpyAsync(S->col[c].p, cols[c], (size_t)n_hits * 4, cudaMemcpyHostToDevice, st));
    }
    HGPU_CUDA(ctx, S->is_rev.ensure(n_hits + 1)); HGPU_CUDA(ctx, S->mapq.ensure(n_hits + 1));
    HGPU_CUDA(ctx, S->cg_off.ensure(n_hits + 1));
    const uint32_t n_ops = n_hits ? hits->cg_off[n_hits] : 0;
    HGPU_CUDA(ctx, S->cg_ops.ensure(n_ops + 1));
    if (n_hits) {
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->is_rev.p, hits->is_rev, n_hits, cudaMemcpyHostToDevice, st));
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->mapq.p, hits->mapq, n_hits, cudaMemcpyHostToDevice, st));
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->cg_ops.p, hits->cg_ops, (size_t)n_ops * 4, cudaMemcpyHostToDevice, st));
    }
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->cg_off.p, hits->cg_off, (size_t)(n_hits + 1) * 4, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, S->read_off.ensure(n_reads + 1));
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->read_off.p, read_off, (size_t)(n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, S->mean_kmer.ensure(n_contigs + 1));
    if (n_contigs) HGPU_CUDA(ctx, cudaMemcpyAsync(S->mean_kmer.p, mean_kmer, (size_t)n_contigs * 8, cudaMemcpyHostToDevice, st));

    // launch geometry: persistent warps, scratch sized by the largest group
    uint32_t blocks = (uint32_t)ctx->sm_count * 8;
    blocks = std::min<uint32_t>(blocks, (n_reads + 3) / 4);
    blocks = std::max<uint32_t>(blocks, 1);
    while (blocks > 1 && (uint64_t)blocks * 4 * max_group * 72 > (4ull << 30)) blocks = (blocks + 1) / 2;   // bound scratch to 4 GB
    const size_t n_warps = (size_t)blocks * 4, sc = n_warps * max_group;
    HGPU_CUDA(ctx, S->idx.ensure(sc)); HGPU_CUDA(ctx, S->hit.ensure(sc)); HGPU_CUDA(ctx, S->dp.ensure(sc));
    HGPU_CUDA(ctx, S->prevc.ensure(sc)); HGPU_CUDA(ctx, S->cand.ensure(sc)); HGPU_CUDA(ctx, S->take.ensure(sc));
    HGPU_CUDA(ctx, S->tmp.ensure(n_hits + 1)); HGPU_CUDA(ctx, S->out.ensure(n_hits + 1));
    HGPU_CUDA(ctx, S->out_cnt.ensure(n_reads + 1)); HGPU_CUDA(ctx, S->out_off.ensure(n_reads + 2));

    K1Args a{};
    a.h = HitCols{S->col[0].p, S->col[1].p, S->col[2].p, S->col[3].p, S->col[4].p, S->col[5].p, S->col[6].p, S->col[7].p,
                  S->is_rev.p, S->mapq.p, S->cg_off.p, S->cg_ops.p};
    a.read_off = S->read_off.p; a.n_reads = n_reads; a.mean_kmer = S->mean_kmer.p;
    a.p = K1Params{prm->min_aln_sim, prm->uniq_freq, prm->max_uniq_dev, prm->min_aln_block, prm->min_aln_mapq};
    a.max_group = max_group;
    a.idx = S->idx.p; a.hit = S->hit.p; a.dp = S->dp.p; a.prevc = S->prevc.p; a.cand = S->cand.p; a.take = S->take.p;
    a.tmp = S->tmp.p; a.out_cnt = S->out_cnt.p;
    k1_compact_lr<<<blocks, 128, 0, st>>>(a);
    HGPU_CUDA(ctx, cudaGetLastError());
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->out_cnt.p, S->out_off.p, n_reads);
    HGPU_CUDA(ctx, cudaGetLastError());
    k1_pack<<<std::min<uint32_t>((n_reads + 7) / 8, (uint32_t)ctx->sm_count * 8), 256, 0, st>>>(S->tmp.p, S->read_off.p, S->out_cnt.p, S->out_off.p, n_reads, S->out.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    HGPU_CUDA(ctx, cudaMemcpyAsync(out_read_off, S->out_off.p, (size_t)(n_reads + 1) * 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t n_out = out_read_off[n_reads];
    *out_n = n_out;
    if (n_out) {
        if (!out_elems) HGPU_FAIL(ctx, HGPU_E_INVALID, "null out_elems");
        HGPU_CUDA(ctx, cudaMemcpyAsync(out_elems, S->out.p, (size_t)n_out * sizeof(ClElem), cudaMemcpyDeviceToHost, st));
        HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return HGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------------------------------------
static constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// claim-or-find the slot of `key`
__device__ __forceinline__ uint32_t hash_upsert(unsigned long long* keys, uint32_t mask, unsigned long long key) {
    uint32_t slot = hash64(key) & mask;
    while (true) {
        unsigned long long old = keys[slot];
        if (old == key) return slot;
        if (old == EMPTY_KEY) {
            old = atomicCAS(&keys[slot], EMPTY_KEY, key);
            if (old == EMPTY_KEY || old == key) return slot;
        }
        slot = (slot + 1) & mask;
    }
}

// one thread per compact-read element j; it owns the pair (j, j+1) when both are in the same read.
// Each pair upserts its key and its twin's key; increments are aggregated across the lanes of a warp that hit
// the same key (match.any) so one atomicAdd per distinct key per warp reaches memory.
__global__ void __launch_bounds__(256) k2_count(const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off, uint32_t n_reads,
                                                uint32_t n_elems, unsigned long long* h_key, uint32_t* h_cnt, uint32_t mask,
                                                uint32_t* slot_of /* [2*n_elems] slot of (pair j, side), or ~0 */,
                                                uint32_t* elem_read /* [n_elems] */) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    unsigned long long k[2] = {EMPTY_KEY, EMPTY_KEY};
    if (j < n_elems) {
        // read of element j: last r with cl_read_off[r] <= j
        uint32_t lo = 0, hi = n_reads;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (cl_read_off[mid] <= j) lo = mid; else hi = mid; }
        elem_read[j] = lo;
        if (j + 1 < cl_read_off[lo + 1]) {
            has = true;
            const uint32_t n1 = cl_tid[j], r1 = cl_rev[j], n2 = cl_tid[j + 1], r2 = cl_rev[j + 1];
            k[0] = ((unsigned long long)((n1 << 1) | r1) << 32) | ((n2 << 1) | r2);
            k[1] = ((unsigned long long)((n2 << 1) | (1 - r2)) << 32) | ((n1 << 1) | (1 - r1));
        }
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const unsigned active = __ballot_sync(FULLM, has);
        uint32_t slot = 0xFFFFFFFFu;
        if (has) {
            const unsigned peers = __match_any_sync(active, k[side]);
            const int leader = __ffs(peers) - 1;
            if ((int)(threadIdx.x & 31) == leader) {
                slot = hash_upsert(h_key, mask, k[side]);
                atomicAdd(&h_cnt[slot], (uint32_t)__popc(peers));
            }
            slot = __shfl_sync(peers, slot, leader);
        }
        if (j < n_elems) slot_of[2 * (size_t)j + side] = slot;
    }
}

// unique keys: histogram over `from`
__global__ void __launch_bounds__(256) k2_from_hist(const unsigned long long* h_key, uint32_t cap, uint32_t* from_hist) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && h_key[s] != EMPTY_KEY) atomicAdd(&from_hist[(uint32_t)(h_key[s] >> 32)], 1u);
}
// scatter unique slots into their `from` bucket (arbitrary order inside a bucket; fixed by k2_bucket_sort)
__global__ void __launch_bounds__(256) k2_scatter_unique(const unsigned long long* h_key, uint32_t cap, const uint32_t* from_off,
                                                         uint32_t* bucket_cur, uint32_t* ent_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && h_key[s] != EMPTY_KEY) {
        const uint32_t f = (uint32_t)(h_key[s] >> 32);
        ent_slot[from_off[f] + atomicAdd(&bucket_cur[f], 1u)] = s;
    }
}
// order each `from` bucket by `to` (bucket = out-degree of one contig end: a handful), then emit entry records
__global__ void __launch_bounds__(256) k2_bucket_sort(const unsigned long long* h_key, const uint32_t* h_cnt, const uint32_t* from_off,
                                                      uint32_t n_from, uint32_t* ent_slot, unsigned long long* ent_key, uint32_t* ent_cnt,
                                                      uint32_t* slot_rank, uint8_t* keep, uint32_t min_edge_sup) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_from) return;
    const uint32_t b = from_off[f], e = from_off[f + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        const uint32_t s = ent_slot[i];
        const unsigned long long ks = h_key[s];
        uint32_t q = i;
        while (q > b && h_key[ent_slot[q - 1]] > ks) { ent_slot[q] = ent_slot[q - 1]; --q; }
        ent_slot[q] = s;
    }
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t s = ent_slot[i];
        ent_key[i] = h_key[s]; ent_cnt[i] = h_cnt[s]; slot_rank[s] = i;
        keep[i] = h_cnt[s] >= min_edge_sup ? 1 : 0;                       // Backbone_graph.cpp:358
    }
}
// supports: (pair j, side) -> its entry, claimed position inside the entry's list
__global__ void __launch_bounds__(256) k2_scatter_supp(const uint32_t* slot_of, const uint32_t* elem_read, const uint32_t* cl_read_off,
                                                       uint32_t n_elems, const uint32_t* slot_rank, const uint32_t* supp_off,
                                                       uint32_t* supp_cur, hgpu_edge_supp* supp) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_elems) return;
    const uint32_t slot = slot_of[t];
    if (slot == 0xFFFFFFFFu) return;
    const uint32_t j = t >> 1, side = t & 1, r = elem_read[j], lj = j - cl_read_off[r];
    const uint32_t ent = slot_rank[slot];
    hgpu_edge_supp s;
    s.lr_id_strand = r | (side << 31);
    s.cmp_head = side ? lj + 1 : lj;                                        // Backbone_graph.cpp:23-24
    s.cmp_tail = side ? lj : lj + 1;
    supp[supp_off[ent] + atomicAdd(&supp_cur[ent], 1u)] = s;
}
// per entry: order supports by (read, pair index, side) = the order bbg_build_graph appends them in
__device__ __forceinline__ unsigned long long supp_order(const hgpu_edge_supp& s) {
    const uint32_t side = s.lr_id_strand >> 31, pj = side ? s.cmp_tail : s.cmp_head;
    return ((unsigned long long)(s.lr_id_strand & 0x7FFFFFFFu) << 32) | ((unsigned long long)pj << 1) | side;
}
__global__ void __launch_bounds__(128) k2_supp_sort(const uint32_t* supp_off, uint32_t n_ent, const hgpu_edge_supp* in, hgpu_edge_supp* out) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t ent = gw; ent < n_ent; ent += nw) {
        const uint32_t b = supp_off[ent], n = supp_off[ent + 1] - b;
        for (uint32_t i = lane; i < n; i += 32) {            // rank sort: keys are distinct
            const hgpu_edge_supp s = in[b + i];
            const unsigned long long k = supp_order(s);
            uint32_t rank = 0;
            for (uint32_t q = 0; q < n; ++q) rank += supp_order(in[b + q]) < k ? 1u : 0u;
            out[b + rank] = s;
        }
    }
}

extern "C" int hgpu_backbone_edges(hgpu_t* ctx, const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off,
                                   uint32_t n_reads, uint32_t min_edge_sup,
                                   uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint8_t* out_keep,
                                   uint64_t* out_n_entries) {
    if (!ctx) return HGPU_E_INVALID;
    if (!cl_read_off || !out_n_entries || !out_supp_off) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    K12State* S = k12_state(ctx);
    cudaStream_t st = ctx->stream;
    *out_n_entries = 0; out_supp_off[0] = 0;
    const uint32_t n_elems = cl_read_off[n_reads];
    uint64_t n_pairs = 0; uint32_t max_tid = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (cl_read_off[r + 1] < cl_read_off[r]) HGPU_FAIL(ctx, HGPU_E_INVALID, "cl_read_off not monotone at read %u", r);
        uint32_t c = cl_read_off[r + 1] - cl_read_off[r];
        if (c > 1) n_pairs += c - 1;
    }
    if (n_pairs == 0) return HGPU_OK;
    if (!cl_tid || !cl_rev || !out_key || !out_supp || !out_keep) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (n_reads >= 0x80000000u) HGPU_FAIL(ctx, HGPU_E_INVALID, "read ids must fit 31 bits");
    for (uint32_t j = 0; j < n_elems; ++j) {
        if (cl_tid[j] >= 0x7FFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "contig id %u does not fit 31 bits", cl_tid[j]);
        if (cl_rev[j] > 1) HGPU_FAIL(ctx, HGPU_E_INVALID, "cl_rev[%u] = %u", j, cl_rev[j]);
        max_tid = std::max(max_tid, cl_tid[j]);
    }
    const uint32_t n_from = 2 * (max_tid + 1);
    uint32_t cap = 1024;
    while ((uint64_t)cap < 4 * n_pairs) cap <<= 1;           // load factor <= 0.5
    const uint32_t mask = cap - 1;

    HGPU_CUDA(ctx, S->cl_tid.ensure(n_elems)); HGPU_CUDA(ctx, S->cl_rev.ensure(n_elems)); HGPU_CUDA(ctx, S->cl_read_off.ensure(n_reads + 1));
    HGPU_CUDA(ctx, S->h_key.ensure(cap)); HGPU_CUDA(ctx, S->h_cnt.ensure(cap)); HGPU_CUDA(ctx, S->slot_rank.ensure(cap));
    HGPU_CUDA(ctx, S->slot_of.ensure(2 * (size_t)n_elems)); HGPU_CUDA(ctx, S->out_cnt.ensure(n_elems));
    HGPU_CUDA(ctx, S->from_hist.ensure(n_from + 1)); HGPU_CUDA(ctx, S->from_off.ensure(n_from + 2)); HGPU_CUDA(ctx, S->bucket_cur.ensure(n_from + 1));
    const size_t ecap = 2 * n_pairs;
    HGPU_CUDA(ctx, S->ent_slot.ensure(ecap)); HGPU_CUDA(ctx, S->ent_key.ensure(ecap)); HGPU_CUDA(ctx, S->ent_cnt.ensure(ecap + 1));
    HGPU_CUDA(ctx, S->keep.ensure(ecap)); HGPU_CUDA(ctx, S->supp_off.ensure(ecap + 2)); HGPU_CUDA(ctx, S->supp_cur.ensure(ecap + 1));
    HGPU_CUDA(ctx, S->supp.ensure(ecap)); HGPU_CUDA(ctx, S->supp_tmp.ensure(ecap));

    HGPU_CUDA(ctx, cudaMemcpyAsync(S->cl_tid.p, cl_tid, (size_t)n_elems * 4, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->cl_rev.p, cl_rev, n_elems, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->cl_read_off.p, cl_read_off, (size_t)(n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->h_key.p, 0xFF, (size_t)cap * 8, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->h_cnt.p, 0, (size_t)cap * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->from_hist.p, 0, (size_t)(n_from + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->bucket_cur.p, 0, (size_t)(n_from + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->supp_cur.p, 0, (size_t)(ecap + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->ent_cnt.p, 0, (size_t)(ecap + 1) * 4, st));

    uint32_t* elem_read = S->out_cnt.p;   // reuse: [n_elems]
    k2_count<<<(n_elems + 255) / 256, 256, 0, st>>>(S->cl_tid.p, S->cl_rev.p, S->cl_read_off.p, n_reads, n_elems, S->h_key.p, S->h_cnt.p, mask,
                                                    S->slot_of.p, elem_read);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_from_hist<<<(cap + 255) / 256, 256, 0, st>>>(S->h_key.p, cap, S->from_hist.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->from_hist.p, S->from_off.p, n_from);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_scatter_unique<<<(cap + 255) / 256, 256, 0, st>>>(S->h_key.p, cap, S->from_off.p, S->bucket_cur.p, S->ent_slot.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_bucket_sort<<<(n_from + 255) / 256, 256, 0, st>>>(S->h_key.p, S->h_cnt.p, S->from_off.p, n_from, S->ent_slot.p, S->ent_key.p, S->ent_cnt.p,
                                                         S->slot_rank.p, S->keep.p, min_edge_sup);
    HGPU_CUDA(ctx, cudaGetLastError());
    uint32_t n_ent = 0;
    HGPU_CUDA(ctx, cudaMemcpyAsync(&n_ent, S->from_off.p + n_from, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->ent_cnt.p, S->supp_off.p, n_ent);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_scatter_supp<<<(2 * n_elems + 255) / 256, 256, 0, st>>>(S->slot_of.p, elem_read, S->cl_read_off.p, n_elems, S->slot_rank.p, S->supp_off.p,
                                                               S->supp_cur.p, S->supp_tmp.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_supp_sort<<<std::min<uint32_t>((n_ent + 3) / 4, (uint32_t)ctx->sm_count * 16), 128, 0, st>>>(S->supp_off.p, n_ent, S->supp_tmp.p, S->supp.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 8;

    HGPU_CUDA(ctx, cudaMemcpyAsync(out_key, S->ent_key.p, (size_t)n_ent * 8, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(out_supp_off, S->supp_off.p, (size_t)(n_ent + 1) * 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(out_keep, S->keep.p, n_ent, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(out_supp, S->supp.p, (size_t)2 * n_pairs * sizeof(hgpu_edge_supp), cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    *out_n_entries = n_ent;
    return HGPU_OK;
}
