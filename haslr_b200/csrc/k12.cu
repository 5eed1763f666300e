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
#include "sort.cuh"

using namespace hgpu;

struct K12State {
    // K1
    DevBuf<uint32_t> col[8], cg_off, cg_ops, read_off, out_cnt, out_off, idx, cg_total, status;
    DevBuf<uint8_t> is_rev, mapq, take, pass;
    DevBuf<double> mean_kmer;
    DevBuf<K1Hit> hit;
    DevBuf<uint32_t> dp, cand; DevBuf<int32_t> prevc;
    DevBuf<ClElem> tmp, out;
    DevBuf<uint32_t> out_tid; DevBuf<uint8_t> out_rev;
    // K2
    DevBuf<uint32_t> cl_tid, cl_read_off, slot_of, h_cnt, from_hist, from_off, bucket_cur, ent_slot, ent_cnt, supp_off, supp_cur, slot_rank;
    DevBuf<uint8_t> cl_rev, keep;
    DevBuf<unsigned long long> h_key, ent_key, sort_key;
    DevBuf<uint32_t> sort_idx;
    DevBuf<hgpu_edge_supp> supp, supp_tmp;
    DevBuf<uint32_t> scalars;
    DevBuf<unsigned long long> scan_tmp;
};
void k12_state_destroy(K12State* s) { delete s; }
static K12State* k12_state(hgpu_t* ctx) { if (!ctx->k12) ctx->k12 = new K12State(); return ctx->k12; }



// ---------------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------------
struct K1Args {
    HitCols h; const uint32_t* read_off; uint32_t n_reads, n_contigs, n_hits; const double* mean_kmer; K1Params p;
    // scratch, one entry per hit row: a read works in the rows it owns, so nothing is sized by the largest group
    uint32_t* idx; K1Hit* hit; uint32_t* dp; int32_t* prevc; uint32_t* cand; uint8_t* take; uint32_t* cg_total;
    ClElem* tmp; uint32_t* out_cnt;
    uint32_t* status;                         // [0] first row (min) that names a contig >= n_contigs, [1] first read whose offsets decrease
};

// K1 runs in three launches so that every step has the parallelism it can use:
//   k1_filter        one thread per PAF row: the load filters F1-F4 and the contig-id check (SoA columns, coalesced);
//   k1_cigar_totals  one warp per surviving row: its expanded CIGAR length (the tail only needs the totals);
//   k1_tail          one THREAD per long read: everything that depends on the order of the read's hits (libstdc++-ordered sort,
//                    palindrome cut, F5, overlap trimming on the run-length CIGARs, chaining) is inherently one instruction stream
//                    per read, a few hundred dependent global accesses long. With a warp per read (rounds 1 and early 2) 31 lanes
//                    sat idle beside it and 4,000 reads were in flight per GPU; with a thread per read it is 150,000, and the
//                    lanes of a warp overlap each other's memory latency wherever their control flow agrees.
// (Staging the read's CIGAR runs, sort keys and scratch in shared memory for the one-lane tail was measured first: 0.94 -> 1.06 ->
// 1.24 ms on config 2 - the shared memory cut the resident warps and with them the latency hiding. profiles/r2G_pool_shape_crit_k1_ab.log)
__global__ void __launch_bounds__(256) k1_filter(K1Args a, uint8_t* pass) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_hits) return;
    uint8_t f;
    if (a.h.t_id[i] >= a.n_contigs) { atomicMin(a.status, i); f = 2; }          // Q1: the reference indexes mean_kmer[t_id] unchecked
    else f = k1_load_filter(a.h, i, a.mean_kmer, a.p) ? 1 : 0;
    pass[i] = f;
}

__global__ void __launch_bounds__(256) k1_cigar_totals(K1Args a, const uint8_t* pass) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = gw; row < a.n_hits; row += nw) {
        if (pass[row] != 1) continue;
        const uint32_t k0 = a.h.cg_off[row], k1 = a.h.cg_off[row + 1];
        uint32_t tot = 0;
        for (uint32_t k = k0 + lane; k < k1; k += 32) tot += a.h.cg_ops[k] >> 2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(FULLM, tot, d);
        if (lane == 0) a.cg_total[row] = tot;
    }
}

__global__ void __launch_bounds__(128) k1_tail(K1Args a, const uint8_t* pass) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads) return;
    const uint32_t b = a.read_off[r], e = a.read_off[r + 1];
    if (e < b || e > a.n_hits) { atomicMin(a.status + 1, r); a.out_cnt[r] = 0; return; }
    uint32_t* idx = a.idx + b;
    uint32_t cnt = 0;
    bool bad = false;
    for (uint32_t i = b; i < e; ++i) {                                          // survivors in PAF order
        const uint8_t f = pass[i];
        if (f == 2) bad = true; else if (f == 1) idx[cnt++] = i;
    }
    if (bad) { a.out_cnt[r] = 0; return; }
    a.out_cnt[r] = k1_process_read(a.h, a.mean_kmer, a.p, idx, cnt, a.hit + b, a.dp + b, a.prevc + b, a.cand + b, a.take + b, a.tmp + b, a.cg_total);
}

// pack per-read element runs (stored at the read's first hit row) into the output order; also the per-element contig id and
// strand of the hit behind it (what the edge table, compact_uniq.txt and the coordinate log need of the hit table)
__global__ void __launch_bounds__(256) k1_pack(const ClElem* tmp, const uint32_t* read_off, const uint32_t* out_cnt, const uint32_t* out_off,
                                               uint32_t n_reads, const uint32_t* t_id, const uint8_t* is_rev, ClElem* out, uint32_t* out_tid, uint8_t* out_rev) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t r = gw; r < n_reads; r += nw) {
        const uint32_t n = out_cnt[r];
        const ClElem* srce = tmp + read_off[r];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(srce);
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + out_off[r]);
        for (uint32_t k = lane; k < n * (sizeof(ClElem) / 4); k += 32) dst[k] = src[k];
        for (uint32_t k = lane; k < n; k += 32) { const uint32_t hrow = srce[k].hit; out_tid[out_off[r] + k] = t_id[hrow]; out_rev[out_off[r] + k] = is_rev[hrow]; }
    }
}

// K1 on device-resident inputs; leaves elements / offsets / per-element (contig, strand) on the device and registers them
// as the context's compact reads
static int compact_lr_run(hgpu_t* ctx, const HitCols& h, uint32_t n_hits, const uint32_t* d_read_off, uint32_t n_reads,
                          const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm, uint32_t* out_read_off, uint64_t* out_n) {
    K12State* S = k12_state(ctx);
    cudaStream_t st = ctx->stream;
    ctx->compact = ResidentCompact();
    ctx->stage.ms_k1 = 0; ctx->stage.launches_k1 = 0; ctx->stage.k1_hits = n_hits; ctx->stage.k1_reads = n_reads; ctx->stage.k1_elems = 0;
    HGPU_CUDA(ctx, S->mean_kmer.ensure(n_contigs + 1));
    HGPU_H2D(ctx, S->mean_kmer.p, mean_kmer, (size_t)n_contigs * 8);
    const size_t sc = (size_t)n_hits + 1;
    HGPU_CUDA(ctx, S->idx.ensure(sc)); HGPU_CUDA(ctx, S->hit.ensure(sc)); HGPU_CUDA(ctx, S->dp.ensure(sc));
    HGPU_CUDA(ctx, S->prevc.ensure(sc)); HGPU_CUDA(ctx, S->cand.ensure(sc)); HGPU_CUDA(ctx, S->take.ensure(sc)); HGPU_CUDA(ctx, S->cg_total.ensure(sc)); HGPU_CUDA(ctx, S->pass.ensure(sc));
    HGPU_CUDA(ctx, S->tmp.ensure(sc)); HGPU_CUDA(ctx, S->out.ensure(sc)); HGPU_CUDA(ctx, S->out_tid.ensure(sc)); HGPU_CUDA(ctx, S->out_rev.ensure(sc));
    HGPU_CUDA(ctx, S->out_cnt.ensure(n_reads + 1)); HGPU_CUDA(ctx, S->out_off.ensure(n_reads + 2)); HGPU_CUDA(ctx, S->status.ensure(4));
    HGPU_CUDA(ctx, S->scan_tmp.ensure(scan_tmp_entries(n_reads)));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->status.p, 0xFF, 16, st));

    K1Args a{};
    a.h = h; a.read_off = d_read_off; a.n_reads = n_reads; a.n_contigs = n_contigs; a.n_hits = n_hits; a.mean_kmer = S->mean_kmer.p;
    a.p = K1Params{prm->min_aln_sim, prm->uniq_freq, prm->max_uniq_dev, prm->min_aln_block, prm->min_aln_mapq};
    a.idx = S->idx.p; a.hit = S->hit.p; a.dp = S->dp.p; a.prevc = S->prevc.p; a.cand = S->cand.p; a.take = S->take.p; a.cg_total = S->cg_total.p;
    a.tmp = S->tmp.p; a.out_cnt = S->out_cnt.p; a.status = S->status.p;
    stage_begin(ctx, ctx->ev_k1);
    if (n_hits) {
        k1_filter<<<(n_hits + 255) / 256, 256, 0, st>>>(a, S->pass.p);
        HGPU_CUDA(ctx, cudaGetLastError());
        k1_cigar_totals<<<std::min<uint32_t>((n_hits + 7) / 8, (uint32_t)ctx->sm_count * 32), 256, 0, st>>>(a, S->pass.p);
        HGPU_CUDA(ctx, cudaGetLastError());
    }
    k1_tail<<<(n_reads + 127) / 128, 128, 0, st>>>(a, S->pass.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    const int scan_launches = scan_u32(st, S->out_cnt.p, S->out_off.p, n_reads, S->scan_tmp.p, nullptr);
    HGPU_CUDA(ctx, cudaGetLastError());
    k1_pack<<<std::max<uint32_t>(1, std::min<uint32_t>((n_reads + 7) / 8, (uint32_t)ctx->sm_count * 8)), 256, 0, st>>>(
        S->tmp.p, d_read_off, S->out_cnt.p, S->out_off.p, n_reads, h.t_id, h.is_rev, S->out.p, S->out_tid.p, S->out_rev.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    stage_end(ctx, ctx->ev_k1);
    ctx->launches += 4 + scan_launches; ctx->stage.launches_k1 = 4 + scan_launches;
    uint32_t status[2];
    HGPU_CUDA(ctx, cudaMemcpyAsync(status, S->status.p, 8, cudaMemcpyDeviceToHost, st));
    HGPU_D2H(ctx, out_read_off, S->out_off.p, (size_t)(n_reads + 1) * 4);
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stage.ms_k1 = stage_ms(ctx, ctx->ev_k1);
    if (status[1] != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "read_off not monotone at read %u", status[1]);
    if (status[0] != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "hit %u names a contig >= n_contigs %u", status[0], n_contigs);
    *out_n = out_read_off[n_reads];
    ctx->stage.k1_elems = *out_n;
    ctx->compact.valid = true; ctx->compact.n_elems = (uint32_t)*out_n; ctx->compact.n_reads = n_reads;
    ctx->compact.elems = reinterpret_cast<const hgpu_cl_elem*>(S->out.p); ctx->compact.read_off = S->out_off.p;
    return HGPU_OK;
}

static int compact_lr_download(hgpu_t* ctx, uint64_t n_out, hgpu_cl_elem* out_elems, uint32_t* out_tid, uint8_t* out_rev) {
    K12State* S = k12_state(ctx);
    if (n_out == 0) return HGPU_OK;
    if (out_elems) HGPU_D2H(ctx, out_elems, S->out.p, (size_t)n_out * sizeof(ClElem));
    if (out_tid) HGPU_D2H(ctx, out_tid, S->out_tid.p, (size_t)n_out * 4);
    if (out_rev) HGPU_D2H(ctx, out_rev, S->out_rev.p, (size_t)n_out);
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HGPU_OK;
}

extern "C" int hgpu_compact_lr(hgpu_t* ctx, const hgpu_hits_t* hits, const uint32_t* read_off, uint32_t n_reads,
                               const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm,
                               hgpu_cl_elem* out_elems, uint32_t* out_read_off, uint64_t* out_n) {
    static_assert(sizeof(ClElem) == sizeof(hgpu_cl_elem), "element layout");
    if (!ctx) return HGPU_E_INVALID;
    if (!hits || !read_off || !prm || !out_read_off || !out_n || (n_contigs && !mean_kmer)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    const uint32_t n_hits = hits->n_hits;
    if (read_off[n_reads] != n_hits) HGPU_FAIL(ctx, HGPU_E_INVALID, "read_off[n_reads] = %u but n_hits = %u", read_off[n_reads], n_hits);
    if (n_reads && read_off[0] != 0) HGPU_FAIL(ctx, HGPU_E_INVALID, "read_off[0] = %u", read_off[0]);
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    K12State* S = k12_state(ctx);
    *out_n = 0;
    for (uint32_t r = 0; r <= n_reads; ++r) out_read_off[r] = 0;
    if (n_reads == 0) return HGPU_OK;

    // host columns -> this stage's own device copies (a resident table made by hgpu_paf_tokenize is left alone)
    const uint32_t* cols[8] = {hits->q_start, hits->q_end, hits->t_id, hits->t_len, hits->t_start, hits->t_end, hits->n_match, hits->n_block};
    for (int c = 0; c < 8; ++c) {
        HGPU_CUDA(ctx, S->col[c].ensure(n_hits + 1));
        HGPU_H2D(ctx, S->col[c].p, cols[c], (size_t)n_hits * 4);
    }
    HGPU_CUDA(ctx, S->is_rev.ensure(n_hits + 1)); HGPU_CUDA(ctx, S->mapq.ensure(n_hits + 1));
    HGPU_CUDA(ctx, S->cg_off.ensure(n_hits + 1));
    const uint32_t n_ops = n_hits ? hits->cg_off[n_hits] : 0;
    HGPU_CUDA(ctx, S->cg_ops.ensure(n_ops + 1));
    HGPU_H2D(ctx, S->is_rev.p, hits->is_rev, n_hits);
    HGPU_H2D(ctx, S->mapq.p, hits->mapq, n_hits);
    HGPU_H2D(ctx, S->cg_ops.p, hits->cg_ops, (size_t)n_ops * 4);
    HGPU_H2D(ctx, S->cg_off.p, hits->cg_off, (size_t)(n_hits + 1) * 4);
    HGPU_CUDA(ctx, S->read_off.ensure(n_reads + 1));
    HGPU_H2D(ctx, S->read_off.p, read_off, (size_t)(n_reads + 1) * 4);
    const HitCols h{S->col[0].p, S->col[1].p, S->col[2].p, S->col[3].p, S->col[4].p, S->col[5].p, S->col[6].p, S->col[7].p,
                    S->is_rev.p, S->mapq.p, S->cg_off.p, S->cg_ops.p};
    // the uploaded columns become the context's resident hit table (the _dev stages that follow read it)
    ctx->hits = ResidentHits();
    ctx->hits.valid = true; ctx->hits.grouped = true; ctx->hits.n_hits = n_hits; ctx->hits.n_reads = n_reads; ctx->hits.n_ops = n_ops;
    ctx->hits.q_start = h.q_start; ctx->hits.q_end = h.q_end; ctx->hits.t_id = h.t_id; ctx->hits.t_len = h.t_len; ctx->hits.t_start = h.t_start;
    ctx->hits.t_end = h.t_end; ctx->hits.n_match = h.n_match; ctx->hits.n_block = h.n_block; ctx->hits.is_rev = h.is_rev; ctx->hits.mapq = h.mapq;
    ctx->hits.cg_off = h.cg_off; ctx->hits.cg_ops = h.cg_ops; ctx->hits.read_off = S->read_off.p;
    int rc = compact_lr_run(ctx, h, n_hits, S->read_off.p, n_reads, mean_kmer, n_contigs, prm, out_read_off, out_n);
    if (rc) return rc;
    if (*out_n && !out_elems) HGPU_FAIL(ctx, HGPU_E_INVALID, "null out_elems");
    return compact_lr_download(ctx, *out_n, out_elems, nullptr, nullptr);
}

extern "C" int hgpu_compact_lr_dev(hgpu_t* ctx, uint32_t n_reads, const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm,
                                   hgpu_cl_elem* out_elems, uint32_t* out_tid, uint8_t* out_rev, uint32_t* out_read_off, uint64_t* out_n) {
    if (!ctx) return HGPU_E_INVALID;
    if (!prm || !out_read_off || !out_n || (n_contigs && !mean_kmer)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    const ResidentHits& rh = ctx->hits;
    if (!rh.valid || !rh.grouped || rh.n_reads != n_reads)
        HGPU_FAIL(ctx, HGPU_E_INVALID, "hgpu_compact_lr_dev needs the resident hit table grouped for %u reads (hgpu_paf_tokenize + hgpu_hits_group)", n_reads);
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    *out_n = 0;
    for (uint32_t r = 0; r <= n_reads; ++r) out_read_off[r] = 0;
    if (n_reads == 0) return HGPU_OK;
    const HitCols h{rh.q_start, rh.q_end, rh.t_id, rh.t_len, rh.t_start, rh.t_end, rh.n_match, rh.n_block, rh.is_rev, rh.mapq, rh.cg_off, rh.cg_ops};
    int rc = compact_lr_run(ctx, h, rh.n_hits, rh.read_off, n_reads, mean_kmer, n_contigs, prm, out_read_off, out_n);
    if (rc) return rc;
    return compact_lr_download(ctx, *out_n, out_elems, out_tid, out_rev);
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
__global__ void __launch_bounds__(128) k2_supp_sort(const uint32_t* supp_off, uint32_t n_ent, const hgpu_edge_supp* in, hgpu_edge_supp* out,
                                                    unsigned long long* sort_key, uint32_t* sort_idx) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t ent = gw; ent < n_ent; ent += nw) {
        const uint32_t b = supp_off[ent], n = supp_off[ent + 1] - b;
        if (n <= SORT_RANK_MAX) {
            for (uint32_t i = lane; i < n; i += 32) {            // rank sort: keys are distinct
                const hgpu_edge_supp s = in[b + i];
                const unsigned long long k = supp_order(s);
                uint32_t rank = 0;
                for (uint32_t q = 0; q < n; ++q) rank += supp_order(in[b + q]) < k ? 1u : 0u;
                out[b + rank] = s;
            }
        } else {                                                  // a contig end with many supports: n log^2 n instead of n^2 (sort.cuh)
            for (uint32_t i = lane; i < n; i += 32) { sort_key[b + i] = supp_order(in[b + i]); sort_idx[b + i] = i; }
            warp_bitonic_u64<true>(sort_key + b, sort_idx + b, n, lane);
            for (uint32_t i = lane; i < n; i += 32) out[b + i] = in[b + sort_idx[b + i]];
        }
    }
}

// in-kernel input checks of the edge table stage: flags[0] = first element whose contig id does not fit 31 bits or whose strand
// is not 0/1, flags[1] = largest contig id, flags[2] (with flags[3] as the high word) = number of adjacent pairs
__global__ void __launch_bounds__(256) k2_scan_input(const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off, uint32_t n_reads,
                                                     uint32_t n_elems, uint32_t* flags) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mx = 0;
    if (j < n_elems) {
        const uint32_t t = cl_tid[j];
        if (t >= 0x7FFFFFFFu || cl_rev[j] > 1) atomicMin(flags, j);
        mx = t;
    }
    mx = __reduce_max_sync(FULLM, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(flags + 1, mx);
    unsigned long long pairs = 0;
    if (j < n_reads) {
        const uint32_t b = cl_read_off[j], e = cl_read_off[j + 1];
        if (e < b || e > n_elems) atomicMin(flags + 4, j);
        else if (e - b > 1) pairs = e - b - 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(FULLM, pairs, d);
    if ((threadIdx.x & 31) == 0 && pairs) atomicAdd(reinterpret_cast<unsigned long long*>(flags + 2), pairs);
}

// K2 on device-resident compact reads (d_tid / d_rev per element, d_read_off per read)
static int backbone_edges_run(hgpu_t* ctx, const uint32_t* d_tid, const uint8_t* d_rev, const uint32_t* d_read_off, uint32_t n_reads, uint32_t n_elems,
                              uint32_t min_edge_sup, uint64_t entry_cap, uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp,
                              uint8_t* out_keep, uint64_t* out_n_entries) {
    K12State* S = k12_state(ctx);
    cudaStream_t st = ctx->stream;
    ctx->stage.ms_k2 = 0; ctx->stage.launches_k2 = 0; ctx->stage.k2_pairs = 0; ctx->stage.k2_entries = 0;
    *out_n_entries = 0; out_supp_off[0] = 0;
    if (n_reads >= 0x80000000u) HGPU_FAIL(ctx, HGPU_E_INVALID, "read ids must fit 31 bits");
    HGPU_CUDA(ctx, S->scalars.ensure(8));
    const uint32_t init[6] = {0xFFFFFFFFu, 0u, 0u, 0u, 0xFFFFFFFFu, 0u};
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->scalars.p, init, sizeof init, cudaMemcpyHostToDevice, st));
    stage_begin(ctx, ctx->ev_k2);
    const uint32_t n_scan = std::max(n_elems, n_reads);
    k2_scan_input<<<(n_scan + 255) / 256, 256, 0, st>>>(d_tid, d_rev, d_read_off, n_reads, n_elems, S->scalars.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    uint32_t flags[6];
    HGPU_CUDA(ctx, cudaMemcpyAsync(flags, S->scalars.p, sizeof flags, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1; ctx->stage.launches_k2 = 1;
    if (flags[4] != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "cl_read_off not monotone at read %u", flags[4]);
    if (flags[0] != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "element %u: contig id does not fit 31 bits or strand is not 0/1", flags[0]);
    const uint64_t n_pairs = ((uint64_t)flags[3] << 32) | flags[2];
    ctx->stage.k2_pairs = n_pairs;
    if (n_pairs == 0) return HGPU_OK;
    if (!out_key || !out_supp || !out_keep) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (2 * n_pairs > entry_cap) HGPU_FAIL(ctx, HGPU_E_NOSPACE, "edge table needs room for %llu entries, caller gave %llu", (unsigned long long)(2 * n_pairs), (unsigned long long)entry_cap);
    if (4 * n_pairs >= (1ull << 31)) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "%llu adjacent pairs: shard the edge table (SURVEY 8e)", (unsigned long long)n_pairs);
    const uint32_t max_tid = flags[1];
    const uint32_t n_from = 2 * (max_tid + 1);
    uint32_t cap = 1024;
    while ((uint64_t)cap < 4 * n_pairs) cap <<= 1;           // load factor <= 0.5
    const uint32_t mask = cap - 1;

    HGPU_CUDA(ctx, S->h_key.ensure(cap)); HGPU_CUDA(ctx, S->h_cnt.ensure(cap)); HGPU_CUDA(ctx, S->slot_rank.ensure(cap));
    HGPU_CUDA(ctx, S->slot_of.ensure(2 * (size_t)n_elems)); HGPU_CUDA(ctx, S->out_cnt.ensure(std::max<size_t>(n_elems, (size_t)n_reads + 1)));
    HGPU_CUDA(ctx, S->from_hist.ensure(n_from + 1)); HGPU_CUDA(ctx, S->from_off.ensure(n_from + 2)); HGPU_CUDA(ctx, S->bucket_cur.ensure(n_from + 1));
    const size_t ecap = 2 * n_pairs;
    HGPU_CUDA(ctx, S->ent_slot.ensure(ecap)); HGPU_CUDA(ctx, S->ent_key.ensure(ecap)); HGPU_CUDA(ctx, S->ent_cnt.ensure(ecap + 1));
    HGPU_CUDA(ctx, S->keep.ensure(ecap)); HGPU_CUDA(ctx, S->supp_off.ensure(ecap + 2)); HGPU_CUDA(ctx, S->supp_cur.ensure(ecap + 1));
    HGPU_CUDA(ctx, S->supp.ensure(ecap)); HGPU_CUDA(ctx, S->supp_tmp.ensure(ecap));
    HGPU_CUDA(ctx, S->sort_key.ensure(ecap)); HGPU_CUDA(ctx, S->sort_idx.ensure(ecap));
    HGPU_CUDA(ctx, S->scan_tmp.ensure(scan_tmp_entries((uint32_t)std::max<uint64_t>(n_from, ecap))));

    HGPU_CUDA(ctx, cudaMemsetAsync(S->h_key.p, 0xFF, (size_t)cap * 8, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->h_cnt.p, 0, (size_t)cap * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->from_hist.p, 0, (size_t)(n_from + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->bucket_cur.p, 0, (size_t)(n_from + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->supp_cur.p, 0, (size_t)(ecap + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->ent_cnt.p, 0, (size_t)(ecap + 1) * 4, st));

    uint32_t* elem_read = S->out_cnt.p;   // reuse: [n_elems] (K1's per-read counts are no longer needed)
    k2_count<<<(n_elems + 255) / 256, 256, 0, st>>>(d_tid, d_rev, d_read_off, n_reads, n_elems, S->h_key.p, S->h_cnt.p, mask, S->slot_of.p, elem_read);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_from_hist<<<(cap + 255) / 256, 256, 0, st>>>(S->h_key.p, cap, S->from_hist.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    int scan_launches = scan_u32(st, S->from_hist.p, S->from_off.p, n_from, S->scan_tmp.p, nullptr);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_scatter_unique<<<(cap + 255) / 256, 256, 0, st>>>(S->h_key.p, cap, S->from_off.p, S->bucket_cur.p, S->ent_slot.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_bucket_sort<<<(n_from + 255) / 256, 256, 0, st>>>(S->h_key.p, S->h_cnt.p, S->from_off.p, n_from, S->ent_slot.p, S->ent_key.p, S->ent_cnt.p,
                                                         S->slot_rank.p, S->keep.p, min_edge_sup);
    HGPU_CUDA(ctx, cudaGetLastError());
    uint32_t n_ent = 0;
    HGPU_CUDA(ctx, cudaMemcpyAsync(&n_ent, S->from_off.p + n_from, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    scan_launches += scan_u32(st, S->ent_cnt.p, S->supp_off.p, n_ent, S->scan_tmp.p, nullptr);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_scatter_supp<<<(2 * n_elems + 255) / 256, 256, 0, st>>>(S->slot_of.p, elem_read, d_read_off, n_elems, S->slot_rank.p, S->supp_off.p,
                                                               S->supp_cur.p, S->supp_tmp.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k2_supp_sort<<<std::min<uint32_t>((n_ent + 3) / 4, (uint32_t)ctx->sm_count * 16), 128, 0, st>>>(S->supp_off.p, n_ent, S->supp_tmp.p, S->supp.p,
                                                                                                    S->sort_key.p, S->sort_idx.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    stage_end(ctx, ctx->ev_k2);
    ctx->launches += 6 + scan_launches; ctx->stage.launches_k2 += 6 + scan_launches;

    HGPU_D2H(ctx, out_key, S->ent_key.p, (size_t)n_ent * 8);
    HGPU_D2H(ctx, out_supp_off, S->supp_off.p, (size_t)(n_ent + 1) * 4);
    HGPU_D2H(ctx, out_keep, S->keep.p, n_ent);
    HGPU_D2H(ctx, out_supp, S->supp.p, (size_t)2 * n_pairs * sizeof(hgpu_edge_supp));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stage.ms_k2 = stage_ms(ctx, ctx->ev_k2);
    ctx->stage.k2_entries = n_ent;
    *out_n_entries = n_ent;
    return HGPU_OK;
}

extern "C" int hgpu_backbone_edges(hgpu_t* ctx, const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off,
                                   uint32_t n_reads, uint32_t min_edge_sup,
                                   uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint8_t* out_keep,
                                   uint64_t* out_n_entries) {
    if (!ctx) return HGPU_E_INVALID;
    if (!cl_read_off || !out_n_entries || !out_supp_off) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    K12State* S = k12_state(ctx);
    *out_n_entries = 0; out_supp_off[0] = 0;
    const uint32_t n_elems = cl_read_off[n_reads];
    if (n_elems == 0) return HGPU_OK;
    if (!cl_tid || !cl_rev) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    HGPU_CUDA(ctx, S->cl_tid.ensure(n_elems)); HGPU_CUDA(ctx, S->cl_rev.ensure(n_elems)); HGPU_CUDA(ctx, S->cl_read_off.ensure(n_reads + 1));
    HGPU_H2D(ctx, S->cl_tid.p, cl_tid, (size_t)n_elems * 4);
    HGPU_H2D(ctx, S->cl_rev.p, cl_rev, n_elems);
    HGPU_H2D(ctx, S->cl_read_off.p, cl_read_off, (size_t)(n_reads + 1) * 4);
    // the documented capacity of the output arrays: two entries per adjacent pair (the kernel counts the pairs)
    return backbone_edges_run(ctx, S->cl_tid.p, S->cl_rev.p, S->cl_read_off.p, n_reads, n_elems, min_edge_sup, ~0ull,
                              out_key, out_supp_off, out_supp, out_keep, out_n_entries);
}

extern "C" int hgpu_backbone_edges_dev(hgpu_t* ctx, uint32_t min_edge_sup, uint64_t entry_cap,
                                       uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint8_t* out_keep,
                                       uint64_t* out_n_entries) {
    if (!ctx) return HGPU_E_INVALID;
    if (!out_n_entries || !out_supp_off) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (!ctx->compact.valid) HGPU_FAIL(ctx, HGPU_E_INVALID, "hgpu_backbone_edges_dev needs the compact reads hgpu_compact_lr[_dev] leaves on the device");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    K12State* S = k12_state(ctx);
    *out_n_entries = 0; out_supp_off[0] = 0;
    if (ctx->compact.n_elems == 0) return HGPU_OK;
    return backbone_edges_run(ctx, S->out_tid.p, S->out_rev.p, ctx->compact.read_off, ctx->compact.n_reads, ctx->compact.n_elems, min_edge_sup,
                              entry_cap, out_key, out_supp_off, out_supp, out_keep, out_n_entries);
}
