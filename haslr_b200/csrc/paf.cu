// (0) PAF text -> hit table: tokeniser kernels + C ABI.
//
// hgpu_paf_tokenize / hgpu_paf_fetch replace the text side of load_alignment (reference
// src/haslr_assemble/src/Longread.cpp:250-289: getline, str_split, str2type per column, cg:Z: search). The text goes to
// the device once; four passes over it, all HBM-streaming byte work:
//   k0_count_nl   every thread counts the line feeds of its 256-byte chunk            (1 B read per byte)
//   k0_line_start after a scan of the counts, the same threads write the line starts
//   k0_scan_lines one thread per line: column boundaries, cg:Z: payload, number of run-length operations
//   k0_emit       after scans of the "row kept" flags and operation counts: one thread per line parses the 12 columns
//                 into the SoA table hgpu_compact_lr takes and the CIGAR into (len << 2) | op words
// Algorithmic bytes: 4 reads of the text + 42 B per row + 4 B per CIGAR run written.
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "paf_core.cuh"
#include "scan.cuh"

using namespace hgpu;

struct PafState {
    DevBuf<char> text;
    DevBuf<uint32_t> chunk_cnt, chunk_off, keep, row_of, nops, line_cg_off, cg_off, cg_ops, col[10], scal, read_off;
    DevBuf<unsigned long long> line_start, scan_tmp, totals;
    DevBuf<uint8_t> is_rev, mapq;
    uint32_t n_rows = 0, n_lines = 0;
    uint64_t n_ops = 0, n_bytes = 0;
    bool ready = false;
};
void paf_state_destroy(PafState* s) { delete s; }

static constexpr uint32_t PAF_CHUNK = 256;

__global__ void __launch_bounds__(256) k0_count_nl(const char* text, uint64_t n, uint32_t n_chunks, uint32_t* cnt) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t b = (uint64_t)c * PAF_CHUNK, e = b + PAF_CHUNK < n ? b + PAF_CHUNK : n;
    uint32_t k = 0;
    if (e - b == PAF_CHUNK) {                                 // 16 bytes at a time (the buffer is 256-byte aligned)
        const uint4* p = reinterpret_cast<const uint4*>(text + b);
#pragma unroll 4
        for (uint32_t i = 0; i < PAF_CHUNK / 16; ++i) {
            const uint4 v = p[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k += __popc(__vcmpeq4(w[j], 0x0A0A0A0Au)) >> 3;   // 0xFF per byte that is a line feed
            }
        }
    } else {
        for (uint64_t i = b; i < e; ++i) k += text[i] == '\n';
    }
    cnt[c] = k;
}
// line l (l >= 1) starts after the l-th line feed; line 0 starts at byte 0
__global__ void __launch_bounds__(256) k0_line_start(const char* text, uint64_t n, uint32_t n_chunks, const uint32_t* off, unsigned long long* line_start) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t b = (uint64_t)c * PAF_CHUNK, e = b + PAF_CHUNK < n ? b + PAF_CHUNK : n;
    uint32_t k = off[c];
    for (uint64_t i = b; i < e; ++i) if (text[i] == '\n') line_start[++k] = i + 1;
    if (c == 0) line_start[0] = 0;
}
__device__ __forceinline__ void line_bounds(const char* text, uint64_t n, const unsigned long long* line_start, uint32_t n_lines, uint32_t l,
                                            const char** b, const char** e) {
    const uint64_t s = line_start[l];
    const uint64_t t = (l + 1 < n_lines) ? line_start[l + 1] - 1 : n;      // without the line feed; the piece after the last one may be empty
    *b = text + s; *e = text + t;
}
__global__ void __launch_bounds__(128) k0_scan_lines(const char* text, uint64_t n, const unsigned long long* line_start, uint32_t n_lines,
                                                     uint32_t* keep, uint32_t* nops, uint32_t* first_bad) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const char *b, *e;
    line_bounds(text, n, line_start, n_lines, l, &b, &e);
    PafLine ln;
    const bool row = paf_scan_line(b, e, &ln);
    if (row && ln.n_cols < 12) atomicMin(first_bad, l);
    keep[l] = row ? 1u : 0u;
    nops[l] = row ? ln.n_ops : 0u;
}
struct PafCols {
    uint32_t* c[10];      // q_id q_len q_start q_end t_id t_len t_start t_end n_match n_block
    uint8_t* is_rev; uint8_t* mapq;
    uint32_t* cg_off; uint32_t* cg_ops;
};
__global__ void __launch_bounds__(128) k0_emit(const char* text, uint64_t n, const unsigned long long* line_start, uint32_t n_lines,
                                               const uint32_t* keep, const uint32_t* row_of, const uint32_t* op_off, PafCols out) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines || !keep[l]) return;
    const char *b, *e;
    line_bounds(text, n, line_start, n_lines, l, &b, &e);
    PafLine ln;
    paf_scan_line(b, e, &ln);
    const uint32_t r = row_of[l];
    const int src[10] = {0, 1, 2, 3, 5, 6, 7, 8, 9, 10};
#pragma unroll
    for (int k = 0; k < 10; ++k) out.c[k][r] = paf_u32(b + ln.f[src[k]], b + ln.fe[src[k]]);
    out.is_rev[r] = (ln.fe[4] > ln.f[4] && b[ln.f[4]] == '-') ? 1 : 0;
    out.mapq[r] = (uint8_t)paf_u32(b + ln.f[11], b + ln.fe[11]);
    out.cg_off[r] = op_off[l];
    paf_emit_ops(b, ln, out.cg_ops + op_off[l]);
}


// read_off[r] = first row of read r (rows are grouped by ascending read id): every row that starts a new read id fills the
// entries of the ids it skips; flags[0] = first row whose id is smaller than its predecessor's, flags[1] = largest id seen
__global__ void __launch_bounds__(256) k0_read_off(const uint32_t* q_id, uint32_t n_rows, uint32_t n_reads, uint32_t* read_off, uint32_t* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_rows) return;
    if (i == n_rows) {                                         // the tail: reads after the last one that has rows
        const uint32_t last = n_rows ? q_id[n_rows - 1] : 0xFFFFFFFFu;
        for (uint32_t r = last + 1; r <= n_reads; ++r) read_off[r] = n_rows;      // last = ~0 wraps to 0: no rows at all
        if (n_rows) atomicMax(flags + 1, last);
        return;
    }
    const uint32_t id = q_id[i];
    if (i == 0) { for (uint32_t r = 0; r <= id && r <= n_reads; ++r) read_off[r] = 0; return; }
    const uint32_t pid = q_id[i - 1];
    if (id < pid) { atomicMin(flags, i); return; }
    for (uint32_t r = pid + 1; r <= id && r <= n_reads; ++r) read_off[r] = i;
}

static void publish_hits(hgpu_ctx* ctx, PafState* S) {
    ResidentHits& h = ctx->hits;
    h = ResidentHits();
    h.valid = true; h.n_hits = S->n_rows; h.n_ops = S->n_ops;
    h.q_id = S->col[0].p; h.q_start = S->col[2].p; h.q_end = S->col[3].p; h.t_id = S->col[4].p; h.t_len = S->col[5].p;
    h.t_start = S->col[6].p; h.t_end = S->col[7].p; h.n_match = S->col[8].p; h.n_block = S->col[9].p;
    h.is_rev = S->is_rev.p; h.mapq = S->mapq.p; h.cg_off = S->cg_off.p; h.cg_ops = S->cg_ops.p;
    ctx->compact = ResidentCompact();
}

extern "C" int hgpu_paf_tokenize(hgpu_t* ctx, const char* text, uint64_t n_bytes, uint64_t* out_n_rows, uint64_t* out_n_ops) {
    if (!ctx) return HGPU_E_INVALID;
    if (!out_n_rows || !out_n_ops || (n_bytes && !text)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    *out_n_rows = 0; *out_n_ops = 0;
    if (!ctx->paf) ctx->paf = new PafState();
    PafState* S = ctx->paf;
    S->ready = false; S->n_rows = 0; S->n_ops = 0; S->n_lines = 0;
    ctx->hits = ResidentHits(); ctx->compact = ResidentCompact();
    ctx->stage.ms_k0 = 0; ctx->stage.launches_k0 = 0; ctx->stage.k0_text_bytes = n_bytes; ctx->stage.k0_rows = 0; ctx->stage.k0_ops = 0;
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_bytes == 0) {
        for (int k = 0; k < 10; ++k) HGPU_CUDA(ctx, S->col[k].ensure(1));
        HGPU_CUDA(ctx, S->is_rev.ensure(1)); HGPU_CUDA(ctx, S->mapq.ensure(1)); HGPU_CUDA(ctx, S->cg_off.ensure(1)); HGPU_CUDA(ctx, S->cg_ops.ensure(1));
        HGPU_CUDA(ctx, cudaMemsetAsync(S->cg_off.p, 0, 4, ctx->stream));
        S->ready = true; publish_hits(ctx, S);
        return HGPU_OK;
    }
    if (n_bytes >= (1ull << 40)) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffers of a terabyte and more must be tokenised in pieces");
    cudaStream_t st = ctx->stream;
    const uint64_t chunks64 = (n_bytes + PAF_CHUNK - 1) / PAF_CHUNK;
    if (chunks64 > 0xFFFFFFF0ull) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffer too large for one call");
    const uint32_t n_chunks = (uint32_t)chunks64;
    HGPU_CUDA(ctx, S->text.ensure(n_bytes + 256)); HGPU_CUDA(ctx, S->chunk_cnt.ensure(n_chunks)); HGPU_CUDA(ctx, S->chunk_off.ensure(n_chunks + 1));
    HGPU_CUDA(ctx, S->scal.ensure(4)); HGPU_CUDA(ctx, S->totals.ensure(4));
    HGPU_CUDA(ctx, S->scan_tmp.ensure(scan_tmp_entries(n_chunks)));
    HGPU_H2D(ctx, S->text.p, text, n_bytes);
    stage_begin(ctx, ctx->ev_k0);
    k0_count_nl<<<(n_chunks + 255) / 256, 256, 0, st>>>(S->text.p, n_bytes, n_chunks, S->chunk_cnt.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    int launches = 1 + scan_u32(st, S->chunk_cnt.p, S->chunk_off.p, n_chunks, S->scan_tmp.p, S->totals.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    unsigned long long n_nl = 0;                              // totals are kept in 64 bits: offsets that would wrap are refused, not truncated
    HGPU_CUDA(ctx, cudaMemcpyAsync(&n_nl, S->totals.p, 8, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_nl >= 0xFFFFFFF0ull) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffer holds %llu lines: tokenise it in pieces of fewer than 2^32", n_nl);
    const uint32_t n_lines = (uint32_t)n_nl + 1;              // the piece after the last line feed counts (it may be empty)
    HGPU_CUDA(ctx, S->line_start.ensure((size_t)n_lines + 1));
    HGPU_CUDA(ctx, S->keep.ensure(n_lines)); HGPU_CUDA(ctx, S->row_of.ensure(n_lines + 1));
    HGPU_CUDA(ctx, S->nops.ensure(n_lines)); HGPU_CUDA(ctx, S->line_cg_off.ensure(n_lines + 1));
    k0_line_start<<<(n_chunks + 255) / 256, 256, 0, st>>>(S->text.p, n_bytes, n_chunks, S->chunk_off.p, S->line_start.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    HGPU_CUDA(ctx, cudaMemsetAsync(S->scal.p, 0xFF, 4, st));
    k0_scan_lines<<<(n_lines + 127) / 128, 128, 0, st>>>(S->text.p, n_bytes, S->line_start.p, n_lines, S->keep.p, S->nops.p, S->scal.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    HGPU_CUDA(ctx, S->scan_tmp.ensure(scan_tmp_entries(n_lines)));
    launches += 2 + scan_u32(st, S->keep.p, S->row_of.p, n_lines, S->scan_tmp.p, S->totals.p + 1);
    HGPU_CUDA(ctx, cudaGetLastError());
    launches += scan_u32(st, S->nops.p, S->line_cg_off.p, n_lines, S->scan_tmp.p, S->totals.p + 2);       // per LINE; k0_emit copies it per row
    HGPU_CUDA(ctx, cudaGetLastError());
    uint32_t bad = 0;
    unsigned long long tot[2] = {0, 0};
    HGPU_CUDA(ctx, cudaMemcpyAsync(&bad, S->scal.p, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(tot, S->totals.p + 1, 16, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += launches; ctx->stage.launches_k0 = launches;
    if (bad != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "PAF line %u has fewer than 12 columns", bad + 1);
    if (const char* e = getenv("HGPU_TEST_PAF_OPS_BIAS")) tot[1] += strtoull(e, nullptr, 10);   // tests: pretend the table is this much larger
    if (tot[1] > 0xFFFFFFFFull)      // cg_off is 32-bit (hgpu_hits_t): a table this large must come in several calls
        HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffer holds %llu CIGAR runs (limit 2^32 - 1 per call): tokenise it in pieces", tot[1]);
    const uint32_t rows = (uint32_t)tot[0], ops = (uint32_t)tot[1];
    // the table itself: it stays on the device for hgpu_compact_lr_dev / hgpu_edge_coords_dev; hgpu_paf_fetch copies it out
    PafCols out;
    for (int k = 0; k < 10; ++k) { HGPU_CUDA(ctx, S->col[k].ensure((size_t)rows + 1)); out.c[k] = S->col[k].p; }
    HGPU_CUDA(ctx, S->is_rev.ensure((size_t)rows + 1)); HGPU_CUDA(ctx, S->mapq.ensure((size_t)rows + 1));
    HGPU_CUDA(ctx, S->cg_off.ensure((size_t)rows + 1)); HGPU_CUDA(ctx, S->cg_ops.ensure((size_t)ops + 1));
    out.is_rev = S->is_rev.p; out.mapq = S->mapq.p; out.cg_off = S->cg_off.p; out.cg_ops = S->cg_ops.p;
    k0_emit<<<(n_lines + 127) / 128, 128, 0, st>>>(S->text.p, n_bytes, S->line_start.p, n_lines, S->keep.p, S->row_of.p, S->line_cg_off.p, out);
    HGPU_CUDA(ctx, cudaGetLastError());
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->cg_off.p + rows, &ops, 4, cudaMemcpyHostToDevice, st));
    stage_end(ctx, ctx->ev_k0);
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1; ctx->stage.launches_k0 += 1;
    ctx->stage.ms_k0 = stage_ms(ctx, ctx->ev_k0);
    ctx->stage.k0_rows = rows; ctx->stage.k0_ops = ops;
    S->n_lines = n_lines; S->n_rows = rows; S->n_ops = ops; S->n_bytes = n_bytes; S->ready = true;
    publish_hits(ctx, S);
    *out_n_rows = rows; *out_n_ops = ops;
    return HGPU_OK;
}

extern "C" int hgpu_paf_fetch(hgpu_t* ctx, uint32_t* q_id, uint32_t* q_len, uint32_t* q_start, uint32_t* q_end, uint8_t* is_rev,
                              uint32_t* t_id, uint32_t* t_len, uint32_t* t_start, uint32_t* t_end, uint32_t* n_match, uint32_t* n_block,
                              uint8_t* mapq, uint32_t* cg_off, uint32_t* cg_ops) {
    if (!ctx) return HGPU_E_INVALID;
    PafState* S = ctx->paf;
    if (!S || !S->ready) HGPU_FAIL(ctx, HGPU_E_INVALID, "hgpu_paf_fetch without a successful hgpu_paf_tokenize");
    if (!cg_off) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    cg_off[0] = 0;
    const uint32_t rows = S->n_rows;
    if (rows == 0) return HGPU_OK;
    uint32_t* host_cols[10] = {q_id, q_len, q_start, q_end, t_id, t_len, t_start, t_end, n_match, n_block};
    for (int k = 0; k < 10; ++k) if (!host_cols[k]) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (!is_rev || !mapq || (S->n_ops && !cg_ops)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int k = 0; k < 10; ++k) HGPU_D2H(ctx, host_cols[k], S->col[k].p, (size_t)rows * 4);
    HGPU_D2H(ctx, is_rev, S->is_rev.p, rows);
    HGPU_D2H(ctx, mapq, S->mapq.p, rows);
    HGPU_D2H(ctx, cg_off, S->cg_off.p, ((size_t)rows + 1) * 4);
    HGPU_D2H(ctx, cg_ops, S->cg_ops.p, (size_t)S->n_ops * 4);
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HGPU_OK;
}

// Group the resident hit table by read (rows must be grouped by ascending read id - Longread.cpp:57-84 slices the hit array
// by cumulative counts in read-id order and silently mis-assigns hits otherwise; here it is an error).
extern "C" int hgpu_hits_group(hgpu_t* ctx, uint32_t n_reads, uint32_t* out_read_off) {
    if (!ctx) return HGPU_E_INVALID;
    ResidentHits& h = ctx->hits;
    if (!h.valid || !ctx->paf || !h.q_id) HGPU_FAIL(ctx, HGPU_E_INVALID, "hgpu_hits_group needs the hit table hgpu_paf_tokenize leaves on the device");
    PafState* S = ctx->paf;
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    HGPU_CUDA(ctx, S->read_off.ensure((size_t)n_reads + 2)); HGPU_CUDA(ctx, S->scal.ensure(4));
    const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->scal.p, init, 8, cudaMemcpyHostToDevice, st));
    k0_read_off<<<(h.n_hits + 1 + 255) / 256, 256, 0, st>>>(h.q_id, h.n_hits, n_reads, S->read_off.p, S->scal.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1; ctx->stage.launches_k0 += 1;
    uint32_t flags[2];
    HGPU_CUDA(ctx, cudaMemcpyAsync(flags, S->scal.p, 8, cudaMemcpyDeviceToHost, st));
    if (out_read_off) HGPU_D2H(ctx, out_read_off, S->read_off.p, ((size_t)n_reads + 1) * 4);
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (flags[0] != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "PAF rows are not grouped by ascending read id (row %u)", flags[0]);
    if (h.n_hits && flags[1] >= n_reads) HGPU_FAIL(ctx, HGPU_E_INVALID, "PAF names read %u but only %u reads were loaded", flags[1], n_reads);
    h.read_off = S->read_off.p; h.n_reads = n_reads; h.grouped = true;
    return HGPU_OK;
}
