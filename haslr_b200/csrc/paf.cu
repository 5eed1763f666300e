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
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "paf_core.cuh"
#include "scan.cuh"

using namespace hgpu;

struct PafState {
    DevBuf<char> text;
    DevBuf<uint32_t> chunk_cnt, chunk_off, keep, row_of, nops, cg_off, cg_ops, col[10], scal;
    DevBuf<unsigned long long> line_start;
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

extern "C" int hgpu_paf_tokenize(hgpu_t* ctx, const char* text, uint64_t n_bytes, uint64_t* out_n_rows, uint64_t* out_n_ops) {
    if (!ctx) return HGPU_E_INVALID;
    if (!out_n_rows || !out_n_ops || (n_bytes && !text)) HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    *out_n_rows = 0; *out_n_ops = 0;
    if (!ctx->paf) ctx->paf = new PafState();
    PafState* S = ctx->paf;
    S->ready = false; S->n_rows = 0; S->n_ops = 0; S->n_lines = 0;
    if (n_bytes == 0) { S->ready = true; return HGPU_OK; }
    if (n_bytes >= (1ull << 40)) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffers of a terabyte and more must be tokenised in pieces");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t chunks64 = (n_bytes + PAF_CHUNK - 1) / PAF_CHUNK;
    if (chunks64 > 0xFFFFFFF0ull) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "PAF buffer too large for one call");
    const uint32_t n_chunks = (uint32_t)chunks64;
    HGPU_CUDA(ctx, S->text.ensure(n_bytes + 256)); HGPU_CUDA(ctx, S->chunk_cnt.ensure(n_chunks)); HGPU_CUDA(ctx, S->chunk_off.ensure(n_chunks + 1));
    HGPU_CUDA(ctx, S->scal.ensure(4));
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->text.p, text, n_bytes, cudaMemcpyHostToDevice, st));
    k0_count_nl<<<(n_chunks + 255) / 256, 256, 0, st>>>(S->text.p, n_bytes, n_chunks, S->chunk_cnt.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->chunk_cnt.p, S->chunk_off.p, n_chunks);
    HGPU_CUDA(ctx, cudaGetLastError());
    uint32_t n_nl = 0;
    HGPU_CUDA(ctx, cudaMemcpyAsync(&n_nl, S->chunk_off.p + n_chunks, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t n_lines = n_nl + 1;                        // the piece after the last line feed counts (it may be empty)
    HGPU_CUDA(ctx, S->line_start.ensure((size_t)n_lines + 1));
    HGPU_CUDA(ctx, S->keep.ensure(n_lines)); HGPU_CUDA(ctx, S->row_of.ensure(n_lines + 1));
    HGPU_CUDA(ctx, S->nops.ensure(n_lines)); HGPU_CUDA(ctx, S->cg_off.ensure(n_lines + 1));
    k0_line_start<<<(n_chunks + 255) / 256, 256, 0, st>>>(S->text.p, n_bytes, n_chunks, S->chunk_off.p, S->line_start.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    HGPU_CUDA(ctx, cudaMemsetAsync(S->scal.p, 0xFF, 4, st));
    k0_scan_lines<<<(n_lines + 127) / 128, 128, 0, st>>>(S->text.p, n_bytes, S->line_start.p, n_lines, S->keep.p, S->nops.p, S->scal.p);
    HGPU_CUDA(ctx, cudaGetLastError());
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->keep.p, S->row_of.p, n_lines);
    HGPU_CUDA(ctx, cudaGetLastError());
    k_exclusive_scan<<<1, 1024, 0, st>>>(S->nops.p, S->cg_off.p, n_lines);       // per LINE; k0_emit copies it per row
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 6;
    uint32_t bad = 0, rows = 0, ops = 0;
    HGPU_CUDA(ctx, cudaMemcpyAsync(&bad, S->scal.p, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(&rows, S->row_of.p + n_lines, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(&ops, S->cg_off.p + n_lines, 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (bad != 0xFFFFFFFFu) HGPU_FAIL(ctx, HGPU_E_INVALID, "PAF line %u has fewer than 12 columns", bad + 1);
    S->n_lines = n_lines; S->n_rows = rows; S->n_ops = ops; S->n_bytes = n_bytes; S->ready = true;
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
    cudaStream_t st = ctx->stream;
    PafCols out;
    for (int k = 0; k < 10; ++k) { HGPU_CUDA(ctx, S->col[k].ensure(rows)); out.c[k] = S->col[k].p; }
    HGPU_CUDA(ctx, S->is_rev.ensure(rows)); HGPU_CUDA(ctx, S->mapq.ensure(rows));
    DevBuf<uint32_t> row_cg;                                  // per-row offsets (cg_off above is per line)
    HGPU_CUDA(ctx, row_cg.alloc((size_t)rows + 1));
    HGPU_CUDA(ctx, S->cg_ops.ensure(S->n_ops + 1));
    out.is_rev = S->is_rev.p; out.mapq = S->mapq.p; out.cg_off = row_cg.p; out.cg_ops = S->cg_ops.p;
    k0_emit<<<(S->n_lines + 127) / 128, 128, 0, st>>>(S->text.p, S->n_bytes, S->line_start.p, S->n_lines, S->keep.p, S->row_of.p, S->cg_off.p, out);
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    for (int k = 0; k < 10; ++k) HGPU_CUDA(ctx, cudaMemcpyAsync(host_cols[k], S->col[k].p, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(is_rev, S->is_rev.p, rows, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(mapq, S->mapq.p, rows, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaMemcpyAsync(cg_off, row_cg.p, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    if (S->n_ops) HGPU_CUDA(ctx, cudaMemcpyAsync(cg_ops, S->cg_ops.p, (size_t)S->n_ops * 4, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    cg_off[rows] = (uint32_t)S->n_ops;
    return HGPU_OK;
}
