// TEST-ONLY host harness for haslr_b200/csrc/paf_core.cuh (the per-line core is __host__ __device__): the passes of
// paf.cu run serially — split at line feeds, scan every line, emit rows and run-length CIGARs. Never shipped.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../haslr_b200/csrc/paf_core.cuh"

using namespace hgpu;

extern "C" int64_t pafhost_parse(const char* text, uint64_t n, uint64_t row_cap, uint64_t op_cap,
                                 uint32_t* q_id, uint32_t* q_len, uint32_t* q_start, uint32_t* q_end, uint8_t* is_rev,
                                 uint32_t* t_id, uint32_t* t_len, uint32_t* t_start, uint32_t* t_end, uint32_t* n_match,
                                 uint32_t* n_block, uint8_t* mapq, uint32_t* cg_off, uint32_t* cg_ops, uint64_t* n_ops_out) {
    uint64_t rows = 0, ops = 0;
    cg_off[0] = 0;
    uint64_t s = 0;
    while (s <= n) {
        const char* nl = s < n ? (const char*)memchr(text + s, '\n', n - s) : nullptr;
        const uint64_t t = nl ? (uint64_t)(nl - text) : n;
        PafLine ln;
        if (paf_scan_line(text + s, text + t, &ln)) {
            if (ln.n_cols < 12) return -2;
            if (rows >= row_cap || ops + ln.n_ops > op_cap) return -1;
            const char* b = text + s;
            uint32_t* c[10] = {q_id, q_len, q_start, q_end, t_id, t_len, t_start, t_end, n_match, n_block};
            const int src[10] = {0, 1, 2, 3, 5, 6, 7, 8, 9, 10};
            for (int k = 0; k < 10; ++k) c[k][rows] = paf_u32(b + ln.f[src[k]], b + ln.fe[src[k]]);
            is_rev[rows] = (ln.fe[4] > ln.f[4] && b[ln.f[4]] == '-') ? 1 : 0;
            mapq[rows] = (uint8_t)paf_u32(b + ln.f[11], b + ln.fe[11]);
            paf_emit_ops(b, ln, cg_ops + ops);
            ops += ln.n_ops;
            cg_off[++rows] = (uint32_t)ops;
        }
        if (!nl) break;
        s = t + 1;
    }
    *n_ops_out = ops;
    return (int64_t)rows;
}
