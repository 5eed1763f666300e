// Per-line core of the PAF tokeniser, __host__ __device__ so the CPU tests can exercise it.
//
// Replaces the text side of load_alignment (reference src/haslr_assemble/src/Longread.cpp:250-289): getline,
// str_split on tabs, str2type<uint32_t> on columns 1-4 and 6-12, the strand character of column 5, and the search for
// the first "cg:Z:" tag among the optional columns. The reference keeps the CIGAR as a string and expands it later
// (expand_cigar, Common.cpp); here it becomes run-length operations (len << 2) | op, op 0 = M, 1 = I, 2 = anything
// else, straight away. Every row is kept: the load filters F1-F4 run in hgpu_compact_lr.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PAF_HD __host__ __device__ __forceinline__
#else
#define PAF_HD inline
#endif

namespace hgpu {

struct PafLine {
    uint32_t f[12];      // byte offsets (from the line start) of columns 1..12
    uint32_t fe[12];     // ... and of their ends
    uint32_t cg_b, cg_e; // the cg:Z: payload, cg_b == cg_e if the row has none
    uint32_t n_ops;      // run-length operations in it
    uint32_t n_cols;     // tab-separated columns found among the first 12
};

// A column as `istringstream >> uint32_t` (str2type<uint32_t>, Common.hpp:126-133) reads it: leading white space skipped,
// an optional sign, decimal digits up to the first other character. No digits -> 0 (the reference returns an
// uninitialised value there); a magnitude above 2^32 - 1 -> 4294967295; a '-' negates modulo 2^32 (libstdc++ num_get).
PAF_HD uint32_t paf_u32(const char* b, const char* e) {
    while (b < e && (*b == ' ' || (*b >= '\t' && *b <= '\r'))) ++b;
    bool neg = false;
    if (b < e && (*b == '+' || *b == '-')) { neg = *b == '-'; ++b; }
    uint64_t v = 0;
    bool over = false;
    for (; b < e && *b >= '0' && *b <= '9'; ++b) {
        v = v * 10 + (uint64_t)(*b - '0');
        if (v > 0xFFFFFFFFull) { over = true; v = 0xFFFFFFFFull; }
    }
    if (over) return 0xFFFFFFFFu;
    return neg ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v;
}

// Scans one line [b, e) (no terminator). Returns false for an empty line (skipped) — and with n_cols < 12 for a
// malformed one (the caller refuses the file; the reference would index past its fields vector).
PAF_HD bool paf_scan_line(const char* b, const char* e, PafLine* out) {
    out->n_cols = 0; out->cg_b = out->cg_e = 0; out->n_ops = 0;
    if (b == e) return false;
    const uint32_t len = (uint32_t)(e - b);
    uint32_t nf = 1, p = 0, opt = len;           // opt = offset of column 13, if any
    out->f[0] = 0;
    for (; p < len; ++p) {
        if (b[p] != '\t') continue;
        out->fe[nf - 1] = p;
        if (nf == 12) { opt = p + 1; break; }
        out->f[nf++] = p + 1;
    }
    if (p == len) out->fe[nf - 1] = len;         // the last column runs to the end of the line
    out->n_cols = nf;
    if (nf < 12) return true;
    // first cg:Z: tag among the optional columns (Longread.cpp:275-283)
    uint32_t q = opt;
    while (q < len) {
        uint32_t t = q;
        while (t < len && b[t] != '\t') ++t;
        if (t - q >= 5 && b[q] == 'c' && b[q + 1] == 'g' && b[q + 2] == ':' && b[q + 3] == 'Z' && b[q + 4] == ':') {
            out->cg_b = q + 5; out->cg_e = t;
            uint32_t c = q + 5, n = 0;
            while (c < t) {
                while (c < t && b[c] >= '0' && b[c] <= '9') ++c;
                if (c >= t) break;
                ++c; ++n;
            }
            out->n_ops = n;
            break;
        }
        q = t + 1;
    }
    return true;
}

// run-length operations of the payload [b + cg_b, b + cg_e) into ops[0 .. n_ops)
PAF_HD void paf_emit_ops(const char* b, const PafLine& ln, uint32_t* ops) {
    uint32_t c = ln.cg_b, k = 0;
    while (c < ln.cg_e) {
        uint32_t n = 0;
        while (c < ln.cg_e && b[c] >= '0' && b[c] <= '9') n = n * 10 + (uint32_t)(b[c++] - '0');
        if (c >= ln.cg_e) break;
        const char op = b[c++];
        ops[k++] = (n << 2) | (op == 'M' ? 0u : op == 'I' ? 1u : 2u);
    }
}

}  // namespace hgpu
