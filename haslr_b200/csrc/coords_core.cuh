// Per-edge core of the edge-coordinate stage, __host__ __device__ so the CPU tests can exercise it.
//
// Replaces, for ONE backbone edge: asm_best_supported_interval_contig1/2 (reference
// src/haslr_assemble/src/Assemble.cpp:24-126), asm_find_lr_pos (:129-155) and the eight strand cases of
// asm_calc_single_edge_coordinates (:157-363). The reference sweeps sorted begin/end lists while copying a std::set of
// supporting reads at every new optimum, and walks CIGARs expanded to one character per operation; here the sweep
// runs once to find WHERE the last optimum is taken and is replayed up to that point on a bitmask, and the walks
// go run by run over the run-length CIGAR window an element carries (hgpu_cl_elem::cg_lo..cg_hi).
#pragma once
#include <stdint.h>

#include "../../include/haslr_b200.h"

#if defined(__CUDACC__)
#define K4_HD __host__ __device__ __forceinline__
#else
#define K4_HD inline
#endif

namespace hgpu {

struct CoordIn {
    const hgpu_edge_supp* supp;
    const hgpu_cl_elem* elems;
    const uint32_t* cl_read_off;
    const uint32_t* read_len;
    const uint8_t* hit_is_rev;
    const uint32_t* cg_off;
    const uint32_t* cg_ops;
};

K4_HD const hgpu_cl_elem& k4_elem(const CoordIn& in, const hgpu_edge_supp& s, bool head) {
    return in.elems[in.cl_read_off[s.lr_id_strand & 0x7FFFFFFFu] + (head ? s.cmp_head : s.cmp_tail)];
}

// sort key of a (position, support index) pair: the order of std::sort on vector<pair<uint32_t, uint32_t>>
K4_HD uint64_t k4_key(uint32_t pos, uint32_t idx) { return ((uint64_t)pos << 32) | idx; }

// The sweep of Assemble.cpp:39-69 over SORTED keys. ge: a later interval of equal depth replaces the optimum (contig1,
// `>=`) or not (contig2, `>`). mask: ceil(n / 32) zeroed words; on return bit k is set iff support k is in best_lrs.
K4_HD void k4_best_interval(const uint64_t* beg, const uint64_t* end, uint32_t n, bool ge, uint32_t* mask, uint32_t* lo, uint32_t* hi) {
    // pass 1: position of the last optimum (i, j at the moment it is taken) and the interval
    int cur = 0, top = 0;
    uint32_t i = 0, j = 0, si = 0xFFFFFFFFu, sj = 0, blo = 0, bhi = 0;
    bool open = false;
    while (i < n && j < n) {
        if ((uint32_t)(beg[i] >> 32) < (uint32_t)(end[j] >> 32)) {
            ++cur;
            if (ge ? cur >= top : cur > top) { top = cur; blo = (uint32_t)(beg[i] >> 32); si = i; sj = j; open = true; }
            ++i;
        } else {
            if (open) { bhi = (uint32_t)(end[j] >> 32); open = false; }
            --cur;
            ++j;
        }
    }
    if (open) bhi = (uint32_t)(end[j] >> 32);
    *lo = blo; *hi = bhi;
    if (si == 0xFFFFFFFFu) return;                   // no support at all
    // pass 2: the set as it stood then — the same inserts and erases, replayed on the bitmask
    i = 0; j = 0;
    while (i < n && j < n) {
        if ((uint32_t)(beg[i] >> 32) < (uint32_t)(end[j] >> 32)) {
            const uint32_t k = (uint32_t)beg[i];
            mask[k >> 5] |= 1u << (k & 31u);
            if (i == si && j == sj) break;
            ++i;
        } else {
            const uint32_t k = (uint32_t)end[j];
            mask[k >> 5] &= ~(1u << (k & 31u));
            ++j;
        }
    }
}

// asm_find_lr_pos (Assemble.cpp:129-155) on the element's run-length CIGAR window: walk until the contig coordinate
// reaches contig_pos. reversed = walk the window from its last run. Returns the read coordinate, or -1.
K4_HD int64_t k4_find_lr_pos(const CoordIn& in, const hgpu_cl_elem& e, bool reversed, uint32_t lr_curr, uint32_t c_curr,
                             int lr_step, int c_step, uint32_t contig_pos) {
    if ((c_step > 0 && c_curr > contig_pos) || (c_step < 0 && c_curr < contig_pos)) return -1;
    uint32_t dist = c_step > 0 ? contig_pos - c_curr : c_curr - contig_pos;     // contig steps still to go
    const uint32_t b = in.cg_off[e.hit];
    if (in.cg_off[e.hit + 1] == b) return (int64_t)lr_curr;                     // PAF row without cg:Z: — nothing to walk
    const uint32_t* ops = in.cg_ops + b;
    const uint32_t n = e.cg_hi - e.cg_lo + 1;
    for (uint32_t k = 0; k < n && dist > 0; ++k) {
        const uint32_t r = reversed ? e.cg_hi - k : e.cg_lo + k;
        const uint32_t op = ops[r] & 3u;
        const uint32_t len = (r == e.cg_lo) ? e.cg_lo_len : (r == e.cg_hi ? e.cg_hi_len : ops[r] >> 2);
        if (op == 0) {                     // M: both move
            const uint32_t take = len < dist ? len : dist;
            lr_curr += (uint32_t)lr_step * take; dist -= take;
        } else if (op == 1) {              // I: only the read moves
            lr_curr += (uint32_t)lr_step * len;
        } else {                           // D / anything else: only the contig moves
            dist -= len < dist ? len : dist;
        }
    }
    return (int64_t)lr_curr;
}

// One member of the best set: which stretch of the read lies between contig1_pos and contig2_pos (Assemble.cpp:255-338)
K4_HD void k4_walk(const CoordIn& in, const hgpu_edge_supp& s, uint32_t rev1, uint32_t rev2, uint32_t c1, uint32_t c2, hgpu_supp_coord* out) {
    const uint32_t rlen = in.read_len[s.lr_id_strand & 0x7FFFFFFFu];
    const hgpu_cl_elem& a1 = k4_elem(in, s, true);
    const hgpu_cl_elem& a2 = k4_elem(in, s, false);
    const uint32_t rstrand = (rev1 == in.hit_is_rev[a1.hit]) ? 0u : 1u;
    const uint32_t q0h = rstrand == 0 ? a1.q_start : rlen - a1.q_end;
    const uint32_t q0t = rstrand == 0 ? a2.q_end - 1 : rlen - a2.q_start - 1;
    out->lr_start = rev1 == 0 ? k4_find_lr_pos(in, a1, false, q0h, a1.t_start, +1, +1, c1)       // cases 1 / 5
                              : k4_find_lr_pos(in, a1, true, q0h, a1.t_end - 1, +1, -1, c1);     // cases 2 / 6
    out->lr_end = rev2 == 0 ? k4_find_lr_pos(in, a2, true, q0t, a2.t_end - 1, -1, -1, c2)        // cases 3 / 7
                            : k4_find_lr_pos(in, a2, false, q0t, a2.t_start, -1, +1, c2);        // cases 4 / 8
    out->lr_strand = rstrand;
    out->in_best = 1;
}

}  // namespace hgpu
