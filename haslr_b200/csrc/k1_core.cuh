// Per-read core of the compact-long-read stage, __host__ __device__ so the CPU tests can exercise it.
//
// Replaces, for ONE long read: the per-read std::sort of load_alignment (reference
// src/haslr_assemble/src/Longread.cpp:256), process_lr_alignment_group (Longread.cpp:182-232),
// fix_overlapping_alignments / find_contig_pos (Longread.cpp:375-512) and find_best_scheduling
// (Longread.cpp:514-610). CIGARs stay run-length encoded: the reference expands them to one char per op
// and edits strings; here a hit carries a window [lo, hi) over its expanded ops and walks run by run.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define K1_HD __host__ __device__ __forceinline__
#else
#define K1_HD inline
#endif

namespace hgpu {

// ---------------------------------------------------------------------------------------------------------
// std::sort as libstdc++ implements it (introsort: median-of-3 + unguarded partition above 16 elements,
// heapsort at depth 2*floor(log2 n), final insertion sort), on an index array. The reference's comparator
// looks at (q_end, q_start) only, so the order of equal keys is whatever this exact algorithm leaves
//  — restating the algorithm is what keeps ties bit-exact.
// ---------------------------------------------------------------------------------------------------------
struct KeyLess {
    const uint32_t* q_end; const uint32_t* q_start;
    K1_HD bool operator()(uint32_t a, uint32_t b) const {
        return (q_end[a] < q_end[b]) || (q_end[a] == q_end[b] && q_start[a] < q_start[b]);
    }
};

template <typename Less>
K1_HD void ls_unguarded_linear_insert(uint32_t* v, int last, const Less& lt) {
    uint32_t val = v[last];
    int next = last - 1;
    while (lt(val, v[next])) { v[last] = v[next]; last = next; --next; }
    v[last] = val;
}
template <typename Less>
K1_HD void ls_insertion_sort(uint32_t* v, int first, int last, const Less& lt) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (lt(v[i], v[first])) {
            uint32_t val = v[i];
            for (int k = i; k > first; --k) v[k] = v[k - 1];
            v[first] = val;
        } else {
            ls_unguarded_linear_insert(v, i, lt);
        }
    }
}
template <typename Less>
K1_HD void ls_push_heap(uint32_t* v, int first, int hole, int top, uint32_t value, const Less& lt) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt(v[first + parent], value)) { v[first + hole] = v[first + parent]; hole = parent; parent = (hole - 1) / 2; }
    v[first + hole] = value;
}
template <typename Less>
K1_HD void ls_adjust_heap(uint32_t* v, int first, int hole, int len, uint32_t value, const Less& lt) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(v[first + child], v[first + child - 1])) child--;
        v[first + hole] = v[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[first + hole] = v[first + child - 1];
        hole = child - 1;
    }
    ls_push_heap(v, first, hole, top, value, lt);
}
template <typename Less>
K1_HD void ls_heap_sort(uint32_t* v, int first, int last, const Less& lt) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            uint32_t val = v[first + parent];
            ls_adjust_heap(v, first, parent, len, val, lt);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) {
        --l;
        uint32_t val = v[l];
        v[l] = v[first];
        ls_adjust_heap(v, first, 0, l - first, val, lt);
    }
}
template <typename Less>
K1_HD void libstdcxx_sort(uint32_t* v, int n, const Less& lt) {
    if (n <= 1) return;
    // explicit stack instead of recursion: sub-ranges are disjoint, so the processing order is immaterial
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    stk_first[sp] = 0; stk_last[sp] = n; stk_depth[sp] = 2 * lg; ++sp;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { ls_heap_sort(v, first, last, lt); break; }
            --depth;
            // median of (first+1, mid, last-1) moved to first
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            int m;
            if (lt(v[a], v[b])) { if (lt(v[b], v[c])) m = b; else if (lt(v[a], v[c])) m = c; else m = a; }
            else if (lt(v[a], v[c])) m = a;
            else if (lt(v[b], v[c])) m = c;
            else m = b;
            { uint32_t t = v[first]; v[first] = v[m]; v[m] = t; }
            int lo = first + 1, hi = last;
            const uint32_t pivot = v[first];
            while (true) {
                while (lt(v[lo], pivot)) ++lo;
                --hi;
                while (lt(pivot, v[hi])) --hi;
                if (!(lo < hi)) break;
                uint32_t t = v[lo]; v[lo] = v[hi]; v[hi] = t;
                ++lo;
            }
            // right part [lo, last) deferred, continue with the left part
            if (sp < 64) { stk_first[sp] = lo; stk_last[sp] = last; stk_depth[sp] = depth; ++sp; }
            last = lo;
        }
    }
    if (n > 16) {
        ls_insertion_sort(v, 0, 16, lt);
        for (int i = 16; i != n; ++i) ls_unguarded_linear_insert(v, i, lt);
    } else {
        ls_insertion_sort(v, 0, n, lt);
    }
}

// ---------------------------------------------------------------------------------------------------------
// One hit of the group while it is being trimmed and chained.
// ---------------------------------------------------------------------------------------------------------
struct K1Hit {
    uint32_t src;                     // row in the hit columns
    uint32_t q_start, q_end, t_start, t_end, n_match, n_block;
    uint32_t lo, hi;                  // kept window over the expanded CIGAR ops, [lo, hi)
    uint32_t is_rev;
};

// Run-length CIGAR of one hit: ops[k] = (len << 2) | op, op 0 = M, 1 = I, other = D (Longread.cpp:384-396).
struct RleCigar { const uint32_t* ops; uint32_t n; };

// find_contig_pos (Longread.cpp:375-420) on the window [h.lo, h.hi) of a run-length CIGAR.
// forward = walk from the window's low end upwards, else from its high end downwards.
// Walk: before each op test lr_curr == lr_pos (stop); M moves both, I moves the read, anything else the contig.
// Then back up to the nearest M at or before the stop position (an index past the end counts as "not M", Q3)
// and keep the ops up to and including it. Returns the number of ops kept; *m_kept = how many of them are M.
K1_HD uint32_t rle_find_contig_pos(const RleCigar& cg, uint32_t lo, uint32_t hi, bool forward,
                                   uint32_t& lr_curr, uint32_t& c_curr, int lr_step, int c_step, uint32_t lr_pos,
                                   uint32_t* m_kept) {
    const uint32_t n = hi - lo;       // ops in the window
    *m_kept = 0;
    if (n == 0) return 0;             // reference: empty string, nothing walked, nothing erased
    const int dir = forward ? 1 : -1;
    // run holding the first op of the walk, and how many of its ops lie inside the window
    int r0; uint32_t len0;
    if (forward) {
        uint32_t pos = 0; r0 = 0;
        while (pos + (cg.ops[r0] >> 2) <= lo) { pos += cg.ops[r0] >> 2; ++r0; }
        len0 = pos + (cg.ops[r0] >> 2) - lo;
    } else {
        uint32_t pos = 0;
        for (uint32_t k = 0; k < cg.n; ++k) pos += cg.ops[k] >> 2;   // expanded end of run r0
        r0 = (int)cg.n - 1;
        while (pos - (cg.ops[r0] >> 2) >= hi) { pos -= cg.ops[r0] >> 2; --r0; }
        len0 = hi - (pos - (cg.ops[r0] >> 2));
    }
    // cursor = (k-th run in walking order, c ops of it consumed); `start` = ops of the window before run k
    int k = 0;
    uint32_t start = 0, c = 0, i = 0, m_cnt = 0;
    uint32_t len = len0 < n ? len0 : n;
    while (i < n) {
        if (lr_curr == lr_pos) break;
        const uint32_t op = cg.ops[r0 + dir * k] & 3u;
        uint32_t take = len - c;
        bool stop = false;
        if (op <= 1) {                                       // M or I: the read coordinate moves
            const uint32_t dist = lr_step > 0 ? lr_pos - lr_curr : lr_curr - lr_pos;   // ops until equality (mod 2^32)
            if (dist < take) { take = dist; stop = true; }
            lr_curr += (uint32_t)lr_step * take;
            if (op == 0) { c_curr += (uint32_t)c_step * take; m_cnt += take; }
        } else {
            c_curr += (uint32_t)c_step * take;
        }
        i += take; c += take;
        if (c == len && i < n) {
            start += len; ++k; c = 0;
            len = cg.ops[r0 + dir * k] >> 2;
            if (len > n - start) len = n - start;
        }
        if (stop) break;
    }
    // back up to the nearest M at or before op[i]; an index past the end counts as "not M" (Q3)
    while (i > 0 && !(i < n && (cg.ops[r0 + dir * k] & 3u) == 0)) {
        if (c == 0) {
            --k;
            len = k == 0 ? len0 : (cg.ops[r0 + dir * k] >> 2);
            start -= len; c = len;
        }
        --c; --i;
        const uint32_t op = cg.ops[r0 + dir * k] & 3u;
        if (op == 0) { c_curr -= (uint32_t)c_step; lr_curr -= (uint32_t)lr_step; --m_cnt; }
        else if (op == 1) { lr_curr -= (uint32_t)lr_step; }
        else { c_curr -= (uint32_t)c_step; }
    }
    // ops [0, i] are kept
    *m_kept = m_cnt + (((cg.ops[r0 + dir * k] & 3u) == 0) ? 1u : 0u);
    return i + 1;
}

// Column view of the PAF hits (device or host pointers) — same fields as hgpu_hits_t.
struct HitCols {
    const uint32_t *q_start, *q_end, *t_id, *t_len, *t_start, *t_end, *n_match, *n_block;
    const uint8_t *is_rev, *mapq;
    const uint32_t *cg_off, *cg_ops;
};

struct K1Params {
    double min_aln_sim, uniq_freq, max_uniq_dev;
    uint32_t min_aln_block, min_aln_mapq;
};

// One compact-long-read element (layout of hgpu_cl_elem / oracle_cl_elem: 11 x uint32)
struct ClElem {
    uint32_t hit, q_start, q_end, t_start, t_end, n_match, n_block, cg_lo, cg_lo_len, cg_hi, cg_hi_len;
};

// load filters F1-F4 (Longread.cpp:262-272) on one PAF row
K1_HD bool k1_load_filter(const HitCols& h, uint32_t i, const double* mean_kmer, const K1Params& p) {
    if (h.n_block[i] < p.min_aln_block) return false;                                          // F1
    if ((double)h.n_match[i] / (double)h.n_block[i] < p.min_aln_sim) return false;            // F2
    if ((uint32_t)h.mapq[i] < p.min_aln_mapq) return false;                                    // F3
    if (mean_kmer[h.t_id[i]] > p.uniq_freq * (3 + p.max_uniq_dev)) return false;               // F4
    return true;
}

// kept expanded window [lo, hi) -> run window of the hit's CIGAR (first/last run and their clipped lengths)
K1_HD void k1_cigar_window(const RleCigar& cg, uint32_t lo, uint32_t hi, ClElem& e) {
    e.cg_lo = e.cg_hi = 0; e.cg_lo_len = e.cg_hi_len = 0;
    uint32_t pos = 0;
    bool have_lo = false;
    for (uint32_t k = 0; k < cg.n; ++k) {
        const uint32_t len = cg.ops[k] >> 2, end = pos + len;
        if (!have_lo && lo < end) { e.cg_lo = k; e.cg_lo_len = (end < hi ? end : hi) - lo; have_lo = true; }
        if (have_lo && hi <= end) { e.cg_hi = k; e.cg_hi_len = hi - (pos > lo ? pos : lo); break; }
        pos = end;
    }
}

// Everything after the load filters for one read. `idx[0..n)` = rows of the read's hits that passed F1-F4,
// in PAF order (it is permuted in place); `hit` = scratch for n K1Hit; `dp`/`prevc`/`cand` = scratch for n words
// each; `take` = scratch for n bytes. Elements are written to `out` (capacity n). Returns their count.
// `cg_total` (may be null): expanded CIGAR length of every row of `idx`, indexed by row, precomputed by the caller (the kernel
// sums the runs lane-parallel); without it the runs are summed here.
K1_HD uint32_t k1_process_read(const HitCols& h, const double* mean_kmer, const K1Params& p,
                               uint32_t* idx, uint32_t n, K1Hit* hit, uint32_t* dp, int32_t* prevc, uint32_t* cand,
                               uint8_t* take, ClElem* out, const uint32_t* cg_total = nullptr) {
    const double uf = p.uniq_freq, dev = p.max_uniq_dev;
    // per-read sort by (q_end, q_start), Longread.cpp:256
    KeyLess lt{h.q_end, h.q_start};
    libstdcxx_sort(idx, (int)n, lt);
    if (n <= 1) return 0;                                                                       // :184
    // palindrome rule :187-202 — cut the group at the second occurrence of a unique contig
    {
        uint32_t cut = n;
        for (uint32_t i = 0; i < n && cut == n; ++i) {
            const uint32_t tid = h.t_id[idx[i]];
            if (mean_kmer[tid] < uf * (1 + dev)) {
                for (uint32_t j = 0; j < i; ++j)
                    if (h.t_id[idx[j]] == tid && mean_kmer[tid] < uf * (1 + dev)) { cut = i; break; }
            }
        }
        n = cut;
    }
    // F5 :207 — interior hits must cover 80% of their contig
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t r = idx[i];
        if (i > 0 && i < n - 1 && (h.t_end[r] - h.t_start[r]) / (double)h.t_len[r] < 0.8) continue;
        K1Hit& x = hit[m++];
        x.src = r; x.q_start = h.q_start[r]; x.q_end = h.q_end[r]; x.t_start = h.t_start[r]; x.t_end = h.t_end[r];
        x.n_match = h.n_match[r]; x.n_block = h.n_block[r]; x.is_rev = h.is_rev[r];
        uint32_t tot = 0;
        if (cg_total) tot = cg_total[r];
        else for (uint32_t k = h.cg_off[r]; k < h.cg_off[r + 1]; ++k) tot += h.cg_ops[k] >> 2;
        x.lo = 0; x.hi = tot;
    }
    // overlap fix :430-512 on adjacent pairs, in place
    for (uint32_t i = 0; i + 1 < m; ++i) {
        K1Hit& a = hit[i];
        K1Hit& b = hit[i + 1];
        if (!(a.q_end > b.q_start)) continue;
        const long long ov = (long long)a.q_end - (long long)b.q_start;
        {
            const RleCigar cg{h.cg_ops + h.cg_off[a.src], h.cg_off[a.src + 1] - h.cg_off[a.src]};
            const uint32_t target = (uint32_t)((long long)a.q_end - ov / 2 - 1);
            uint32_t rq, rt, mk;
            if (a.is_rev == 0) {
                rq = a.q_start; rt = a.t_start;
                const uint32_t kept = rle_find_contig_pos(cg, a.lo, a.hi, true, rq, rt, +1, +1, target, &mk);
                a.q_end = rq + 1; a.t_end = rt + 1;
                a.hi = a.lo + kept; a.n_block = kept; a.n_match = mk;
            } else {
                rq = a.q_start; rt = a.t_end - 1;
                const uint32_t kept = rle_find_contig_pos(cg, a.lo, a.hi, false, rq, rt, +1, -1, target, &mk);
                a.q_end = rq + 1; a.t_start = rt;
                a.lo = a.hi - kept; a.n_block = kept; a.n_match = mk;
            }
        }
        {
            const RleCigar cg{h.cg_ops + h.cg_off[b.src], h.cg_off[b.src + 1] - h.cg_off[b.src]};
            const uint32_t target = (uint32_t)((long long)b.q_start + (ov - ov / 2));
            uint32_t rq, rt, mk;
            if (b.is_rev == 0) {
                rq = b.q_end - 1; rt = b.t_end - 1;
                const uint32_t kept = rle_find_contig_pos(cg, b.lo, b.hi, false, rq, rt, -1, -1, target, &mk);
                b.q_start = rq; b.t_start = rt;
                b.lo = b.hi - kept; b.n_block = kept; b.n_match = mk;
            } else {
                rq = b.q_end - 1; rt = b.t_start;
                const uint32_t kept = rle_find_contig_pos(cg, b.lo, b.hi, true, rq, rt, -1, +1, target, &mk);
                b.q_start = rq; b.t_end = rt + 1;
                b.hi = b.lo + kept; b.n_block = kept; b.n_match = mk;
            }
        }
    }
    // weighted interval scheduling :524-610 over unique, long-enough hits (weights = post-fix n_match)
    uint32_t nc = 0;
    for (uint32_t i = 0; i < m; ++i) {
        if (hit[i].n_block < p.min_aln_block) continue;
        if (mean_kmer[h.t_id[hit[i].src]] > uf * (1 + dev)) continue;
        cand[nc++] = i;
    }
    if (nc == 0) return 0;
    dp[0] = hit[cand[0]].n_match; take[0] = 1; prevc[0] = -1;
    for (uint32_t i = 1; i < nc; ++i) {
        int j;
        for (j = (int)i - 1; j >= 0; --j) if (hit[cand[j]].q_end <= hit[cand[i]].q_start) break;
        prevc[i] = j;
        const uint32_t v = hit[cand[i]].n_match + (j >= 0 ? dp[j] : 0);
        if (v > dp[i - 1]) { dp[i] = v; take[i] = 1; } else { dp[i] = dp[i - 1]; take[i] = 0; }
    }
    // backtrack (selected candidates come out last-first; reverse while writing)
    uint32_t ns = 0;
    for (int i = (int)nc - 1; i >= 0;) { if (take[i]) { ++ns; i = prevc[i]; } else { --i; } }
    uint32_t w = ns;
    for (int i = (int)nc - 1; i >= 0;) {
        if (take[i]) {
            const K1Hit& x = hit[cand[i]];
            ClElem e;
            e.hit = x.src; e.q_start = x.q_start; e.q_end = x.q_end; e.t_start = x.t_start; e.t_end = x.t_end;
            e.n_match = x.n_match; e.n_block = x.n_block;
            const RleCigar cg{h.cg_ops + h.cg_off[x.src], h.cg_off[x.src + 1] - h.cg_off[x.src]};
            k1_cigar_window(cg, x.lo, x.hi, e);
            out[--w] = e;
            i = prevc[i];
        } else {
            --i;
        }
    }
    return ns;
}

}  // namespace hgpu
