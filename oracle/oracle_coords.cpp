// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h). Edge coordinates, restated from the reference's
// src/haslr_assemble/src/Assemble.cpp the way the reference computes them: sorted begin/end lists swept with a
// std::set of supporting reads (asm_best_supported_interval_contig1/2, Assemble.cpp:24-126), CIGARs expanded to one
// character per operation and walked character by character (asm_find_lr_pos, Assemble.cpp:129-155), the eight
// strand cases of asm_calc_single_edge_coordinates (Assemble.cpp:157-363). Pinned against log_coordinate.txt of the
// reference binary (tests/golden/syn200k_coords.txt, tests/test_oracle_golden.py).
#include <algorithm>
#include <iterator>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {

typedef std::pair<uint32_t, uint32_t> P;

// Assemble.cpp:24-74 (ge = true, "curr_supp >= best_supp") and :76-126 (ge = false, ">")
void best_interval(std::vector<P>& beg_list, std::vector<P>& end_list, bool ge, P& best_int, std::set<uint32_t>& best_lrs) {
    std::sort(beg_list.begin(), beg_list.end());
    std::sort(end_list.begin(), end_list.end());
    int curr_supp = 0, best_supp = 0, i = 0, j = 0;
    const int len = (int)beg_list.size();
    uint32_t beg_best = 0, end_best = 0;          // uninitialised in the reference; every edge has >= 1 support
    bool interval_started = false;
    std::set<uint32_t> curr_lrs;
    while (i < len && j < len) {
        if (beg_list[i].first < end_list[j].first) {
            curr_supp++;
            curr_lrs.insert(beg_list[i].second);
            if (ge ? curr_supp >= best_supp : curr_supp > best_supp) {
                best_supp = curr_supp;
                beg_best = beg_list[i].first;
                best_lrs = curr_lrs;
                interval_started = true;
            }
            i++;
        } else {
            if (interval_started) { end_best = end_list[j].first; interval_started = false; }
            curr_supp--;
            curr_lrs.erase(end_list[j].second);
            j++;
        }
    }
    if (interval_started) end_best = end_list[j].first;
    best_int = P(beg_best, end_best);
}

// Assemble.cpp:129-155, on the expanded CIGAR
long long find_lr_pos(const std::string& cigar_str, uint32_t lr_curr, uint32_t c_curr, int lr_step, int c_step, uint32_t contig_pos) {
    if ((c_step > 0 && c_curr > contig_pos) || (c_step < 0 && c_curr < contig_pos)) return -1;
    for (size_t i = 0; i < cigar_str.size(); i++) {
        if (c_curr == contig_pos) break;
        if (cigar_str[i] == 'M') { c_curr += c_step; lr_curr += lr_step; }
        else if (cigar_str[i] == 'I') { lr_curr += lr_step; }
        else { c_curr += c_step; }
    }
    return lr_curr;
}

// the CIGAR an element carries after the overlap fix, one character per operation (expand_cigar, Common.cpp)
std::string expand(const oracle_cl_elem& e, const uint32_t* cg_off, const uint32_t* cg_ops) {
    std::string s;
    if (cg_off[e.hit + 1] == cg_off[e.hit]) return s;        // PAF row without a cg:Z: tag
    const uint32_t* ops = cg_ops + cg_off[e.hit];
    for (uint32_t r = e.cg_lo; r <= e.cg_hi; ++r) {
        const uint32_t len = r == e.cg_lo ? e.cg_lo_len : (r == e.cg_hi ? e.cg_hi_len : ops[r] >> 2);
        s.append(len, "MID"[std::min<uint32_t>(ops[r] & 3u, 2u)]);
    }
    return s;
}

}  // namespace

extern "C" int oracle_edge_coords(uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const oracle_edge_supp* supp,
                                  const oracle_cl_elem* elems, const uint32_t* cl_read_off, const uint32_t* read_len,
                                  const uint8_t* hit_is_rev, const uint32_t* cg_off, const uint32_t* cg_ops,
                                  oracle_edge_coord* out_edge, oracle_supp_coord* out_supp) {
    for (uint32_t e = 0; e < n_edges; ++e) {
        const uint32_t rev1 = edge_rev[e] & 1u, rev2 = (edge_rev[e] >> 1) & 1u;
        const oracle_edge_supp* es = supp + supp_off[e];
        const uint32_t n = supp_off[e + 1] - supp_off[e];
        oracle_supp_coord* os = out_supp + supp_off[e];
        auto elem = [&](uint32_t i, bool head) -> const oracle_cl_elem& {
            const uint32_t rid = es[i].lr_id_strand & 0x7FFFFFFFu;
            return elems[cl_read_off[rid] + (head ? es[i].cmp_head : es[i].cmp_tail)];
        };
        std::vector<P> beg1, end1, beg2, end2;                 // Assemble.cpp:196-226
        for (uint32_t i = 0; i < n; ++i) {
            beg1.push_back(P(elem(i, true).t_start, i)); end1.push_back(P(elem(i, true).t_end, i));
            beg2.push_back(P(elem(i, false).t_start, i)); end2.push_back(P(elem(i, false).t_end, i));
        }
        P int1, int2;
        std::set<uint32_t> lrs1, lrs2;
        best_interval(beg1, end1, true, int1, lrs1);
        best_interval(beg2, end2, false, int2, lrs2);
        const uint32_t c1 = rev1 == 0 ? int1.second - 1 : int1.first;      // Assemble.cpp:228-238
        const uint32_t c2 = rev2 == 0 ? int2.first : int2.second - 1;
        std::vector<uint32_t> best;
        std::set_intersection(lrs1.begin(), lrs1.end(), lrs2.begin(), lrs2.end(), std::inserter(best, best.begin()));
        for (uint32_t i = 0; i < n; ++i) { os[i].lr_start = -1; os[i].lr_end = -1; os[i].lr_strand = 0; os[i].in_best = 0; }
        uint32_t n_cns = 0;
        for (uint32_t bi : best) {                               // Assemble.cpp:255-338
            const uint32_t rid = es[bi].lr_id_strand & 0x7FFFFFFFu, rlen = read_len[rid];
            const oracle_cl_elem& a1 = elem(bi, true);
            const oracle_cl_elem& a2 = elem(bi, false);
            const uint32_t rstrand = (rev1 == hit_is_rev[a1.hit]) ? 0 : 1;
            long long lr_start = -1, lr_end = -1;
            std::string cg = expand(a1, cg_off, cg_ops), cg_rev(cg.rbegin(), cg.rend());
            const uint32_t q0h = rstrand == 0 ? a1.q_start : rlen - a1.q_end;
            if (rev1 == 0) lr_start = find_lr_pos(cg, q0h, a1.t_start, +1, +1, c1);            // cases 1 / 5
            else lr_start = find_lr_pos(cg_rev, q0h, a1.t_end - 1, +1, -1, c1);               // cases 2 / 6
            cg = expand(a2, cg_off, cg_ops); cg_rev.assign(cg.rbegin(), cg.rend());
            const uint32_t q0t = rstrand == 0 ? a2.q_end - 1 : rlen - a2.q_start - 1;
            if (rev2 == 0) lr_end = find_lr_pos(cg_rev, q0t, a2.t_end - 1, -1, -1, c2);        // cases 3 / 7
            else lr_end = find_lr_pos(cg, q0t, a2.t_start, -1, +1, c2);                        // cases 4 / 8
            os[bi].lr_start = lr_start; os[bi].lr_end = lr_end; os[bi].lr_strand = rstrand; os[bi].in_best = 1;
            if (lr_start != -1 && lr_end != -1) ++n_cns;
        }
        oracle_edge_coord& oe = out_edge[e];
        oe.int1_lo = int1.first; oe.int1_hi = int1.second; oe.int2_lo = int2.first; oe.int2_hi = int2.second;
        oe.c1 = c1; oe.c2 = c2; oe.n_best = (uint32_t)best.size(); oe.n_cns = n_cns;
    }
    return 0;
}
