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

// Assemble.cpp:24-74 (later_wins = true: a later interval of EQUAL depth replaces the optimum, the `>=` of contig1) and
// :76-126 (later_wins = false, the `>` of contig2). Two sorted event lists are merged; the set of reads currently
// covering the sweep position is copied whenever a new optimum is reached; the optimum's interval closes at the next
// end event.
void best_interval(std::vector<P>& starts, std::vector<P>& stops, bool later_wins, P& interval, std::set<uint32_t>& members) {
    std::sort(starts.begin(), starts.end());
    std::sort(stops.begin(), stops.end());
    const int n = (int)starts.size();
    int depth = 0, top = 0, a = 0, z = 0;
    uint32_t lo = 0, hi = 0;                      // uninitialised in the reference; every edge has >= 1 support
    bool open = false;
    std::set<uint32_t> covering;
    while (a < n && z < n) {
        const bool start_event = starts[a].first < stops[z].first;      // on a tie the stop is taken first
        if (start_event) {
            ++depth;
            covering.insert(starts[a].second);
            const bool better = later_wins ? depth >= top : depth > top;
            if (better) { top = depth; lo = starts[a].first; members = covering; open = true; }
            ++a;
        } else {
            if (open) { hi = stops[z].first; open = false; }
            --depth;
            covering.erase(stops[z].second);
            ++z;
        }
    }
    if (open) hi = stops[z].first;
    interval = P(lo, hi);
}

// Assemble.cpp:129-155, on the expanded CIGAR: step through the operations until the contig coordinate arrives at
// `target`; M moves both coordinates, I the read only, anything else the contig only. -1 if the start is already past it.
long long find_lr_pos(const std::string& ops, uint32_t read_pos, uint32_t contig_at, int read_dir, int contig_dir, uint32_t target) {
    const bool past = contig_dir > 0 ? contig_at > target : contig_at < target;
    if (past) return -1;
    for (size_t k = 0; k < ops.size() && contig_at != target; ++k) {
        const char op = ops[k];
        if (op != 'I') contig_at += contig_dir;
        if (op == 'M' || op == 'I') read_pos += read_dir;
    }
    return read_pos;
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
