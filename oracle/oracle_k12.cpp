// ORACLE — TEST INFRASTRUCTURE ONLY. See oracle.h.
//
// CPU restatement of the reference's per-read hit pipeline and backbone edge table:
//   load filters F1-F4 + per-read sort        Longread.cpp:234-302 (load_alignment)
//   group rule, palindrome truncation, F5     Longread.cpp:182-232 (process_lr_alignment_group)
//   overlap fix by CIGAR walk                 Longread.cpp:375-512 (find_contig_pos, fix_overlapping_alignments)
//   weighted interval scheduling              Longread.cpp:514-610 (find_best_scheduling)
//   edge table + weak-edge rule               Backbone_graph.cpp:10-25,148-171,348-375
// Pinned against outputs of the reference binary (oracle/_ref/haslr_assemble_ref): tests/golden/.
//
// Input precondition (the reference's silent assumption): PAF hits are grouped by read and
// reads appear in ascending id, so a "group" is exactly one read's hits.
#include "oracle.h"

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Hit {
    uint32_t src;  // index into the PAF hit arrays
    uint32_t q_start, q_end, t_id, t_len, t_start, t_end, n_match, n_block;
    uint8_t is_rev;
    std::string exp;  // current expanded CIGAR, one char per base-op
    uint32_t cut_left = 0, cut_right = 0;  // expanded ops removed so far at either end
};

std::string expand(const oracle_hits_t* h, uint32_t i) {
    std::string s;
    for (uint32_t k = h->cg_off[i]; k < h->cg_off[i + 1]; ++k) {
        uint32_t op = h->cg_ops[k] & 3u, len = h->cg_ops[k] >> 2;
        s.append(len, op == 0 ? 'M' : op == 1 ? 'I' : 'D');
    }
    return s;
}

// Longread.cpp:375-420. Walks `c` until lr_curr == lr_pos, backs up to the nearest 'M', truncates after it.
void find_contig_pos(std::string& c, uint32_t& lr_curr, uint32_t& c_curr, int lr_step, int c_step, uint32_t lr_pos) {
    uint32_t i;
    for (i = 0; i < c.size(); i++) {
        if (lr_curr == lr_pos) break;
        if (c[i] == 'M') { c_curr += c_step; lr_curr += lr_step; }
        else if (c[i] == 'I') { lr_curr += lr_step; }
        else { c_curr += c_step; }
    }
    // c[c.size()] reads the terminating '\0' in the reference (Q3); model it as "not M"
    while (i > 0 && (i >= c.size() || c[i] != 'M')) {
        if (c[i - 1] == 'M') { c_curr -= c_step; lr_curr -= lr_step; }
        else if (c[i - 1] == 'I') { lr_curr -= lr_step; }
        else if (c[i - 1] == 'D') { c_curr -= c_step; }
        i--;
    }
    if (i + 1 < c.size()) c.erase(i + 1);
}

uint32_t count_m(const std::string& s) { return (uint32_t)std::count(s.begin(), s.end(), 'M'); }

// Longread.cpp:430-512
void fix_overlaps(std::vector<Hit>& a) {
    for (int i = 0; i + 1 < (int)a.size(); i++) {
        if (!(a[i].q_end > a[i + 1].q_start)) continue;
        long long ov = (long long)a[i].q_end - (long long)a[i + 1].q_start;
        uint32_t rq, rt;
        {
            Hit& h = a[i];
            uint32_t target = (uint32_t)((long long)h.q_end - ov / 2 - 1);
            size_t before = h.exp.size();
            if (h.is_rev == 0) {
                rq = h.q_start; rt = h.t_start;
                find_contig_pos(h.exp, rq, rt, +1, +1, target);
                h.q_end = rq + 1; h.t_end = rt + 1;
                h.cut_right += (uint32_t)(before - h.exp.size());
            } else {
                std::reverse(h.exp.begin(), h.exp.end());
                rq = h.q_start; rt = h.t_end - 1;
                find_contig_pos(h.exp, rq, rt, +1, -1, target);
                h.q_end = rq + 1; h.t_start = rt;
                std::reverse(h.exp.begin(), h.exp.end());
                h.cut_left += (uint32_t)(before - h.exp.size());
            }
            h.n_block = (uint32_t)h.exp.size();
            h.n_match = count_m(h.exp);
        }
        {
            Hit& h = a[i + 1];
            uint32_t target = (uint32_t)((long long)h.q_start + (ov - ov / 2));
            size_t before = h.exp.size();
            if (h.is_rev == 0) {
                std::reverse(h.exp.begin(), h.exp.end());
                rq = h.q_end - 1; rt = h.t_end - 1;
                find_contig_pos(h.exp, rq, rt, -1, -1, target);
                h.q_start = rq; h.t_start = rt;
                std::reverse(h.exp.begin(), h.exp.end());
                h.cut_left += (uint32_t)(before - h.exp.size());
            } else {
                rq = h.q_end - 1; rt = h.t_start;
                find_contig_pos(h.exp, rq, rt, -1, +1, target);
                h.q_start = rq; h.t_end = rt + 1;
                h.cut_right += (uint32_t)(before - h.exp.size());
            }
            h.n_block = (uint32_t)h.exp.size();
            h.n_match = count_m(h.exp);
        }
    }
}

// kept expanded range [cut_left, total - cut_right) -> window over the hit's run-length ops
void cigar_window(const oracle_hits_t* h, const Hit& hit, oracle_cl_elem& e) {
    uint32_t b = h->cg_off[hit.src], n = h->cg_off[hit.src + 1] - b;
    uint64_t lo = hit.cut_left, hi = 0;
    for (uint32_t k = 0; k < n; ++k) hi += h->cg_ops[b + k] >> 2;
    hi -= hit.cut_right;  // exclusive
    e.cg_lo = e.cg_hi = 0; e.cg_lo_len = e.cg_hi_len = 0;
    uint64_t pos = 0;
    bool have_lo = false;
    for (uint32_t k = 0; k < n; ++k) {
        uint64_t len = h->cg_ops[b + k] >> 2, end = pos + len;
        if (!have_lo && lo < end) { e.cg_lo = k; e.cg_lo_len = (uint32_t)(std::min(end, hi) - lo); have_lo = true; }
        if (have_lo && hi <= end) { e.cg_hi = k; e.cg_hi_len = (uint32_t)(hi - std::max(pos, lo)); break; }
        pos = end;
    }
}

}  // namespace

extern "C" int64_t oracle_compact_lr(const oracle_hits_t* hits, const uint32_t* read_off, uint32_t n_reads,
                                     const double* mean_kmer, const oracle_k1_params* prm,
                                     oracle_cl_elem* out_elems, uint32_t* out_read_off) {
    int64_t n_out = 0;
    const double uf = prm->uniq_freq, dev = prm->max_uniq_dev;
    for (uint32_t r = 0; r < n_reads; ++r) {
        out_read_off[r] = (uint32_t)n_out;
        std::vector<Hit> grp;
        for (uint32_t i = read_off[r]; i < read_off[r + 1]; ++i) {
            if (hits->n_block[i] < prm->min_aln_block) continue;                                       // F1 :262
            if ((double)hits->n_match[i] / (double)hits->n_block[i] < prm->min_aln_sim) continue;       // F2 :265
            if ((uint32_t)hits->mapq[i] < prm->min_aln_mapq) continue;                                   // F3 :268
            if (mean_kmer[hits->t_id[i]] > uf * (3 + dev)) continue;                                     // F4 :272
            Hit h;
            h.src = i; h.q_start = hits->q_start[i]; h.q_end = hits->q_end[i]; h.t_id = hits->t_id[i]; h.t_len = hits->t_len[i];
            h.t_start = hits->t_start[i]; h.t_end = hits->t_end[i]; h.n_match = hits->n_match[i]; h.n_block = hits->n_block[i];
            h.is_rev = hits->is_rev[i];
            grp.push_back(h);
        }
        // :256 — the comparator looks at (q_end, q_start) only; ties resolve however libstdc++'s introsort leaves them (Q2)
        std::sort(grp.begin(), grp.end(), [](const Hit& a, const Hit& b) {
            return (a.q_end < b.q_end) || (a.q_end == b.q_end && a.q_start < b.q_start);
        });
        if (grp.size() <= 1) continue;                                                                   // :184
        {   // palindrome rule :187-202
            std::unordered_map<uint32_t, uint32_t> seen;
            for (uint32_t i = 0; i < grp.size(); i++) {
                uint32_t tid = grp[i].t_id;
                if (mean_kmer[tid] < uf * (1 + dev)) {
                    if (seen.count(tid) > 0) grp.resize(i); else seen[tid] = i;
                }
            }
        }
        std::vector<Hit> kept;
        for (uint32_t i = 0; i < grp.size(); i++) {                                                      // F5 :207
            if (i > 0 && i < grp.size() - 1 && (grp[i].t_end - grp[i].t_start) / (double)grp[i].t_len < 0.8) continue;
            kept.push_back(grp[i]);
        }
        for (auto& h : kept) h.exp = expand(hits, h.src);
        fix_overlaps(kept);
        // chaining :524-610
        std::vector<uint32_t> cand;
        for (uint32_t i = 0; i < kept.size(); i++) {
            if (kept[i].n_block < prm->min_aln_block) continue;
            if (mean_kmer[kept[i].t_id] > uf * (1 + dev)) continue;
            cand.push_back(i);
        }
        if (cand.empty()) continue;
        const int n = (int)cand.size();
        std::vector<uint32_t> dp(n);
        std::vector<int> prevc(n, -1);
        std::vector<uint8_t> take(n, 0);
        dp[0] = kept[cand[0]].n_match; take[0] = 1;
        for (int i = 1; i < n; i++) {
            int j;
            for (j = i - 1; j >= 0; j--) if (kept[cand[j]].q_end <= kept[cand[i]].q_start) break;
            prevc[i] = j;
            uint32_t v = kept[cand[i]].n_match + (j >= 0 ? dp[j] : 0);
            if (v > dp[i - 1]) { dp[i] = v; take[i] = 1; } else { dp[i] = dp[i - 1]; }
        }
        std::vector<int> sel;
        for (int i = n - 1; i >= 0;) { if (take[i]) { sel.push_back(i); i = prevc[i]; } else { i--; } }
        std::reverse(sel.begin(), sel.end());
        for (int s : sel) {
            const Hit& h = kept[cand[s]];
            oracle_cl_elem e;
            e.hit = h.src; e.q_start = h.q_start; e.q_end = h.q_end; e.t_start = h.t_start; e.t_end = h.t_end;
            e.n_match = h.n_match; e.n_block = h.n_block;
            cigar_window(hits, h, e);
            out_elems[n_out++] = e;
        }
    }
    out_read_off[n_reads] = (uint32_t)n_out;
    return n_out;
}

extern "C" int64_t oracle_backbone_edges(const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off, uint32_t n_reads,
                                         uint32_t min_edge_sup,
                                         uint64_t* out_key, uint32_t* out_supp_off, oracle_edge_supp* out_supp, uint8_t* out_keep) {
    // graph[node].edges[rev][to] as one ordered map keyed like the reference's nested iteration order
    std::map<uint64_t, std::vector<oracle_edge_supp>> tab;
    for (uint32_t r = 0; r < n_reads; ++r) {
        uint32_t b = cl_read_off[r], e = cl_read_off[r + 1];
        if (e - b <= 1) continue;
        for (uint32_t j = b; j + 1 < e; ++j) {  // Backbone_graph.cpp:166-167 (every element is unique by construction)
            uint32_t node1 = cl_tid[j], rev1 = cl_rev[j], node2 = cl_tid[j + 1], rev2 = cl_rev[j + 1];
            uint32_t to1 = (node2 << 1) | rev2, to2 = (node1 << 1) | (1 - rev1);
            uint64_t k1 = ((uint64_t)((node1 << 1) | rev1) << 32) | to1;
            uint64_t k2 = ((uint64_t)((node2 << 1) | (1 - rev2)) << 32) | to2;
            tab[k1].push_back({r, j - b, j + 1 - b});                       // :23
            tab[k2].push_back({r | 0x80000000u, j + 1 - b, j - b});          // :24
        }
    }
    int64_t n = 0;
    uint32_t so = 0;
    for (auto& kv : tab) {
        out_key[n] = kv.first;
        out_supp_off[n] = so;
        for (auto& s : kv.second) out_supp[so++] = s;
        out_keep[n] = kv.second.size() >= min_edge_sup ? 1 : 0;             // :358
        ++n;
    }
    out_supp_off[n] = so;
    return n;
}
