// Edge coordinates, consensus calling (GPU) and stitching.
//
// Restates the behaviour of the reference's src/haslr_assemble/src/Assemble.cpp: the canonical edge enumeration
// (:365-434), the best-supported-interval sweeps (:24-126), the eight CIGAR-walk cases that map a contig position to
// a read position (:129-155,253-338), the per-edge consensus call (:479-560 — here ONE batched hgpu_poa_batch call
// instead of a SPOA engine per edge per thread) and the simple-path stitching (:607-810,1045-1077).
#include <algorithm>
#include <cstring>
#include <deque>
#include <iterator>
#include <set>
#include <thread>
#include <ctime>

#include "haslr.hpp"

namespace haslr {

static inline char sgn(uint32_t s) { return s ? '-' : '+'; }

// every undirected edge once, in (node*2+strand, map key) order; edge and twin get `flag` (Assemble.cpp:365-434)
void enumerate_edges(Graph& g, uint32_t flag, std::vector<EdgeRef>& out) {
    out.clear();
    for (uint32_t v = 0; v < 2 * g.size(); ++v) {
        const uint32_t node1 = v / 2, rev1 = v % 2;
        for (auto& kv : g[node1].edges[rev1]) {
            if (kv.second.flag == flag) continue;
            const uint32_t node2 = kv.first >> 1, rev2 = kv.first & 1;
            kv.second.flag = flag;
            auto tw = g[node2].edges[1 - rev2].find((node1 << 1) | (1 - rev1));
            if (tw != g[node2].edges[1 - rev2].end()) tw->second.flag = flag;
            out.push_back({node1, rev1, node2, rev2});
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// best supported interval on one anchor contig: sweep over sorted begin/end lists (Assemble.cpp:24-126).
// ge = true for the head contig (a later interval of equal depth wins), false for the tail contig.
// ---------------------------------------------------------------------------------------------------------
static void best_interval(std::vector<std::pair<uint32_t, uint32_t>>& beg, std::vector<std::pair<uint32_t, uint32_t>>& end, bool ge,
                          std::pair<uint32_t, uint32_t>& best, std::set<uint32_t>& best_lrs) {
    std::sort(beg.begin(), beg.end());
    std::sort(end.begin(), end.end());
    int cur = 0, top = 0, i = 0, j = 0;
    const int len = (int)beg.size();
    uint32_t lo = 0, hi = 0;
    bool open = false;
    std::set<uint32_t> cur_lrs;
    while (i < len && j < len) {
        if (beg[i].first < end[j].first) {
            ++cur;
            cur_lrs.insert(beg[i].second);
            if (ge ? cur >= top : cur > top) { top = cur; lo = beg[i].first; best_lrs = cur_lrs; open = true; }
            ++i;
        } else {
            if (open) { hi = end[j].first; open = false; }
            --cur;
            cur_lrs.erase(end[j].second);
            ++j;
        }
    }
    if (open) hi = end[j].first;
    best = {lo, hi};
}

// The kept part of a compact element's CIGAR as a run list: runs cg_lo..cg_hi of the hit, first/last clipped.
struct RunView {
    const uint32_t* ops; uint32_t lo, hi, lo_len, hi_len;
    uint32_t n() const { return hi - lo + 1; }
    // k-th run in forward order
    void get(uint32_t k, uint32_t& op, uint32_t& len) const {
        const uint32_t r = lo + k;
        op = ops[r] & 3u;
        len = (r == lo) ? lo_len : (r == hi ? hi_len : ops[r] >> 2);
    }
};
static RunView run_view(const PafTable& paf, const hgpu_cl_elem& e) {
    if (paf.cg_off[e.hit + 1] == paf.cg_off[e.hit]) return RunView{paf.cg_ops.data(), 1, 0, 0, 0};   // row without cg:Z: — no ops to walk
    return RunView{paf.cg_ops.data() + paf.cg_off[e.hit], e.cg_lo, e.cg_hi, e.cg_lo_len, e.cg_hi_len};
}

// asm_find_lr_pos (Assemble.cpp:129-155) on runs: walk until the contig coordinate reaches `contig_pos`
static long long find_lr_pos(const RunView& cg, bool reversed, uint32_t lr_curr, uint32_t c_curr, int lr_step, int c_step, uint32_t contig_pos) {
    if ((c_step > 0 && c_curr > contig_pos) || (c_step < 0 && c_curr < contig_pos)) return -1;
    uint32_t dist = c_step > 0 ? contig_pos - c_curr : c_curr - contig_pos;     // contig steps still to go
    const uint32_t n = cg.n();
    for (uint32_t k = 0; k < n && dist > 0; ++k) {
        uint32_t op, len;
        cg.get(reversed ? n - 1 - k : k, op, len);
        if (op == 0) {                     // M: both move
            const uint32_t take = std::min(len, dist);
            lr_curr += (uint32_t)lr_step * take; dist -= take;
        } else if (op == 1) {              // I: only the read moves
            lr_curr += (uint32_t)lr_step * len;
        } else {                           // D / anything else: only the contig moves
            dist -= std::min(len, dist);
        }
    }
    return lr_curr;
}

void calc_edge_coordinates(Graph& g, const std::vector<EdgeRef>& edges, const ContigStore& contigs, const SeqStore& reads,
                           const CompactReads& cl, const PafTable& paf, const std::string& logpath) {
    if (edges.empty()) return;             // the reference returns before opening the log (quirk Q12)
    FILE* fp = logpath.empty() ? nullptr : open_write(logpath);
#define LOG(...) do { if (fp) fprintf(fp, __VA_ARGS__); } while (0)
    auto elem = [&](uint32_t rid, uint32_t cmp) -> const hgpu_cl_elem& { return cl.elems[cl.off[rid] + cmp]; };
    for (const EdgeRef& er : edges) {
        const uint32_t node1 = er.node1, rev1 = er.rev1, node2 = er.node2, rev2 = er.rev2;
        LOG("calc_coords th_id:%d %u:%c -> %u:%c\n", 0, node1, sgn(rev1), node2, sgn(rev2));
        Edge& edge1 = g[node1].edges[rev1][(node2 << 1) | rev2];
        Edge& edge2 = g[node2].edges[1 - rev2][(node1 << 1) | (1 - rev1)];
        LOG("edge      %u:%c -> %u:%c\n", node1, sgn(rev1), node2, sgn(rev2));
        LOG("edge_twin %u:%c -> %u:%c\n", node2, sgn(1 - rev2), node1, sgn(1 - rev1));
        const std::vector<EdgeSupp>& es = edge1.edge_supp;
        LOG("\tedge_supp size:%zu\n", es.size());
        std::vector<std::pair<uint32_t, uint32_t>> beg1, end1, beg2, end2;
        for (uint32_t i = 0; i < es.size(); ++i) {
            const hgpu_cl_elem& h = elem(es[i].lr_id, es[i].cmp_head_id);
            const hgpu_cl_elem& t = elem(es[i].lr_id, es[i].cmp_tail_id);
            LOG("\tsupp_detail head\t%u\t%u\t%c\ttail\t%u\t%u\t%c\n", h.t_start, h.t_end, sgn(paf.is_rev[h.hit]), t.t_start, t.t_end, sgn(paf.is_rev[t.hit]));
            beg1.push_back({h.t_start, i}); end1.push_back({h.t_end, i});
            beg2.push_back({t.t_start, i}); end2.push_back({t.t_end, i});
        }
        std::pair<uint32_t, uint32_t> int1, int2;
        std::set<uint32_t> lrs1, lrs2;
        best_interval(beg1, end1, true, int1, lrs1);
        LOG("    @@@ best interval contig1 %u %u\n", int1.first, int1.second);
        best_interval(beg2, end2, false, int2, lrs2);
        LOG("    @@@ best_interval contig2 %u %u\n", int2.first, int2.second);
        const uint32_t c1 = rev1 == 0 ? int1.second - 1 : int1.first;      // last shared base on the head contig
        const uint32_t c2 = rev2 == 0 ? int2.first : int2.second - 1;      // first shared base on the tail contig
        std::vector<uint32_t> best;
        std::set_intersection(lrs1.begin(), lrs1.end(), lrs2.begin(), lrs2.end(), std::back_inserter(best));
        LOG("coordinates contig1_pos: %u\tcontig2_pos: %u\n", c1, c2);
        LOG("supproting_lr: %lu\n", (unsigned long)best.size());
        auto no_support = [&]() {
            edge1.cns_supp.clear(); edge2.cns_supp.clear();
            edge1.head_end = edge2.tail_beg = (rev1 == 0 ? contigs.len(node1) - 1 : 0);
            edge1.tail_beg = edge2.head_end = (rev2 == 0 ? 0 : contigs.len(node2) - 1);
        };
        if (best.empty()) { no_support(); continue; }
        for (uint32_t bi : best) {
            const uint32_t rid = es[bi].lr_id, rlen = reads.len(rid);
            const hgpu_cl_elem& a1 = elem(rid, es[bi].cmp_head_id);
            const hgpu_cl_elem& a2 = elem(rid, es[bi].cmp_tail_id);
            const uint32_t rstrand = (rev1 == paf.is_rev[a1.hit]) ? 0 : 1;
            LOG("    +++ lr:%u len:%u strand:%c\n", rid, rlen, sgn(rstrand));
            const RunView cg1 = run_view(paf, a1), cg2 = run_view(paf, a2);
            long long lr_start, lr_end;
            // head anchor: where on the (oriented) read does contig position c1 fall; cases 1/2 and 5/6 differ only in q0
            const uint32_t q0h = rstrand == 0 ? a1.q_start : rlen - a1.q_end;
            const uint32_t q0t = rstrand == 0 ? a2.q_end - 1 : rlen - a2.q_start - 1;
            if (rev1 == 0) { LOG("        case %d\n", rstrand ? 5 : 1); lr_start = find_lr_pos(cg1, false, q0h, a1.t_start, +1, +1, c1); }
            else           { LOG("        case %d\n", rstrand ? 6 : 2); lr_start = find_lr_pos(cg1, true, q0h, a1.t_end - 1, +1, -1, c1); }
            if (rev2 == 0) { LOG("        case %d\n", rstrand ? 7 : 3); lr_end = find_lr_pos(cg2, true, q0t, a2.t_end - 1, -1, -1, c2); }
            else           { LOG("        case %d\n", rstrand ? 8 : 4); lr_end = find_lr_pos(cg2, false, q0t, a2.t_start, -1, +1, c2); }
            if (lr_start != -1 && lr_end != -1) {
                LOG("        [coordinate] subseq_len:%lld lr_start:%lld lr_end:%lld\n", lr_end - lr_start - 1, lr_start + 1, lr_end - 1);
                edge1.cns_supp.push_back({rid, rstrand, uint32_t(lr_start + 1), uint32_t(lr_end - 1)});
                edge2.cns_supp.push_back({rid, 1 - rstrand, uint32_t(rlen - (lr_end - 1) - 1), uint32_t(rlen - (lr_start + 1) - 1)});
            } else {
                LOG("        [coordinate] could not extract subseq\n");
            }
        }
        if (!edge1.cns_supp.empty()) {
            edge1.head_end = edge2.tail_beg = c1;
            edge1.tail_beg = edge2.head_end = c2;
        } else {
            no_support();
        }
        LOG("\n");
    }
#undef LOG
    // (the reference never closes this file; it is flushed at exit — same bytes)
    if (fp) fclose(fp);
}

// ---------------------------------------------------------------------------------------------------------
// consensus: gather every edge's segments, ONE batched POA call per GPU, scatter the strings back
// ---------------------------------------------------------------------------------------------------------
// substr(spos, epos - spos + 1) on the read (strand 0) or its reverse complement (strand 1), uint32 arithmetic
// as in Assemble.cpp:529-532 (quirk Q7: a wrapped length takes the tail)
static uint32_t segment_length(const SeqStore& reads, const CnsSupp& s) {
    const uint32_t len = reads.len(s.lr_id);
    if (s.spos > len) { fprintf(stderr, "[ERROR] segment start %u beyond read %u of length %u\n", s.spos, s.lr_id, len); exit(EXIT_FAILURE); }
    const uint32_t want = s.epos - s.spos + 1;
    return std::min<uint32_t>(want, len - s.spos);
}
static void write_segment(const SeqStore& reads, const CnsSupp& s, uint32_t cnt, char* out) {
    const uint32_t len = reads.len(s.lr_id);
    const char* r = reads.data(s.lr_id);
    if (s.lr_strand == 0) {
        memcpy(out, r + s.spos, cnt);
    } else {
        // revcomp(read)[spos .. spos+cnt) = complement of read[len-1-spos], read[len-2-spos], ...
        for (uint32_t k = 0; k < cnt; ++k) {
            const char c = r[len - 1 - s.spos - k];
            out[k] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        }
    }
}

static double now_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

int call_consensus(Graph& g, const std::vector<EdgeRef>& edges, const SeqStore& reads, const std::vector<hgpu_t*>& ctxs,
                   const std::string& logpath, bool write_log, unsigned threads) {
    const size_t n = edges.size();
    const double t0 = now_s();
    std::string bases;
    std::vector<uint64_t> seg_off{0};
    std::vector<uint32_t> edge_seg_off{0};
    std::vector<Edge*> e1(n), e2(n);
    std::vector<const CnsSupp*> seg_src;
    for (size_t e = 0; e < n; ++e) {
        const EdgeRef& er = edges[e];
        e1[e] = &g[er.node1].edges[er.rev1][(er.node2 << 1) | er.rev2];
        e2[e] = &g[er.node2].edges[1 - er.rev2][(er.node1 << 1) | (1 - er.rev1)];
        for (const CnsSupp& s : e1[e]->cns_supp) { seg_src.push_back(&s); seg_off.push_back(seg_off.back() + segment_length(reads, s)); }
        edge_seg_off.push_back((uint32_t)(seg_off.size() - 1));
    }
    bases.resize(seg_off.back());
    {   // the copies (and reverse complements) are independent: all host threads
        const size_t n_seg = seg_src.size();
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n_seg / 256 + 1));
        auto work = [&](unsigned t) {
            for (size_t q = n_seg * t / nt; q < n_seg * (t + 1) / nt; ++q)
                write_segment(reads, *seg_src[q], (uint32_t)(seg_off[q + 1] - seg_off[q]), &bases[seg_off[q]]);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    // shard: edges dealt to GPUs by estimated DP cost (sum of len^2-ish), largest first; one host thread per GPU
    const size_t G = std::max<size_t>(1, ctxs.size());
    std::vector<std::vector<uint32_t>> shard(G);
    {
        std::vector<std::pair<double, uint32_t>> cost(n);
        for (size_t e = 0; e < n; ++e) {
            double c = 0, v = 0;
            for (uint32_t s = edge_seg_off[e]; s < edge_seg_off[e + 1]; ++s) {
                const double l = (double)(seg_off[s + 1] - seg_off[s]);
                c += v * l; v = std::max(v, l) + 0.1 * l;
            }
            cost[e] = {c, (uint32_t)e};
        }
        std::sort(cost.begin(), cost.end(), [](const std::pair<double, uint32_t>& a, const std::pair<double, uint32_t>& b) {
            return a.first != b.first ? a.first > b.first : a.second < b.second;
        });
        std::vector<double> load(G, 0);
        for (const auto& c : cost) {
            size_t best = std::min_element(load.begin(), load.end()) - load.begin();
            shard[best].push_back(c.second); load[best] += c.first;
        }
        for (auto& s : shard) std::sort(s.begin(), s.end());
    }
    std::vector<std::string> cons(n);
    std::vector<int> rc(G, 0);
    const double t1 = now_s();
    auto run = [&](size_t gi) {
        const std::vector<uint32_t>& my = shard[gi];
        if (my.empty()) return;
        std::string b; std::vector<uint64_t> so{0}; std::vector<uint32_t> eso{0};
        for (uint32_t e : my) {
            for (uint32_t s = edge_seg_off[e]; s < edge_seg_off[e + 1]; ++s) {
                b.append(bases, seg_off[s], seg_off[s + 1] - seg_off[s]);
                so.push_back(b.size());
            }
            eso.push_back((uint32_t)(so.size() - 1));
        }
        std::vector<uint8_t> out(b.size() + 64);
        std::vector<uint64_t> off(my.size() + 1);
        std::vector<uint32_t> status(my.size());
        hgpu_poa_set_timing(ctxs[gi], 1);
        int r = hgpu_poa_batch(ctxs[gi], (const uint8_t*)b.data(), so.data(), eso.data(), (uint32_t)my.size(), 5, -4, -8, 0,   // Assemble.cpp:8-11
                               out.data(), out.size(), off.data(), status.data());
        if (r != HGPU_OK) { fprintf(stderr, "[ERROR] hgpu_poa_batch (gpu %zu): %s\n", gi, hgpu_last_error(ctxs[gi])); rc[gi] = r; return; }
        hgpu_poa_stats st;
        if (hgpu_poa_get_stats(ctxs[gi], &st) == HGPU_OK)
            fprintf(stderr, "       gpu %zu: %zu edges, %llu alignments (%llu rel16, %llu int32), %.1f Mbases in, %.3e DP cells, kernel %.1f ms (%.0f GCUPS), %llu launches\n",
                    gi, my.size(), (unsigned long long)st.alignments, (unsigned long long)st.alignments_rel16, (unsigned long long)st.alignments_i32, st.bases_in / 1e6, (double)st.cells,
                    st.ms_dp, st.ms_dp > 0 ? st.cells / (st.ms_dp * 1e6) : 0.0, (unsigned long long)st.dp_launches);
        for (size_t k = 0; k < my.size(); ++k) {
            if (status[k] != 0) { fprintf(stderr, "[ERROR] POA failed for edge %u with status %u\n", my[k], status[k]); rc[gi] = HGPU_E_INTERNAL; }
            cons[my[k]].assign((const char*)out.data() + off[k], off[k + 1] - off[k]);
        }
    };
    if (G == 1) run(0);
    else {
        std::vector<std::thread> th;
        for (size_t gi = 0; gi < G; ++gi) th.emplace_back(run, gi);
        for (auto& t : th) t.join();
    }
    for (int r : rc) if (r) return r;
    const double t2 = now_s();
    fprintf(stderr, "       %zu segments gathered in %.2f s, POA calls %.2f s\n", seg_off.size() - 1, t1 - t0, t2 - t1);
    FILE* fp = write_log ? open_write(logpath) : nullptr;
    for (size_t e = 0; e < n; ++e) {
        const EdgeRef& er = edges[e];
        if (fp) {
            fprintf(fp, "calc_cns th_id:%d %u:%c -> %u:%c\n", 0, er.node1, sgn(er.rev1), er.node2, sgn(er.rev2));
            fprintf(fp, "[shared_region] head_end:%u\ttail_beg:%u\n", e1[e]->head_end, e1[e]->tail_beg);
            uint32_t s = edge_seg_off[e];
            for (const CnsSupp& c : e1[e]->cns_supp) {
                fprintf(fp, "        [debug] lr_id:%u lr_len:%u region_start:%u region_end:%u subseq_len:%u\n", c.lr_id, reads.len(c.lr_id), c.spos, c.epos, c.epos - c.spos + 1);
                fprintf(fp, ">%u %c %u %u %u\n", c.lr_id, sgn(c.lr_strand), c.spos, c.epos, c.epos - c.spos + 1);
                fwrite(bases.data() + seg_off[s], 1, seg_off[s + 1] - seg_off[s], fp);
                fputc('\n', fp);
                ++s;
            }
            fprintf(fp, ">CONSENSUS\n%s\n", cons[e].c_str());
        }
        e1[e]->cns_seq = cons[e];
        e2[e]->cns_seq = revcomp(cons[e]);       // empty stays empty (Assemble.cpp:545-556)
    }
    if (fp) fclose(fp);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// stitching
// ---------------------------------------------------------------------------------------------------------
typedef std::pair<uint32_t, uint32_t> NS;        // (node, strand)

static void simple_path_from(const Graph& g, uint32_t src_node, uint32_t src_strand, std::map<uint32_t, Edge>::const_iterator it,
                             std::deque<NS>& path) {                                                  // Assemble.cpp:607-622
    path.clear();
    path.push_back({src_node, src_strand});
    uint32_t node = it->first >> 1, strand = it->first & 1;
    while (true) {
        path.push_back({node, strand});
        if (g[node].edges[strand].empty()) break;
        if (g[node].edges[strand].size() > 1 || g[node].edges[1 - strand].size() > 1) break;
        it = g[node].edges[strand].begin();
        node = it->first >> 1; strand = it->first & 1;
    }
}

static void extract_simple_paths(Graph& g, uint32_t flag, std::vector<std::deque<NS>>& paths) {      // Assemble.cpp:757-810
    for (uint32_t i = 0; i < g.size(); ++i) {
        if (g[i].edges[0].size() == 1 && g[i].edges[1].size() == 1) continue;       // interior of a path
        if (g[i].edges[0].size() > 1 && g[i].edges[1].size() > 1) paths.push_back(std::deque<NS>{{i, 0}});
        for (uint32_t rev = 0; rev < 2; ++rev)
            for (auto it = g[i].edges[rev].begin(); it != g[i].edges[rev].end(); ++it) {
                if (it->second.flag == flag) continue;
                std::deque<NS> path;
                simple_path_from(g, i, rev, it, path);
                for (size_t j = 0; j + 1 < path.size(); ++j) {
                    const uint32_t n1 = path[j].first, r1 = path[j].second, n2 = path[j + 1].first, r2 = path[j + 1].second;
                    g[n1].edges[r1][(n2 << 1) | r2].flag = flag;
                    g[n2].edges[1 - r2][(n1 << 1) | (1 - r1)].flag = flag;
                }
                // branching ends belong to no path
                if (g[path.front().first].edges[path.front().second].size() > 1) path.pop_front();
                if (g[path.back().first].edges[1 - path.back().second].size() > 1) path.pop_back();
                if (!path.empty()) paths.push_back(path);
            }
    }
}

static void assemble_path(const std::deque<NS>& path, Graph& g, const ContigStore& contigs, int& nb_ctg, FILE* fa, FILE* ann, FILE* log) {
    auto cstr = [&](uint32_t id) { return std::string(contigs.data(id), contigs.len(id)); };
    if (path.size() == 1) {                                                                           // Assemble.cpp:626-635
        const uint32_t c = path.front().first, s = path.front().second;
        const std::string str = cstr(c);
        fprintf(log, ">%d from:%u:%c to:%u:%c\n%s\n\n", nb_ctg, c, sgn(s), c, sgn(s), str.c_str());
        fprintf(fa, ">%d from:%u:%c to:%u:%c\n%s\n", nb_ctg, c, sgn(s), c, sgn(s), str.c_str());
        ++nb_ctg;
        return;
    }
    std::string assembled;
    uint32_t src = path[0].first, src_strand = path[0].second;
    uint32_t start = src_strand == 0 ? 0 : contigs.len(src) - 1;                  // where the current contig resumes
    const uint32_t tgt = path.back().first, tgt_strand = path.back().second;
    size_t i;
    for (i = 0; i + 1 < path.size(); ++i) {                                                           // Assemble.cpp:667-736
        const uint32_t c1 = path[i].first, s1 = path[i].second, c2 = path[i + 1].first, s2 = path[i + 1].second;
        const std::string c1s = cstr(c1);
        Edge& edge = g[c1].edges[s1][(c2 << 1) | s2];
        std::string prefix;
        if (edge.cns_supp.empty()) {       // no read bridges this anchor pair: end the contig here, start a new one
            fprintf(log, "[breaking] contig1_len:%zu    contig1_start:%u    prev_end:%u     next_beg:%u\n", c1s.size(), start, edge.head_end, edge.tail_beg);
            if (s1 == 0) {
                prefix = c1s.substr(start);
                fprintf(ann, "%d\t%zu\t%zu\tctg\t+\t%u\t%zu\t%u\t%zu\n", nb_ctg, assembled.size(), assembled.size() + prefix.size(), c1, c1s.size(), start, c1s.size());
            } else {
                prefix = c1s.substr(0, start + 1);
                fprintf(ann, "%d\t%zu\t%zu\tctg\t-\t%u\t%zu\t%u\t%u\n", nb_ctg, assembled.size(), assembled.size() + prefix.size(), c1, c1s.size(), 0, start + 1);
                prefix = revcomp(prefix);
            }
            assembled += prefix;
            fprintf(log, ">%d from:%u:%c to:%u:%c\n%s\n\n", nb_ctg, src, sgn(src_strand), c1, sgn(s1), assembled.c_str());
            fprintf(fa, ">%d from:%u:%c to:%u:%c\n%s\n", nb_ctg, src, sgn(src_strand), c1, sgn(s1), assembled.c_str());
            ++nb_ctg;
            assembled.clear();
            src = c2; src_strand = s2;
            start = src_strand == 0 ? 0 : contigs.len(src) - 1;
            fprintf(stderr, "[WARNING] breaking assembly for path %u:%c --> %u:%c between anchors %u:%c --> %u:%c\n", src, sgn(src_strand), tgt, sgn(tgt_strand), c1, sgn(s1), c2, sgn(s2));
        } else {
            fprintf(log, "[stitching] contig1_len:%zu    contig1_start:%u    prev_end:%u     next_beg:%u\n", c1s.size(), start, edge.head_end, edge.tail_beg);
            if (s1 == 0) {
                prefix = c1s.substr(start, edge.head_end - start + 1);
                fprintf(ann, "%d\t%zu\t%zu\tctg\t+\t%u\t%zu\t%u\t%zu\n", nb_ctg, assembled.size(), assembled.size() + prefix.size(), c1, c1s.size(), start, start + prefix.size());
            } else {
                prefix = c1s.substr(edge.head_end, start - edge.head_end + 1);
                fprintf(ann, "%d\t%zu\t%zu\tctg\t-\t%u\t%zu\t%u\t%zu\n", nb_ctg, assembled.size(), assembled.size() + prefix.size(), c1, c1s.size(), edge.head_end, edge.head_end + prefix.size());
                prefix = revcomp(prefix);
            }
            assembled += prefix;
            fprintf(ann, "%d\t%zu\t%zu\tcns\t%zu\t%zu\n", nb_ctg, assembled.size(), assembled.size() + edge.cns_seq.size(), edge.cns_seq.size(), edge.cns_supp.size());
            assembled += edge.cns_seq;
            start = edge.tail_beg;
        }
    }
    const uint32_t c2 = path[i].first, s2 = path[i].second;                                           // Assemble.cpp:737-754
    const std::string c2s = cstr(c2);
    std::string suffix;
    if (s2 == 0) {
        suffix = c2s.substr(start);
        fprintf(ann, "%d\t%zu\t%zu\tctg\t+\t%u\t%zu\t%u\t%zu\n", nb_ctg, assembled.size(), assembled.size() + suffix.size(), c2, c2s.size(), start, c2s.size());
    } else {
        suffix = c2s.substr(0, start + 1);
        fprintf(ann, "%d\t%zu\t%zu\tctg\t-\t%u\t%zu\t%u\t%u\n", nb_ctg, assembled.size(), assembled.size() + suffix.size(), c2, c2s.size(), 0, start + 1);
        suffix = revcomp(suffix);
    }
    assembled += suffix;
    fprintf(log, ">%d from:%u:%c to:%u:%c\n%s\n\n", nb_ctg, src, sgn(src_strand), c2, sgn(s2), assembled.c_str());
    fprintf(fa, ">%d from:%u:%c to:%u:%c\n%s\n", nb_ctg, src, sgn(src_strand), c2, sgn(s2), assembled.c_str());
    ++nb_ctg;
}

void write_assembly(Graph& g, const ContigStore& contigs, const std::string& out_dir) {              // Assemble.cpp:1045-1077
    FILE* fa = open_write(out_dir + "/asm.final.fa");
    FILE* ann = open_write(out_dir + "/asm.final.ann");
    FILE* log = open_write(out_dir + "/log_asmfinal.txt");
    std::vector<std::deque<NS>> paths;
    extract_simple_paths(g, 21, paths);
    for (uint32_t i = 0; i < paths.size(); ++i)
        fprintf(log, "simple_path %u size:%zu\tfrom:%u:%c\tto:%u:%c\n", i, paths[i].size(), paths[i].front().first, sgn(paths[i].front().second),
                paths[i].back().first, sgn(paths[i].back().second));
    int nb_ctg = 0;
    for (const auto& p : paths) assemble_path(p, g, contigs, nb_ctg, fa, ann, log);
    fclose(log); fclose(ann); fclose(fa);
}

}  // namespace haslr
