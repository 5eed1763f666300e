// Edge coordinates, consensus calling (GPU) and stitching.
//
// Restates the behaviour of the reference's src/haslr_assemble/src/Assemble.cpp: the canonical edge enumeration
// (:365-434), the edge coordinates (:24-363 — the interval sweeps and CIGAR walks run on the GPU, one hgpu_edge_coords
// call for all edges; this file turns the answer into cns_supp lists and the reference's log), the per-edge consensus
// call (:479-560 — here ONE batched hgpu_poa_batch call instead of a SPOA engine per edge per thread) and the
// simple-path stitching (:607-810,1045-1077).
//
// Provenance note: asm.final.fa / .ann and the logs must match the reference byte for byte, so assemble_path and the log block of
// calc_edge_coordinates follow the statement order and format strings of asm_assemble_single_path / asm_calc_single_edge_coordinates
// (Assemble.cpp:624-755,157-363, GPLv3) closely: a behaviour-identical host port, not a redesign. It lives only in the host
// binary / libhaslr_path.so; the CUDA library (libhaslr_b200.so) contains no reference-derived text.
#include <algorithm>
#include <cstring>
#include <deque>
#include <iterator>
#include <memory>
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
// edge coordinates: one hgpu_edge_coords call for all edges (best supported intervals on both anchors, the reads in
// both best sets, the CIGAR walks to the anchor positions — Assemble.cpp:24-363); the host fills the graph's
// cns_supp / head_end / tail_beg from the answer and writes the reference's log_coordinate.txt from it.
// ---------------------------------------------------------------------------------------------------------
int calc_edge_coordinates(Graph& g, const std::vector<EdgeRef>& edges, const ContigStore& contigs, const SeqStore& reads,
                          const CompactReads& cl, hgpu_t* ctx, const std::string& logpath) {
    if (edges.empty()) return 0;           // the reference returns before opening the log (quirk Q12)
    const size_t n = edges.size();
    std::vector<uint8_t> edge_rev(n);
    std::vector<uint32_t> supp_off(n + 1, 0), read_len(reads.size());
    std::vector<hgpu_edge_supp> supp;
    std::vector<Edge*> e1(n), e2(n);
    for (size_t e = 0; e < n; ++e) {
        const EdgeRef& er = edges[e];
        e1[e] = &g[er.node1].edges[er.rev1][(er.node2 << 1) | er.rev2];
        e2[e] = &g[er.node2].edges[1 - er.rev2][(er.node1 << 1) | (1 - er.rev1)];
        edge_rev[e] = (uint8_t)(er.rev1 | (er.rev2 << 1));
        for (const EdgeSupp& s : e1[e]->edge_supp) supp.push_back({s.lr_id, s.cmp_head_id, s.cmp_tail_id});
        supp_off[e + 1] = (uint32_t)supp.size();
    }
    for (size_t r = 0; r < reads.size(); ++r) read_len[r] = reads.len(r);
    std::vector<hgpu_edge_coord> oe(n);
    std::vector<hgpu_supp_coord> os(supp.size() + 1);
    // compact reads and the hit table (strands, run-length CIGARs) are still on the device: only edges and supports go up
    const int rc = hgpu_edge_coords_dev(ctx, (uint32_t)n, edge_rev.data(), supp_off.data(), supp.data(), read_len.data(), (uint32_t)reads.size(),
                                        oe.data(), os.data());
    if (rc != HGPU_OK) { fprintf(stderr, "[ERROR] hgpu_edge_coords_dev: %s\n", hgpu_last_error(ctx)); return rc; }

    FILE* fp = logpath.empty() ? nullptr : open_write(logpath);
#define LOG(...) do { if (fp) fprintf(fp, __VA_ARGS__); } while (0)
    auto elem = [&](uint32_t rid, uint32_t cmp) -> const hgpu_cl_elem& { return cl.elems[cl.off[rid] + cmp]; };
    auto erev = [&](uint32_t rid, uint32_t cmp) -> uint32_t { return cl.rev[cl.off[rid] + cmp]; };
    for (size_t e = 0; e < n; ++e) {
        const uint32_t node1 = edges[e].node1, rev1 = edges[e].rev1, node2 = edges[e].node2, rev2 = edges[e].rev2;
        Edge& edge1 = *e1[e];
        Edge& edge2 = *e2[e];
        const std::vector<EdgeSupp>& es = edge1.edge_supp;
        const hgpu_edge_coord& c = oe[e];
        const hgpu_supp_coord* sc = os.data() + supp_off[e];
        LOG("calc_coords th_id:%d %u:%c -> %u:%c\n", 0, node1, sgn(rev1), node2, sgn(rev2));
        LOG("edge      %u:%c -> %u:%c\n", node1, sgn(rev1), node2, sgn(rev2));
        LOG("edge_twin %u:%c -> %u:%c\n", node2, sgn(1 - rev2), node1, sgn(1 - rev1));
        LOG("\tedge_supp size:%zu\n", es.size());
        if (fp) for (uint32_t i = 0; i < es.size(); ++i) {
            const hgpu_cl_elem& h = elem(es[i].lr_id, es[i].cmp_head_id);
            const hgpu_cl_elem& t = elem(es[i].lr_id, es[i].cmp_tail_id);
            LOG("\tsupp_detail head\t%u\t%u\t%c\ttail\t%u\t%u\t%c\n", h.t_start, h.t_end, sgn(erev(es[i].lr_id, es[i].cmp_head_id)), t.t_start, t.t_end,
                sgn(erev(es[i].lr_id, es[i].cmp_tail_id)));
        }
        LOG("    @@@ best interval contig1 %u %u\n", c.int1_lo, c.int1_hi);
        LOG("    @@@ best_interval contig2 %u %u\n", c.int2_lo, c.int2_hi);
        LOG("coordinates contig1_pos: %u\tcontig2_pos: %u\n", c.c1, c.c2);
        LOG("supproting_lr: %lu\n", (unsigned long)c.n_best);
        auto no_support = [&]() {
            edge1.cns_supp.clear(); edge2.cns_supp.clear();
            edge1.head_end = edge2.tail_beg = (rev1 == 0 ? contigs.len(node1) - 1 : 0);
            edge1.tail_beg = edge2.head_end = (rev2 == 0 ? 0 : contigs.len(node2) - 1);
        };
        if (c.n_best == 0) { no_support(); continue; }           // the reference returns here, without the blank line
        edge1.cns_supp.clear(); edge2.cns_supp.clear();
        for (uint32_t i = 0; i < es.size(); ++i) {
            if (!sc[i].in_best) continue;
            const uint32_t rid = es[i].lr_id, rlen = reads.len(rid), rstrand = sc[i].lr_strand;
            LOG("    +++ lr:%u len:%u strand:%c\n", rid, rlen, sgn(rstrand));
            LOG("        case %d\n", rev1 == 0 ? (rstrand ? 5 : 1) : (rstrand ? 6 : 2));
            LOG("        case %d\n", rev2 == 0 ? (rstrand ? 7 : 3) : (rstrand ? 8 : 4));
            const long long lr_start = sc[i].lr_start, lr_end = sc[i].lr_end;
            if (lr_start != -1 && lr_end != -1) {
                LOG("        [coordinate] subseq_len:%lld lr_start:%lld lr_end:%lld\n", lr_end - lr_start - 1, lr_start + 1, lr_end - 1);
                edge1.cns_supp.push_back({rid, rstrand, uint32_t(lr_start + 1), uint32_t(lr_end - 1)});
                edge2.cns_supp.push_back({rid, 1 - rstrand, uint32_t(rlen - (lr_end - 1) - 1), uint32_t(rlen - (lr_start + 1) - 1)});
            } else {
                LOG("        [coordinate] could not extract subseq\n");
            }
        }
        if (!edge1.cns_supp.empty()) {
            edge1.head_end = edge2.tail_beg = c.c1;
            edge1.tail_beg = edge2.head_end = c.c2;
        } else {
            no_support();
        }
        LOG("\n");
    }
#undef LOG
    // (the reference never closes this file; it is flushed at exit — same bytes)
    if (fp) fclose(fp);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// consensus: gather every edge's segments, ONE batched POA call per GPU, scatter the strings back
// ---------------------------------------------------------------------------------------------------------
// substr(spos, epos - spos + 1) on the read (strand 0) or its reverse complement (strand 1), uint32 arithmetic
// as in Assemble.cpp:529-532 (quirk Q7: a wrapped length takes the tail)
uint32_t segment_length(const SeqStore& reads, const CnsSupp& s) {
    const uint32_t len = reads.len(s.lr_id);
    if (s.spos > len) { fprintf(stderr, "[ERROR] segment start %u beyond read %u of length %u\n", s.spos, s.lr_id, len); exit(EXIT_FAILURE); }
    const uint32_t want = s.epos - s.spos + 1;
    return std::min<uint32_t>(want, len - s.spos);
}
namespace { struct CompTable { char t[256]; CompTable() { for (int i = 0; i < 256; ++i) t[i] = 'A'; t['A'] = 'T'; t['C'] = 'G'; t['G'] = 'C'; t['T'] = 'A'; } } COMP; }
void write_segment(const SeqStore& reads, const CnsSupp& s, uint32_t cnt, char* out) {
    const uint32_t len = reads.len(s.lr_id);
    const char* r = reads.data(s.lr_id);
    if (s.lr_strand == 0) {
        memcpy(out, r + s.spos, cnt);
    } else {
        // revcomp(read)[spos .. spos+cnt) = complement of read[len-1-spos], read[len-2-spos], ...
        const unsigned char* p = reinterpret_cast<const unsigned char*>(r) + (len - 1 - s.spos);
        for (uint32_t k = 0; k < cnt; ++k) out[k] = COMP.t[p[-(long)k]];
    }
}

static double now_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

int call_consensus(Graph& g, const std::vector<EdgeRef>& edges, const SeqStore& reads, const std::vector<hgpu_t*>& ctxs,
                   const std::string& logpath, bool write_log, unsigned threads, uint64_t* bases_in) {
    const size_t n = edges.size();
    const double t0 = now_s();
    std::vector<uint64_t> seg_off{0};            // global segment offsets (lengths only: the bases go straight into the per-GPU buffers)
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
    if (bases_in) *bases_in = seg_off.back();
    // shard: edges dealt to GPUs by estimated DP cost (sum of len^2-ish), largest first (LPT); one host thread per GPU
    const size_t G = std::max<size_t>(1, ctxs.size());
    std::vector<std::vector<uint32_t>> shard(G);
    if (G == 1) {
        shard[0].resize(n);
        for (size_t e = 0; e < n; ++e) shard[0][e] = (uint32_t)e;
    } else {
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
    // per-GPU segment tables; every segment is copied (or reverse-complemented) ONCE, from the read into its GPU's buffer
    // the segment bytes live in uninitialised storage: a std::string would zero-fill gigabytes on one thread before the gather threads
    // touch (and page in) their own parts
    // ... and that storage is the context's page-locked staging buffer, so the upload is one DMA at link speed (a pageable source
    // goes through the driver's bounce buffer: 25-50 ms for config 2's 133 MB against 3 ms); plain memory if it cannot be had
    struct Shard { char* b = nullptr; std::unique_ptr<char[]> own; uint64_t n = 0; std::vector<uint64_t> so{0}; std::vector<uint32_t> eso{0}; std::vector<uint32_t> seg; };
    std::vector<Shard> sh(G);
    std::vector<std::pair<uint32_t, uint64_t>> seg_home(seg_src.size());      // segment -> (gpu, offset in its buffer), for the log
    for (size_t gi = 0; gi < G; ++gi) {
        Shard& S = sh[gi];
        for (uint32_t e : shard[gi]) {
            for (uint32_t s = edge_seg_off[e]; s < edge_seg_off[e + 1]; ++s) {
                seg_home[s] = {(uint32_t)gi, S.so.back()};
                S.seg.push_back(s);
                S.so.push_back(S.so.back() + (seg_off[s + 1] - seg_off[s]));
            }
            S.eso.push_back((uint32_t)(S.so.size() - 1));
        }
        S.n = S.so.back();
        void* pinned = nullptr;
        if (!shard[gi].empty() && hgpu_host_staging(ctxs[gi], 0, S.n + 64, &pinned) == HGPU_OK) S.b = (char*)pinned;
        else { S.own.reset(new char[S.n + 64]); S.b = S.own.get(); }
    }
    {   // the copies are independent: all host threads
        const size_t n_seg = seg_src.size();
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n_seg / 256 + 1));
        auto work = [&](unsigned t) {
            for (size_t q = n_seg * t / nt; q < n_seg * (t + 1) / nt; ++q)
                write_segment(reads, *seg_src[q], (uint32_t)(seg_off[q + 1] - seg_off[q]), sh[seg_home[q].first].b + seg_home[q].second);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    std::vector<std::string> cons(n);
    std::vector<int> rc(G, 0);
    const double t1 = now_s();
    auto run = [&](size_t gi) {
        const std::vector<uint32_t>& my = shard[gi];
        if (my.empty()) return;
        Shard& S = sh[gi];
        const uint64_t out_cap = S.n + 64;                      // worst case (every consensus as long as its segments together); never touched beyond the result
        std::unique_ptr<uint8_t[]> out(new uint8_t[out_cap]);
        std::vector<uint64_t> off(my.size() + 1);
        std::vector<uint32_t> status(my.size());
        hgpu_poa_set_timing(ctxs[gi], 1);      // one event pair per scheduling pass, read after the pass's own synchronisation
        int r = hgpu_poa_batch(ctxs[gi], (const uint8_t*)S.b, S.so.data(), S.eso.data(), (uint32_t)my.size(), 5, -4, -8, 0,   // Assemble.cpp:8-11
                               out.get(), out_cap, off.data(), status.data());
        if (r != HGPU_OK) { fprintf(stderr, "[ERROR] hgpu_poa_batch (gpu %zu): %s\n", gi, hgpu_last_error(ctxs[gi])); rc[gi] = r; return; }
        hgpu_poa_stats st;
        if (hgpu_poa_get_stats(ctxs[gi], &st) == HGPU_OK)
            fprintf(stderr, "       gpu %zu: %zu edges, %llu alignments (%llu rel16, %llu int32), %.1f Mbases in, %.3e DP cells, kernel %.1f ms (%.0f GCUPS), %llu launches\n",
                    gi, my.size(), (unsigned long long)st.alignments, (unsigned long long)st.alignments_rel16, (unsigned long long)st.alignments_i32, st.bases_in / 1e6, (double)st.cells,
                    st.ms_dp, st.ms_dp > 0 ? st.cells / (st.ms_dp * 1e6) : 0.0, (unsigned long long)st.dp_launches);
        for (size_t k = 0; k < my.size(); ++k) {
            if (status[k] != 0) { fprintf(stderr, "[ERROR] POA failed for edge %u with status %u\n", my[k], status[k]); rc[gi] = HGPU_E_INTERNAL; }
            cons[my[k]].assign((const char*)out.get() + off[k], off[k + 1] - off[k]);
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
                fwrite(sh[seg_home[s].first].b + seg_home[s].second, 1, seg_off[s + 1] - seg_off[s], fp);
                fputc('\n', fp);
                ++s;
            }
            fprintf(fp, ">CONSENSUS\n%s\n", cons[e].c_str());
        }
    }
    if (fp) fclose(fp);
    {   // edge and twin get the string and its reverse complement; every undirected edge once, so the edges are independent: all host threads
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n / 64 + 1));
        auto work = [&](unsigned t) {
            for (size_t e = n * t / nt; e < n * (t + 1) / nt; ++e) {
                std::string rc = revcomp(cons[e]);   // empty stays empty (Assemble.cpp:545-556)
                e1[e]->cns_seq = std::move(cons[e]);
                e2[e]->cns_seq = std::move(rc);      // after the edge's own, as in the reference (an edge that is its own twin keeps the reverse complement)
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    fprintf(stderr, "       consensus strings scattered to the graph in %.3f s\n", now_s() - t2);
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
