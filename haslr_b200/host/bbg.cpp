// Backbone graph on the host: filled from the GPU edge table, written as GFA/stat, simplified.
//
// The graph container mirrors the reference's (vector of nodes, two ordered maps per node keyed by
// (other_node << 1) | other_strand, every edge stored with its twin) because the cleaning passes and all writers
// are defined in terms of that iteration order. The passes restate the behaviour of the reference's
// src/haslr_assemble/src/Cleaning.cpp and Backbone_graph.cpp; each function cites what it mirrors.
//
// Provenance note: the drop-in contract is byte-identical logs and GFA files, so detect_super_bubble / clean_super_bubbles and
// clean_small_bubbles below follow the statement order of the reference's functions (Cleaning.cpp:488-560,563-648,7-57, GPLv3)
// closely, down to its arithmetic quirks. They are a behaviour-identical host port, not a redesign, live only in the host
// binary / libhaslr_path.so, and are never linked into the CUDA library (libhaslr_b200.so), which contains no reference-derived text.
#include <algorithm>
#include <queue>
#include <set>
#include <stack>
#include <tuple>
#include <unordered_map>

#include "haslr.hpp"

namespace haslr {

static inline char sgn(uint32_t s) { return s ? '-' : '+'; }

// entries come sorted by key64 = (from << 32) | to, which is the nested iteration order node -> strand -> map key
void graph_from_edge_table(Graph& g, size_t n_contigs, const std::vector<uint64_t>& key, const std::vector<uint32_t>& supp_off,
                           const std::vector<hgpu_edge_supp>& supp, const std::vector<uint8_t>* keep) {
    g.assign(n_contigs, Node());
    for (size_t e = 0; e < key.size(); ++e) {
        if (keep && !(*keep)[e]) continue;
        const uint32_t from = (uint32_t)(key[e] >> 32), to = (uint32_t)key[e];
        auto& m = g[from >> 1].edges[from & 1];
        Edge& ed = m.emplace_hint(m.end(), to, Edge())->second;
        ed.edge_supp.reserve(supp_off[e + 1] - supp_off[e]);
        for (uint32_t s = supp_off[e]; s < supp_off[e + 1]; ++s)
            ed.edge_supp.push_back({supp[s].lr_id_strand & 0x7FFFFFFFu, supp[s].lr_id_strand >> 31, supp[s].cmp_head, supp[s].cmp_tail});
    }
}

static void remove_edge(Graph& g, uint32_t node1, uint32_t rev1, uint32_t node2, uint32_t rev2) {   // Backbone_graph.cpp:44-50
    g[node1].edges[rev1].erase((node2 << 1) | rev2);
    g[node2].edges[1 - rev2].erase((node1 << 1) | (1 - rev1));
}

int remove_weak_edges(Graph& g, uint32_t min_edge_sup) {                                             // Backbone_graph.cpp:348-375
    int removed = 0;
    for (uint32_t i = 0; i < g.size(); ++i)
        for (uint32_t rev1 = 0; rev1 < 2; ++rev1)
            for (auto it = g[i].edges[rev1].begin(); it != g[i].edges[rev1].end();) {
                if (it->second.edge_supp.size() < min_edge_sup) {
                    const uint32_t node2 = it->first >> 1, rev2 = it->first & 1;
                    it = g[i].edges[rev1].erase(it);
                    g[node2].edges[1 - rev2].erase((i << 1) | (1 - rev1));
                    ++removed;
                } else {
                    ++it;
                }
            }
    return removed;
}

void write_compact(const CompactReads& cl, const std::string& path) {          // Longread.cpp:675-693
    FILE* fp = open_write(path);
    for (size_t r = 0; r + 1 < cl.off.size(); ++r) {
        fprintf(fp, ">%zu\t", r);
        for (uint32_t j = cl.off[r]; j < cl.off[r + 1]; ++j) {
            const hgpu_cl_elem& e = cl.elems[j];
            fprintf(fp, "%u-%u:%u:%c:%u-%u\t", e.q_start, e.q_end, cl.tid[j], sgn(cl.rev[j]), e.t_start, e.t_end);
        }
        fprintf(fp, "\n");
    }
    fclose(fp);
}

void write_stats(const Graph& g, const ContigStore& contigs, const std::string& path) {             // Backbone_graph.cpp:595-659
    FILE* fp = open_write(path);
    const uint32_t num = (uint32_t)g.size();
    uint32_t nb_node = 0, nb_edge = 0;
    for (uint32_t i = 0; i < num; ++i) {
        nb_node += (g[i].edges[0].size() > 0 || g[i].edges[1].size() > 0);
        nb_edge += (uint32_t)(g[i].edges[0].size() + g[i].edges[1].size());
    }
    fprintf(fp, "nodes: %d\n", nb_node);
    fprintf(fp, "edges: %d\n", nb_edge / 2);
    std::vector<bool> visited(num, false);
    std::vector<std::tuple<uint32_t, uint32_t, uint32_t>> comps;     // (bases, nodes, first node), 32-bit like the reference's tuple
    for (uint32_t i = 0; i < num; ++i) {
        if (visited[i] || (g[i].edges[0].empty() && g[i].edges[1].empty())) continue;
        uint64_t cc_size = contigs.len(i), cc_node = 1;
        std::queue<uint32_t> q;
        q.push(i); visited[i] = true;
        while (!q.empty()) {
            const uint32_t cur = q.front(); q.pop();
            for (int rev = 0; rev < 2; ++rev)
                for (const auto& kv : g[cur].edges[rev]) {
                    const uint32_t nx = kv.first >> 1;
                    if (!visited[nx]) { q.push(nx); ++cc_node; cc_size += contigs.len(nx); visited[nx] = true; }
                }
        }
        comps.push_back(std::make_tuple((uint32_t)cc_size, (uint32_t)cc_node, i));
    }
    // descending by size only; equal sizes stay in whatever order std::sort leaves them (as in the reference)
    std::sort(comps.begin(), comps.end(), [](const std::tuple<uint32_t, uint32_t, uint32_t>& a, const std::tuple<uint32_t, uint32_t, uint32_t>& b) {
        return std::get<0>(a) > std::get<0>(b);
    });
    fprintf(fp, "connected_components: %zu\n", comps.size());
    for (uint32_t i = 0; i < comps.size(); ++i)
        fprintf(fp, "\tcomponent:%u\tsize:%u\tnodes:%u\trepresentative:%u\n", i, std::get<0>(comps[i]), std::get<1>(comps[i]), std::get<2>(comps[i]));
    fclose(fp);
}

void write_gfa(const Graph& g, const ContigStore& contigs, const std::string& path) {               // Backbone_graph.cpp:540-588
    FILE* fp = open_write(path);
    std::set<uint32_t> to_print;
    for (uint32_t i = 0; i < g.size(); ++i)
        for (int rev = 0; rev < 2; ++rev)
            for (const auto& kv : g[i].edges[rev]) { to_print.insert(i); to_print.insert(kv.first >> 1); }
    for (uint32_t id : to_print) {
        fprintf(fp, "S\t%u\t", id);
        fwrite(contigs.data(id), 1, contigs.len(id), fp);
        fprintf(fp, "\tLN:i:%zu\tKC:i:%u\n", (size_t)contigs.len(id), contigs.kmer_count[id]);
    }
    for (uint32_t i = 0; i < g.size(); ++i)
        for (int rev = 0; rev < 2; ++rev)
            for (const auto& kv : g[i].edges[rev])
                fprintf(fp, "L\t%u\t%c\t%u\t%c\t0M\n", i, sgn(rev), kv.first >> 1, sgn(kv.first & 1));
    fclose(fp);
}

void report_branching(const Graph& g, const std::string& logpath) {                                  // Backbone_graph.cpp:682-694
    FILE* fp = open_write(logpath);
    for (uint32_t i = 0; i < g.size(); ++i)
        if (g[i].edges[0].size() >= 2 || g[i].edges[1].size() >= 2)
            fprintf(fp, "node:%u\tincoming:%zu\toutgoing:%zu\n", i, g[i].edges[0].size(), g[i].edges[1].size());   // (labels swapped in the reference too)
    fclose(fp);
}

// ---------------------------------------------------------------------------------------------------------
// simple path from (src_node, src_strand) through edge `it`, at most max_depth edges — Backbone_graph.cpp:378-402
// path elements are (node, strand); cov = mean support of the edges walked
// ---------------------------------------------------------------------------------------------------------
typedef std::pair<uint32_t, uint32_t> NS;    // (node, strand)
static bool simple_path_from(const Graph& g, uint32_t src_node, uint32_t src_strand, std::map<uint32_t, Edge>::const_iterator it,
                             int max_depth, std::vector<NS>& path, float& cov) {
    path.clear();
    cov = 0;
    path.push_back({src_node, src_strand});
    uint32_t node = it->first >> 1, strand = it->first & 1;
    int depth = 1;
    while (depth <= max_depth) {
        path.push_back({node, strand});
        cov += it->second.edge_supp.size();
        if (g[node].edges[strand].empty()) break;
        if (g[node].edges[strand].size() > 1 || g[node].edges[1 - strand].size() > 1) break;
        it = g[node].edges[strand].begin();
        node = it->first >> 1; strand = it->first & 1;
        ++depth;
    }
    if (depth > max_depth) return false;
    cov = cov / depth;
    return true;
}

int clean_tips(Graph& g, int max_depth, const std::string& logpath) {                               // Cleaning.cpp:59-96
    FILE* fp = max_depth == 1 ? open_write(logpath) : open_append(logpath);
    int removed = 0;
    const uint32_t num = (uint32_t)g.size();
    for (uint32_t i = 0; i < num; ++i) {
        uint32_t src_strand;
        if (g[i].edges[1].empty() && g[i].edges[0].size() == 1) src_strand = 0;
        else if (g[i].edges[1].size() == 1 && g[i].edges[0].empty()) src_strand = 1;
        else continue;
        std::vector<NS> path; float cov;
        if (simple_path_from(g, i, src_strand, g[i].edges[src_strand].begin(), max_depth, path, cov)) {
            if (g[path.back().first].edges[path.back().second].empty()) continue;       // isolated chain, not a tip
            fprintf(fp, "tip_len:%zu\t%u:%c -> %u:%c\n", path.size() - 1, path.front().first, sgn(path.front().second), path.back().first, sgn(path.back().second));
            for (size_t j = 0; j + 1 < path.size(); ++j) remove_edge(g, path[j].first, path[j].second, path[j + 1].first, path[j + 1].second);
            ++removed;
        }
    }
    fclose(fp);
    return removed;
}

// two-way bubbles whose arms are simple paths of at most max_depth edges — Cleaning.cpp:98-184 (clean_simple_bubbles_old)
int clean_simple_bubbles(Graph& g, int max_depth, const std::string& logpath) {
    FILE* fp = open_write(logpath);
    int removed = 0;
    const uint32_t num = (uint32_t)g.size();
    for (uint32_t i = 0; i < num; ++i) {
        if (g[i].edges[0].size() < 2 && g[i].edges[1].size() < 2) continue;
        bool again = false;
        for (uint32_t side = 0; side < 2 && !again; ++side) {
            if (g[i].edges[side].size() != 2) continue;
            std::vector<NS> p1, p2; float c1, c2;
            auto it = g[i].edges[side].begin();
            const bool f1 = simple_path_from(g, i, side, it, max_depth, p1, c1);
            ++it;
            const bool f2 = simple_path_from(g, i, side, it, max_depth, p2, c2);
            if (!(f1 && f2) || p1.back() != p2.back()) continue;
            fprintf(fp, "simple_bubble cov:%.2lf ", c1);
            for (const NS& x : p1) fprintf(fp, "%u:%c ", x.first, sgn(x.second));
            fprintf(fp, "\n              cov:%.2lf ", c2);
            for (const NS& x : p2) fprintf(fp, "%u:%c ", x.first, sgn(x.second));
            fprintf(fp, "\n");
            const std::vector<NS>& drop = c1 < c2 ? p1 : p2;
            for (size_t j = 0; j + 1 < drop.size(); ++j) remove_edge(g, drop[j].first, drop[j].second, drop[j + 1].first, drop[j + 1].second);
            ++removed;
            again = true;                  // the reference re-examines the same node (its `i--; continue;`)
        }
        if (again) --i;
    }
    fclose(fp);
    return removed;
}

// miniasm-style bubble detection from vertex (src_node, src_rev) — Cleaning.cpp:488-560 (detect_super_bubble)
static bool detect_super_bubble(const Graph& g, uint32_t src_node, uint32_t src_rev, std::vector<uint32_t>& best_path,
                                std::set<std::pair<uint32_t, uint32_t>>& bubble_edges) {
    std::stack<uint32_t> S;
    S.push((src_node << 1) | src_rev);
    std::unordered_map<uint32_t, int32_t> visited, pending;            // pending = unvisited incoming edges
    std::unordered_map<uint32_t, std::vector<uint32_t>> path;          // best-supported path up to a vertex
    std::unordered_map<uint32_t, uint32_t> support;                    // its summed edge support
    visited[S.top()] = 1;
    path[S.top()].push_back(S.top());
    support[S.top()] = 0;
    int p = 0;                                                          // visited vertices never pushed
    while (!S.empty()) {
        const uint32_t v = S.top();
        const uint32_t cn = v >> 1, cr = v & 1;
        S.pop();
        for (const auto& kv : g[cn].edges[cr]) {
            bubble_edges.insert({v, kv.first});
            const uint32_t nn = kv.first >> 1, nr = kv.first & 1;
            const uint32_t ns = (uint32_t)kv.second.edge_supp.size();
            const uint32_t w = kv.first;
            if (nn == cn) return false;                                 // a cycle through the current node
            if (visited.count(w) == 0) {
                pending[w] = (int32_t)g[nn].edges[1 - nr].size();
                visited[w] = 1;
                ++p;
            }
            // same arithmetic as the reference, including its division by (|path(v)| - 1)
            if (support.count(w) == 0 ||
                double(support[v] + ns) / path[v].size() > double(support[w]) / (path[v].size() - 1)) {
                support[w] = support[v] + ns;
                path[w] = path[v];
                path[w].push_back(w);
            }
            pending[w]--;
            if (pending[w] == 0 && !g[nn].edges[nr].empty()) { S.push(w); --p; }
        }
        if (S.size() == 1 && p == 0) { best_path = path[S.top()]; return true; }
    }
    return false;
}

int clean_super_bubbles(Graph& g, const std::string& logpath) {                                     // Cleaning.cpp:563-648
    FILE* fp = open_write(logpath);
    int removed = 0;
    const uint32_t num = (uint32_t)g.size();
    for (uint32_t i = 0; i < num; ++i) {
        if (g[i].edges[0].size() < 2 && g[i].edges[1].size() < 2) continue;
        bool again = false;
        for (uint32_t side = 0; side < 2 && !again; ++side) {
            if (g[i].edges[side].size() < 2) continue;
            std::vector<uint32_t> best;
            std::set<std::pair<uint32_t, uint32_t>> edges;
            if (!detect_super_bubble(g, i, side, best, edges)) continue;
            fprintf(fp, "bubble_src %u:%c\tbubble_sink %u:%c\n", i, sgn(side), best.back() >> 1, sgn(best.back() & 1));
            fprintf(fp, "\tbest_path ");
            for (uint32_t x : best) fprintf(fp, "%u:%c ", x >> 1, sgn(x & 1));
            fprintf(fp, "\n");
            for (size_t j = 0; j + 1 < best.size(); ++j) edges.erase({best[j], best[j + 1]});
            fprintf(fp, "\tremoved_edges:\n");
            for (const auto& e : edges) {
                remove_edge(g, e.first >> 1, e.first & 1, e.second >> 1, e.second & 1);
                fprintf(fp, "\t\t%u:%c -> %u:%c\n", e.first >> 1, sgn(e.first & 1), e.second >> 1, sgn(e.second & 1));
            }
            fprintf(fp, "\n");
            ++removed;
            again = true;
        }
        if (again) --i;
    }
    fclose(fp);
    return removed;
}

// triangles a -> i -> b with a direct a -> b edge: drop the less supported way — Cleaning.cpp:7-57
int clean_small_bubbles(Graph& g, const std::string& logpath) {
    FILE* fp = open_write(logpath);
    int removed = 0;
    for (uint32_t i = 0; i < g.size(); ++i) {
        if (g[i].edges[1].empty() || g[i].edges[0].empty()) continue;
        bool done = false;
        for (auto in = g[i].edges[1].begin(); in != g[i].edges[1].end(); ++in) {
            for (auto out = g[i].edges[0].begin(); out != g[i].edges[0].end(); ++out) {
                const uint32_t node1 = in->first >> 1, rev1 = in->first & 1;
                const uint32_t to = out->first, node2 = to >> 1, rev2 = to & 1;
                auto direct = g[node1].edges[1 - rev1].find(to);
                if (direct == g[node1].edges[1 - rev1].end()) continue;
                const double short_cov = (double)direct->second.edge_supp.size();
                const double long_cov = (in->second.edge_supp.size() + out->second.edge_supp.size()) / 2.0;
                fprintf(fp, "small_bubble cov:%.2lf %u:%c -> %u:%c\n", short_cov, node1, sgn(1 - rev1), node2, sgn(rev2));
                fprintf(fp, "             cov:%.2lf %u:%c -> %u:%c -> %u:%c\n", long_cov, node1, sgn(1 - rev1), i, '+', node2, sgn(rev2));
                if (short_cov < long_cov) {
                    remove_edge(g, node1, 1 - rev1, node2, rev2);
                } else {
                    remove_edge(g, node1, 1 - rev1, i, 0);
                    remove_edge(g, i, 0, node2, rev2);
                }
                ++removed;
                done = true;
                break;                     // the iterators may be gone: leave both loops before touching them, as the reference does
            }
            if (done) break;
        }
    }
    fclose(fp);
    return removed;
}

}  // namespace haslr
