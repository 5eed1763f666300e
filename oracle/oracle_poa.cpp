// ORACLE — TEST INFRASTRUCTURE ONLY. See oracle.h.
//
// Per-edge POA consensus exactly as asm_calc_single_cns_seq drives SPOA
// (reference src/haslr_assemble/src/Assemble.cpp:499-554): one engine + one graph per edge, segments in the
// given order, empty segments skipped, generate_consensus at the end; worker threads grab one edge at a time
// from a mutex-guarded cursor like asm_get_next_edge (Assemble.cpp:386-434).
// PARITY UNPINNED: the SPOA semantics come from spoa_restated/spoa.hpp.
#include "oracle.h"
#include "spoa_restated/spoa.hpp"

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
inline char norm_base(uint8_t c) {
    // Compressed_sequence.cpp:10-19,57: non-ACGT packs to code 0 = 'A'; case is folded by the 2-bit table
    switch (c) {
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 'A';
    }
}
std::string seg_string(const uint8_t* bases, uint64_t b, uint64_t e) {
    std::string s(e - b, 'A');
    for (uint64_t i = b; i < e; ++i) s[i - b] = norm_base(bases[i]);
    return s;
}
}  // namespace

extern "C" int oracle_poa_batch(const uint8_t* bases, const uint64_t* seg_off, const uint32_t* edge_seg_off, uint32_t n_edges,
                                int match, int mismatch, int gap, int simd, int threads,
                                uint8_t* out_cons, uint64_t out_cap, uint64_t* out_cons_off,
                                uint64_t* out_cells, uint32_t* out_nodes) {
    std::vector<std::string> cons(n_edges);
    std::atomic<uint64_t> cells(0);
    std::mutex lock;
    uint32_t cursor = 0;
    auto worker = [&]() {
        while (true) {
            uint32_t e;
            {
                std::lock_guard<std::mutex> g(lock);
                if (cursor >= n_edges) return;
                e = cursor++;
            }
            auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, (int8_t)match, (int8_t)mismatch, (int8_t)gap);
            engine->set_simd(simd != 0);
            auto graph = spoa::createGraph();
            uint32_t cnt_non_empty = 0;
            for (uint32_t s = edge_seg_off[e]; s < edge_seg_off[e + 1]; ++s) {
                std::string sub = seg_string(bases, seg_off[s], seg_off[s + 1]);
                if (sub.size() > 0) {
                    auto alignment = engine->align_sequence_with_graph(sub, graph);
                    graph->add_alignment(alignment, sub);
                    cnt_non_empty++;
                }
            }
            if (cnt_non_empty > 0) cons[e] = graph->generate_consensus();
            if (out_nodes) out_nodes[e] = (uint32_t)graph->nodes().size();
            cells += engine->cells();
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    uint64_t off = 0;
    for (uint32_t e = 0; e < n_edges; ++e) {
        out_cons_off[e] = off;
        if (off + cons[e].size() > out_cap) return -1;
        memcpy(out_cons + off, cons[e].data(), cons[e].size());
        off += cons[e].size();
    }
    out_cons_off[n_edges] = off;
    if (out_cells) *out_cells = cells.load();
    return 0;
}

extern "C" int oracle_poa_debug(const uint8_t* bases, const uint64_t* seg_off, uint32_t n_segs, uint32_t n_prior,
                                int match, int mismatch, int gap,
                                int32_t* H, uint64_t H_cap,
                                int32_t* aln_node, int32_t* aln_pos, uint32_t aln_cap,
                                uint32_t* rank2node, uint8_t* node_code, uint32_t* pred_off, uint32_t* pred_node, uint32_t* pred_weight,
                                uint32_t node_cap, uint32_t edge_cap, oracle_poa_dbg_sizes* sizes) {
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, (int8_t)match, (int8_t)mismatch, (int8_t)gap);
    engine->set_simd(false);
    auto graph = spoa::createGraph();
    uint32_t done = 0, s = 0;
    for (; s < n_segs && done < n_prior; ++s) {
        std::string sub = seg_string(bases, seg_off[s], seg_off[s + 1]);
        if (sub.empty()) continue;
        auto alignment = engine->align_sequence_with_graph(sub, graph);
        graph->add_alignment(alignment, sub);
        ++done;
    }
    const auto& nodes = graph->nodes();
    const auto& edges = graph->edges();
    const auto& r2n = graph->rank_to_node_id();
    if (nodes.size() > node_cap || edges.size() > edge_cap) return -1;
    sizes->n_nodes = (uint32_t)nodes.size();
    sizes->n_edges = (uint32_t)edges.size();
    sizes->aln_len = 0;
    sizes->L = 0;
    // graph in RANK order: rank2node, code by node id, predecessor CSR by rank (in-edge order), as node ids
    uint32_t pe = 0;
    for (uint32_t r = 0; r < nodes.size(); ++r) {
        uint32_t nid = r2n[r];
        if (rank2node) rank2node[r] = nid;
        if (pred_off) pred_off[r] = pe;
        for (uint32_t ei : nodes[nid].in_edges) {
            if (pred_node) pred_node[pe] = edges[ei].begin_node_id;
            if (pred_weight) pred_weight[pe] = (uint32_t)edges[ei].total_weight;
            ++pe;
        }
    }
    if (pred_off) pred_off[nodes.size()] = pe;
    if (node_code) for (uint32_t i = 0; i < nodes.size(); ++i) node_code[i] = (uint8_t)graph->decoder(nodes[i].code);
    // next non-empty segment is the query
    while (s < n_segs && seg_off[s + 1] == seg_off[s]) ++s;
    if (s >= n_segs) return (int)nodes.size();
    std::string sub = seg_string(bases, seg_off[s], seg_off[s + 1]);
    sizes->L = (uint32_t)sub.size();
    // scalar H fill restated here so the matrix can be exported (same recurrences as spoa.hpp align_int32)
    if (H) {
        const uint32_t W = (uint32_t)sub.size() + 1;
        if ((uint64_t)(nodes.size() + 1) * W > H_cap) return -1;
        std::vector<uint32_t> n2r(nodes.size());
        for (uint32_t r = 0; r < nodes.size(); ++r) n2r[r2n[r]] = r;
        for (uint32_t j = 0; j < W; ++j) H[j] = (int32_t)j * gap;
        for (uint32_t r = 0; r < nodes.size(); ++r) {
            const auto& node = nodes[r2n[r]];
            int32_t* Hr = H + (uint64_t)(r + 1) * W;
            char nb = graph->decoder(node.code);
            std::vector<uint32_t> preds;
            if (node.in_edges.empty()) preds.push_back(0);
            for (uint32_t ei : node.in_edges) preds.push_back(n2r[edges[ei].begin_node_id] + 1);
            int32_t best0 = INT32_MIN;
            for (uint32_t p : preds) best0 = std::max(best0, H[(uint64_t)p * W]);
            Hr[0] = best0 + gap;
            for (uint32_t j = 1; j < W; ++j) {
                int32_t v = INT32_MIN;
                for (uint32_t p : preds) {
                    const int32_t* Hp = H + (uint64_t)p * W;
                    v = std::max(v, std::max(Hp[j - 1] + (nb == sub[j - 1] ? match : mismatch), Hp[j] + gap));
                }
                Hr[j] = std::max(v, Hr[j - 1] + gap);
            }
        }
    }
    auto alignment = engine->align_sequence_with_graph(sub, graph);
    sizes->aln_len = (uint32_t)alignment.size();
    if (aln_node && aln_pos) {
        if (alignment.size() > aln_cap) return -1;
        for (size_t i = 0; i < alignment.size(); ++i) { aln_node[i] = alignment[i].first; aln_pos[i] = alignment[i].second; }
    }
    return (int)nodes.size();
}
