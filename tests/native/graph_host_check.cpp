// TEST-ONLY host harness for haslr_b200/csrc/poa_graph.cuh (the POA graph mutators are __host__ __device__).
// Runs one edge's POA on the CPU with the PRODUCT's graph code (add_alignment, topological sort, per-rank DP
// records, consensus) driven by a scalar graph-NW written here against the same per-rank records the CUDA fill
// consumes. Lets `-m "not gpu"` tests compare the graph logic with the oracle without a GPU. Never shipped.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../haslr_b200/csrc/poa_graph.cuh"

using namespace hgpu;

namespace {
struct Store {
    std::vector<uint8_t> code, mark, check;
    std::vector<uint32_t> in_head, in_tail, out_head, aligned, e_begin, e_end, e_w, e_next_in, e_next_out, rank2node, node2rank,
        meta0, pred_off, pred_rank, stack, sinks;
    std::vector<int32_t> aln_rank, aln_pos, pred;
    std::vector<int64_t> score;
    uint32_t n_nodes = 0, n_edges = 0, aln_len = 0, n_sinks = 0;
    GraphView g; GraphScratch s;
    Store(uint32_t ncap, uint32_t ecap) {
        code.resize(ncap); mark.resize(ncap); check.resize(ncap);
        for (auto* v : {&in_head, &in_tail, &out_head, &rank2node, &node2rank, &meta0}) v->resize(ncap);
        aligned.resize(3 * (size_t)ncap); pred_off.resize(ncap + 1);
        for (auto* v : {&e_begin, &e_end, &e_w, &e_next_in, &e_next_out, &pred_rank}) v->resize(ecap);
        stack.resize(ecap + 4 * (size_t)ncap + 8);
        aln_rank.resize(ncap); aln_pos.resize(ncap); pred.resize(ncap); score.resize(ncap);
        g.ncap = ncap; g.ecap = ecap;
        g.code = code.data(); g.in_head = in_head.data(); g.in_tail = in_tail.data(); g.out_head = out_head.data();
        g.aligned = aligned.data(); g.e_begin = e_begin.data(); g.e_end = e_end.data(); g.e_w = e_w.data();
        g.e_next_in = e_next_in.data(); g.e_next_out = e_next_out.data(); g.rank2node = rank2node.data();
        g.node2rank = node2rank.data(); g.meta0 = meta0.data(); g.pred_off = pred_off.data();
        g.pred_rank = pred_rank.data(); g.aln_rank = aln_rank.data(); g.aln_pos = aln_pos.data();
        g.n_nodes = &n_nodes; g.n_edges = &n_edges; g.aln_len = &aln_len; sinks.resize(ncap); g.sinks = sinks.data(); g.n_sinks = &n_sinks;
        s.mark = mark.data(); s.check = check.data(); s.stack = stack.data(); s.stack_cap = (uint32_t)stack.size();
        s.score = score.data(); s.pred = pred.data();
    }
};

// predecessor rows (1-based H rows) of rank r from the packed records, exactly as the CUDA fill decodes them
void pred_rows(const GraphView& g, uint32_t r, std::vector<uint32_t>& out) {
    out.clear();
    const uint32_t m0 = g.meta0[r], m1 = meta_d1(m0), npc = (m0 >> 3) & 3u, d0 = meta_d0(m0), i = r + 1;
    if (npc == 0) { out.push_back(0); return; }
    if (npc == 3) { for (uint32_t x = g.pred_off[r]; x < g.pred_off[r + 1]; ++x) out.push_back(g.pred_rank[x] + 1); return; }
    out.push_back(i - d0);
    if (npc == 2) out.push_back(i - m1);
}

bool align_scalar(GraphView& g, const uint8_t* seq, uint32_t L, int m, int x, int gp) {
    const uint32_t V = *g.n_nodes, W = L + 1;
    std::vector<int32_t> H((size_t)(V + 1) * W);
    for (uint32_t j = 0; j < W; ++j) H[j] = (int32_t)j * gp;
    std::vector<uint32_t> pr;
    for (uint32_t r = 0; r < V; ++r) {
        pred_rows(g, r, pr);
        int32_t* Hr = &H[(size_t)(r + 1) * W];
        const uint32_t c = g.meta0[r] & 3u;
        int32_t b0 = INT32_MIN;
        for (uint32_t p : pr) b0 = std::max(b0, H[(size_t)p * W]);
        Hr[0] = b0 + gp;
        for (uint32_t j = 1; j < W; ++j) {
            int32_t v = INT32_MIN;
            const int sc = (base_code(seq[j - 1]) == c) ? m : x;
            for (uint32_t p : pr) v = std::max(v, std::max(H[(size_t)p * W + j - 1] + sc, H[(size_t)p * W + j] + gp));
            Hr[j] = std::max(v, Hr[j - 1] + gp);
        }
    }
    int32_t best = INT32_MIN; uint32_t bi = 0;
    for (uint32_t r = 0; r < V; ++r) if ((g.meta0[r] & META_SINK) && H[(size_t)(r + 1) * W + L] > best) { best = H[(size_t)(r + 1) * W + L]; bi = r + 1; }
    if (bi == 0) return false;
    uint32_t i = bi, j = L, n = 0;
    while (!(i == 0 && j == 0)) {
        const int32_t val = H[(size_t)i * W + j];
        uint32_t pi = i, pj = j; bool found = false;
        if (i != 0) {
            pred_rows(g, i - 1, pr);
            const uint32_t c = g.meta0[i - 1] & 3u;
            if (j != 0) {
                const int sc = (base_code(seq[j - 1]) == c) ? m : x;
                for (uint32_t p : pr) if (val == H[(size_t)p * W + j - 1] + sc) { pi = p; pj = j - 1; found = true; break; }
            }
            if (!found) for (uint32_t p : pr) if (val == H[(size_t)p * W + j] + gp) { pi = p; pj = j; found = true; break; }
        }
        if (!found && j != 0 && val == H[(size_t)i * W + j - 1] + gp) { pi = i; pj = j - 1; found = true; }
        if (!found || n >= g.ncap) return false;
        g.aln_rank[n] = (pi == i) ? -1 : (int32_t)(i - 1);
        g.aln_pos[n] = (pj == j) ? -1 : (int32_t)(j - 1);
        ++n; i = pi; j = pj;
    }
    *g.aln_len = n;
    return true;
}
}  // namespace

// Returns consensus length (>= 0) or a negative error. Also exports the final graph in rank order.
extern "C" int graphtest_poa(const uint8_t* bases, const uint64_t* seg_off, uint32_t n_segs, int m, int x, int gp,
                             uint8_t* out_cons, uint32_t out_cap, uint32_t* out_n_nodes,
                             uint32_t* rank2node, uint32_t* pred_off, uint32_t* pred_node, uint32_t* pred_weight, uint32_t cap) {
    uint64_t total = seg_off[n_segs] - seg_off[0];
    uint32_t ncap = (uint32_t)total + 64, ecap = ncap + ncap / 4 + 64;
    Store st(ncap, ecap);
    GraphView& g = st.g;
    bool first = true;
    for (uint32_t s = 0; s < n_segs; ++s) {
        const uint8_t* seq = bases + seg_off[s];
        const uint32_t L = (uint32_t)(seg_off[s + 1] - seg_off[s]);
        if (L == 0) continue;
        if (first) {
            *g.aln_len = 0;
            if (!g_add_alignment(g, seq, L)) return -1;
            first = false;
        } else {
            if (!align_scalar(g, seq, L, m, x, gp)) return -2;
            if (!g_add_alignment(g, seq, L)) return -1;
        }
        if (!g_toposort(g, st.s)) return -3;
        g_build_meta(g);
    }
    *out_n_nodes = *g.n_nodes;
    if (first) return 0;
    const uint32_t N = *g.n_nodes;
    if (rank2node && N <= cap) {
        uint32_t pe = 0;
        for (uint32_t r = 0; r < N; ++r) {
            uint32_t v = g.rank2node[r];
            rank2node[r] = v; pred_off[r] = pe;
            for (uint32_t e = g.in_head[v]; e != NIL; e = g.e_next_in[e]) {
                if (pe < 2 * cap) { pred_node[pe] = g.e_begin[e]; pred_weight[pe] = g.e_w[e]; }
                ++pe;
            }
        }
        pred_off[N] = pe;
    }
    std::vector<uint32_t> ids(ncap);
    uint32_t n = g_consensus(g, st.s, ids.data());
    if (n > out_cap) return -4;
    for (uint32_t i = 0; i < n; ++i) out_cons[i] = (uint8_t)"ACGT"[g.code[ids[i]]];
    return (int)n;
}
