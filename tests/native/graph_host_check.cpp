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

// The parallelisable form of the topological sort described in DESIGN.md section 7 (not shipped: checked here against the
// serial walk on every graph the CPU tests build). (1) claim[u] = smallest node id that reaches u over in-edges and aligned
// links: a min-propagation, swept until nothing changes (every sweep is independent per node: one lane per node on a GPU).
// (2) Every node that claims itself is a root of SPOA's outer loop; its walk only ever enters nodes it claims (whatever an
// earlier root reaches is emitted before it starts), so the walks touch disjoint node sets and could run one per lane.
// (3) The emission lists are concatenated in root id order. Returns false if the order differs from g.rank2node.
uint32_t g_claim_sweeps_max = 0, g_claim_roots = 0, g_claim_nodes = 0;
bool toposort_by_claims_matches(const GraphView& g) {
    const uint32_t N = *g.n_nodes;
    std::vector<uint32_t> claim(N);
    for (uint32_t v = 0; v < N; ++v) claim[v] = v;
    uint32_t sweeps = 0;
    for (bool changed = true; changed; ++sweeps) {
        changed = false;
        std::vector<uint32_t> next(claim);                                  // Jacobi sweep: what a parallel pass over all nodes computes
        for (uint32_t v = 0; v < N; ++v) {
            const uint32_t c = claim[v];
            for (uint32_t x = g.in_head[v]; x != NIL; x = g.e_next_in[x]) if (next[g.e_begin[x]] > c) { next[g.e_begin[x]] = c; changed = true; }
            for (int q = 0; q < 3; ++q) { const uint32_t o = g.aligned[3 * v + q]; if (o == NIL) break; if (next[o] > c) { next[o] = c; changed = true; } }
        }
        claim.swap(next);
    }
    g_claim_sweeps_max = std::max(g_claim_sweeps_max, sweeps);
    std::vector<uint8_t> mark(N, 0), check(N, 1);
    std::vector<uint32_t> order, stack;
    order.reserve(N);
    for (uint32_t i = 0; i < N; ++i) {
        if (claim[i] != i) continue;                                        // reached by a smaller id: not a root
        ++g_claim_roots;
        // the walk of root i, looking only at its own nodes: anything claimed by a smaller root counts as emitted
        auto done = [&](uint32_t u) { return claim[u] < i || mark[u] == 2; };
        stack.assign(1, i);
        while (!stack.empty()) {
            const uint32_t v = stack.back();
            bool valid = true;
            if (mark[v] != 2) {
                if (claim[v] != i) return false;                            // a walk left its own node set: the claim rule is wrong
                for (uint32_t x = g.in_head[v]; x != NIL; x = g.e_next_in[x]) if (!done(g.e_begin[x])) { stack.push_back(g.e_begin[x]); valid = false; }
                if (check[v])
                    for (int q = 0; q < 3; ++q) {
                        const uint32_t o = g.aligned[3 * v + q];
                        if (o == NIL) break;
                        if (!done(o)) { stack.push_back(o); check[o] = 0; valid = false; }
                    }
                if (!valid && mark[v] == 1) return false;
                if (valid) {
                    mark[v] = 2;
                    if (check[v]) {
                        order.push_back(v);
                        for (int q = 0; q < 3; ++q) { const uint32_t o = g.aligned[3 * v + q]; if (o == NIL) break; order.push_back(o); }
                    }
                } else mark[v] = 1;
            }
            if (valid) stack.pop_back();
        }
    }
    g_claim_nodes += N;
    if (order.size() != N) return false;
    for (uint32_t r = 0; r < N; ++r) if (order[r] != g.rank2node[r]) return false;
    return true;
}

// The same sort laid out the way the CUDA function does it (haslr_b200/csrc/poa_topo_claims.cuh: w_toposort_claims): the
// per-node records of w_build_trec (four in-edge tails, three aligned nodes, the fifth in-edge), the "only a dependency with a
// larger id can be lowered" rule of the sweeps, roots that claim only themselves ranked by the scan alone, every walk on a
// stack region of 5 words per claimed node at 5 x its first rank, the step guard. Lanes are run one after the other here.
// Returns 1 ranks equal, 0 the function would have declined (the caller falls back), -1 ranks differ.
uint32_t g_dev_declined = 0, g_dev_stack_peak_x100 = 0;
int toposort_claims_device_layout(const GraphView& g) {
    const uint32_t N = *g.n_nodes;
    struct Rec { uint32_t p[4], a[3], more; };
    std::vector<Rec> rec(N);
    for (uint32_t v = 0; v < N; ++v) {
        Rec r; for (auto& x : r.p) x = NIL; for (auto& x : r.a) x = NIL;
        uint32_t x = g.in_head[v];
        for (int q = 0; q < 4 && x != NIL; ++q) { r.p[q] = g.e_begin[x]; x = g.e_next_in[x]; }
        r.more = x;
        for (int q = 0; q < 3; ++q) { if (g.aligned[3 * v + q] == NIL) break; r.a[q] = g.aligned[3 * v + q]; }
        rec[v] = r;
    }
    std::vector<uint32_t> claim(N), cnt(N, 0), base(N, 0), roots, r2n(N, NIL), stk(5 * (size_t)N + 8);
    std::vector<uint8_t> mark(N, 0), check(N, 1);
    for (uint32_t v = 0; v < N; ++v) claim[v] = v;
    bool settled = false;
    for (uint32_t sweep = 0; sweep < 256 && !settled; ++sweep) {
        bool changed = false;
        for (uint32_t v = 0; v < N; ++v) {
            const uint32_t c = claim[v];
            const uint32_t d[7] = {rec[v].p[0], rec[v].p[1], rec[v].p[2], rec[v].p[3], rec[v].a[0], rec[v].a[1], rec[v].a[2]};
            for (int q = 0; q < 7; ++q) if (d[q] != NIL && d[q] > c && claim[d[q]] > c) { claim[d[q]] = c; changed = true; }
            for (uint32_t x = rec[v].more; x != NIL; x = g.e_next_in[x]) { const uint32_t b = g.e_begin[x]; if (b > c && claim[b] > c) { claim[b] = c; changed = true; } }
        }
        settled = !changed;
    }
    if (!settled) { ++g_dev_declined; return 0; }
    for (uint32_t u = 0; u < N; ++u) ++cnt[claim[u]];
    uint32_t running = 0;
    for (uint32_t i = 0; i < N; ++i) {
        const uint32_t c = cnt[i];
        if (c != 0) base[i] = running;
        if (c == 1) r2n[running] = i;
        if (c > 1) roots.push_back(i);
        running += c;
    }
    if (running != N) { ++g_dev_declined; return 0; }
    for (uint32_t i : roots) {
        const uint32_t b = base[i], n_own = cnt[i], cap = 5 * n_own, limit = 64 * n_own + 64;
        uint32_t* st = stk.data() + 5 * (size_t)b;
        uint32_t sp = 1, k = 0, guard = 0, peak = 1;
        st[0] = i;
        bool ok = true;
        while (sp > 0 && ok) {
            if (++guard > limit) { ok = false; break; }
            const uint32_t v = st[sp - 1];
            if (mark[v] == 2) { --sp; continue; }
            const Rec& r = rec[v];
            const uint32_t d[7] = {r.p[0], r.p[1], r.p[2], r.p[3], r.a[0], r.a[1], r.a[2]};
            const bool chk = check[v] != 0;
            bool valid = true;
            auto dep = [&](uint32_t u, bool aligned_link) {
                const uint32_t cu = claim[u];
                if (cu < i || mark[u] == 2) return;
                if (cu != i || sp >= cap) { ok = false; return; }
                st[sp++] = u; valid = false; peak = std::max(peak, sp);
                if (aligned_link) check[u] = 0;
            };
            for (int q = 0; q < 4; ++q) if (d[q] != NIL) dep(d[q], false);
            for (uint32_t x = r.more; x != NIL && ok; x = g.e_next_in[x]) dep(g.e_begin[x], false);
            if (chk) for (int q = 4; q < 7; ++q) if (d[q] != NIL) dep(d[q], true);
            if (!ok) break;
            if (!valid) { if (mark[v] == 1) { ok = false; break; } mark[v] = 1; continue; }
            mark[v] = 2;
            if (chk) {
                const uint32_t na = (d[4] != NIL) + (d[5] != NIL) + (d[6] != NIL);
                if (k + 1 + na > n_own) { ok = false; break; }
                r2n[b + k++] = v;
                for (int q = 4; q < 7; ++q) if (d[q] != NIL) r2n[b + k++] = d[q];
            }
            --sp;
        }
        g_dev_stack_peak_x100 = std::max(g_dev_stack_peak_x100, 100 * peak / n_own);
        if (!ok || k != n_own) { ++g_dev_declined; return 0; }
    }
    for (uint32_t r = 0; r < N; ++r) if (r2n[r] != g.rank2node[r]) return -1;
    return 1;
}
}  // namespace

extern "C" void graphtest_claim_device_stats(uint32_t* declined, uint32_t* stack_peak_x100) { *declined = g_dev_declined; *stack_peak_x100 = g_dev_stack_peak_x100; }

extern "C" void graphtest_claim_stats(uint32_t* sweeps_max, uint32_t* roots, uint32_t* nodes) {
    *sweeps_max = g_claim_sweeps_max; *roots = g_claim_roots; *nodes = g_claim_nodes;
}

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
        if (!toposort_by_claims_matches(g)) return -5;
        if (toposort_claims_device_layout(g) < 0) return -6;
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
