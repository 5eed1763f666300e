// POA graph state and its serial mutators, shared by device kernels and (for unit tests) host code.
//
// Replaces the SPOA 1.1.3 Graph object that the reference builds per backbone edge
// (reference src/haslr_assemble/src/Assemble.cpp:500,540,554): add_alignment, the DFS topological
// sort that keeps aligned nodes adjacent, and the heaviest-bundle consensus with branch completion.
// The graph of one backbone edge lives in flat structure-of-arrays storage in HBM (see GraphView);
// adjacency is kept as intrusive singly linked lists so that a read can append in O(1) without
// reallocation, and in-edge order == insertion order, which is what fixes DP predecessor order and
// every tie-break downstream.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HGPU_HD __host__ __device__ __forceinline__
#else
#define HGPU_HD inline
#endif
#define HGPU_HD_FWD HGPU_HD

namespace hgpu {

static constexpr uint32_t NIL = 0xFFFFFFFFu;

// meta0 layout (one word per rank, consumed by the DP/traceback kernel)
//   bits 0-1  node code (A,C,G,T = 0..3)
//   bit  2    sink (no out-edges)
//   bits 3-4  predecessor class: 0 none, 1 one, 2 two, 3 walk the CSR (three or more, or a distance >= 8192)
//   bit  5    fast row: its only predecessor is the previous rank (or it is rank 0 without predecessors)
//   bits 6-18  rank distance to the first predecessor  (classes 1, 2)
//   bits 19-31 rank distance to the second predecessor (class 2)
static constexpr uint32_t META_SINK = 4u;
static constexpr uint32_t META_FAST = 32u;
static constexpr int META_D0_SHIFT = 6, META_D1_SHIFT = 19;
static constexpr uint32_t META_DMAX = 1u << 13;
HGPU_HD_FWD uint32_t meta_d0(uint32_t m0) { return (m0 >> META_D0_SHIFT) & (META_DMAX - 1u); }
HGPU_HD_FWD uint32_t meta_d1(uint32_t m0) { return m0 >> META_D1_SHIFT; }
// record of a rank with code/sink bits `base`, np predecessors, the first two at distances d0, d1
HGPU_HD_FWD uint32_t meta_pack(uint32_t base, uint32_t rank, uint32_t np, uint32_t d0, uint32_t d1) {
    if (np == 0) return base | (rank == 0 ? META_FAST : 0u);
    if (np > 2 || d0 >= META_DMAX || (np == 2 && d1 >= META_DMAX)) return base | (3u << 3);
    if (np == 1) return base | (1u << 3) | (d0 << META_D0_SHIFT) | (d0 == 1 ? META_FAST : 0u);
    return base | (2u << 3) | (d0 << META_D0_SHIFT) | (d1 << META_D1_SHIFT);
}

struct GraphView {
    uint32_t ncap;        // capacity in nodes (== capacity of the aln arrays)
    uint32_t ecap;        // capacity in edges
    uint8_t* code;        // [ncap] node base code
    uint32_t* in_head;    // [ncap] first in-edge (edge id) or NIL
    uint32_t* in_tail;    // [ncap] last in-edge
    uint32_t* out_head;   // [ncap] first out-edge
    uint32_t* aligned;    // [3*ncap] aligned node ids, NIL-terminated, insertion order
    uint32_t* e_begin;    // [ecap] edge pool
    uint32_t* e_end;
    uint32_t* e_w;
    uint32_t* e_next_in;  // next edge in the end node's in-list
    uint32_t* e_next_out; // next edge in the begin node's out-list
    uint32_t* rank2node;  // [ncap]
    uint32_t* node2rank;  // [ncap]
    uint32_t* meta0;      // [ncap] by rank
    uint32_t* pred_off;   // [ncap+1] by rank, CSR into pred_rank
    uint32_t* pred_rank;  // [ecap]
    uint32_t* sinks;      // [ncap] ranks of the nodes without out-edges, ascending (the traceback's end-cell candidates)
    uint32_t* n_sinks;
    int32_t* aln_rank;    // [ncap] alignment, traceback order (last pair first): rank or -1
    int32_t* aln_pos;     // [ncap] sequence position or -1
    uint32_t* n_nodes;    // scalars of this edge
    uint32_t* n_edges;
    uint32_t* aln_len;
};

// Scratch owned by whichever worker currently updates a graph.
struct GraphScratch {
    uint8_t* mark;     // [ncap] 0 unmarked, 1 temporary, 2 permanent
    uint8_t* check;    // [ncap] check_aligned_nodes
    uint32_t* stack;   // [stack_cap]
    uint32_t stack_cap;
    int64_t* score;    // [ncap] consensus
    int32_t* pred;     // [ncap] consensus
};

HGPU_HD uint32_t base_code(uint8_t c) {
    // Reference Compressed_sequence.cpp:10-19,57: case-folded; anything but A/C/G/T packs to code 0 ('A').
    c &= 0xDF;
    return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}

HGPU_HD uint32_t g_add_node(GraphView& g, uint32_t code) {
    uint32_t id = *g.n_nodes;
    g.code[id] = (uint8_t)code;
    g.in_head[id] = NIL; g.in_tail[id] = NIL; g.out_head[id] = NIL;
    g.aligned[3 * id] = NIL; g.aligned[3 * id + 1] = NIL; g.aligned[3 * id + 2] = NIL;
    *g.n_nodes = id + 1;
    return id;
}

HGPU_HD void g_add_edge(GraphView& g, uint32_t b, uint32_t e, uint32_t w) {
    for (uint32_t x = g.out_head[b]; x != NIL; x = g.e_next_out[x]) {
        if (g.e_end[x] == e) { g.e_w[x] += w; return; }
    }
    uint32_t x = *g.n_edges;
    g.e_begin[x] = b; g.e_end[x] = e; g.e_w[x] = w;
    g.e_next_in[x] = NIL;
    g.e_next_out[x] = g.out_head[b];
    g.out_head[b] = x;
    if (g.in_tail[e] == NIL) g.in_head[e] = x; else g.e_next_in[g.in_tail[e]] = x;
    g.in_tail[e] = x;
    *g.n_edges = x + 1;
}

// Chain of seq[begin..end) with unit base weights (edge weight 2). Returns first node id or NIL.
HGPU_HD uint32_t g_add_sequence(GraphView& g, const uint8_t* seq, uint32_t begin, uint32_t end) {
    if (begin == end) return NIL;
    uint32_t first = g_add_node(g, base_code(seq[begin]));
    for (uint32_t i = begin + 1; i < end; ++i) {
        uint32_t id = g_add_node(g, base_code(seq[i]));
        g_add_edge(g, id - 1, id, 2);
    }
    return first;
}

// SPOA Graph::add_alignment with unit weights. The alignment is read from g.aln_rank/g.aln_pos, which hold it
// in traceback order (index aln_len-1 is the first pair) with graph positions given as ranks of the topological
// order the alignment was computed on; they are translated through rank2node before anything is modified.
// Returns false when the node/edge capacity would be exceeded.
HGPU_HD bool g_add_alignment(GraphView& g, const uint8_t* seq, uint32_t L) {
    if (L == 0) return true;
    const uint32_t n = *g.aln_len;
    if (*g.n_nodes + L > g.ncap || *g.n_edges + L + 1 > g.ecap) return false;  // conservative: every base may create a node and an edge
    if (n == 0) {
        g_add_sequence(g, seq, 0, L);
        return true;
    }
    // first / last aligned sequence position
    int32_t front = -1, back = -1;
    for (uint32_t t = n; t-- > 0;) if (g.aln_pos[t] != -1) { front = g.aln_pos[t]; break; }
    for (uint32_t t = 0; t < n; ++t) if (g.aln_pos[t] != -1) { back = g.aln_pos[t]; break; }
    uint32_t before = *g.n_nodes;
    g_add_sequence(g, seq, 0, (uint32_t)front);
    uint32_t head = (*g.n_nodes == before) ? NIL : *g.n_nodes - 1;
    uint32_t tail = g_add_sequence(g, seq, (uint32_t)back + 1, L);
    for (uint32_t t = n; t-- > 0;) {
        int32_t pos = g.aln_pos[t];
        if (pos == -1) continue;
        uint32_t c = base_code(seq[pos]);
        int32_t rk = g.aln_rank[t];
        uint32_t nid;
        if (rk == -1) {
            nid = g_add_node(g, c);
        } else {
            uint32_t a = g.rank2node[rk];
            if (g.code[a] == c) {
                nid = a;
            } else {
                uint32_t hit = NIL;
                for (int q = 0; q < 3; ++q) {
                    uint32_t o = g.aligned[3 * a + q];
                    if (o == NIL) break;
                    if (g.code[o] == c) { hit = o; break; }
                }
                if (hit == NIL) {
                    nid = g_add_node(g, c);
                    int cnt = 0;
                    for (int q = 0; q < 3; ++q) {
                        uint32_t o = g.aligned[3 * a + q];
                        if (o == NIL) break;
                        g.aligned[3 * nid + cnt++] = o;
                        for (int z = 0; z < 3; ++z) if (g.aligned[3 * o + z] == NIL) { g.aligned[3 * o + z] = nid; break; }
                    }
                    g.aligned[3 * nid + cnt] = a;
                    for (int z = 0; z < 3; ++z) if (g.aligned[3 * a + z] == NIL) { g.aligned[3 * a + z] = nid; break; }
                } else {
                    nid = hit;
                }
            }
        }
        if (head != NIL) g_add_edge(g, head, nid, 2);
        head = nid;
    }
    if (tail != NIL) g_add_edge(g, head, tail, 2);
    return true;
}

// SPOA Graph::topological_sort: iterative DFS over in-edges in node-id order; the aligned nodes of a node are
// visited with it and emitted right after it. Returns false on stack overflow or a cycle.
HGPU_HD bool g_toposort(GraphView& g, GraphScratch& s) {
    const uint32_t N = *g.n_nodes;
    for (uint32_t i = 0; i < N; ++i) { s.mark[i] = 0; s.check[i] = 1; }
    uint32_t nr = 0, sp = 0;
    for (uint32_t i = 0; i < N; ++i) {
        if (s.mark[i] != 0) continue;
        s.stack[sp++] = i;
        while (sp > 0) {
            uint32_t v = s.stack[sp - 1];
            bool valid = true;
            if (s.mark[v] != 2) {
                for (uint32_t x = g.in_head[v]; x != NIL; x = g.e_next_in[x]) {
                    uint32_t b = g.e_begin[x];
                    if (s.mark[b] != 2) {
                        if (sp >= s.stack_cap) return false;
                        s.stack[sp++] = b; valid = false;
                    }
                }
                if (s.check[v]) {
                    for (int q = 0; q < 3; ++q) {
                        uint32_t o = g.aligned[3 * v + q];
                        if (o == NIL) break;
                        if (s.mark[o] != 2) {
                            if (sp >= s.stack_cap) return false;
                            s.stack[sp++] = o; s.check[o] = 0; valid = false;
                        }
                    }
                }
                if (!valid && s.mark[v] == 1) return false;  // not a DAG
                if (valid) {
                    s.mark[v] = 2;
                    if (s.check[v]) {
                        g.rank2node[nr] = v; g.node2rank[v] = nr; ++nr;
                        for (int q = 0; q < 3; ++q) {
                            uint32_t o = g.aligned[3 * v + q];
                            if (o == NIL) break;
                            g.rank2node[nr] = o; g.node2rank[o] = nr; ++nr;
                        }
                    }
                } else {
                    s.mark[v] = 1;
                }
            }
            if (valid) --sp;
        }
    }
    return nr == N;
}

// Per-rank records for the DP kernel: code, sink flag, predecessor ranks in in-edge order.
HGPU_HD void g_build_meta(GraphView& g) {
    const uint32_t N = *g.n_nodes;
    uint32_t pe = 0, ns = 0;
    for (uint32_t r = 0; r < N; ++r) {
        uint32_t v = g.rank2node[r];
        const uint32_t base = g.code[v] | (g.out_head[v] == NIL ? META_SINK : 0u);
        g.pred_off[r] = pe;
        uint32_t np = 0, d0 = 0, d1 = 0;
        for (uint32_t x = g.in_head[v]; x != NIL; x = g.e_next_in[x]) {
            uint32_t pr = g.node2rank[g.e_begin[x]];
            g.pred_rank[pe++] = pr;
            if (np == 0) d0 = r - pr;
            if (np == 1) d1 = r - pr;
            ++np;
        }
        g.meta0[r] = meta_pack(base, r, np, d0, d1);
        if (base & META_SINK) g.sinks[ns++] = r;
    }
    g.pred_off[N] = pe;
    *g.n_sinks = ns;
}

// SPOA Graph::traverse_heaviest_bundle + branch_completion. Writes node ids of the consensus path into `out`
// (capacity ncap) and returns its length.
HGPU_HD uint32_t g_branch_completion(GraphView& g, GraphScratch& s, uint32_t rank) {
    const uint32_t N = *g.n_nodes;
    uint32_t v = g.rank2node[rank];
    for (uint32_t x = g.out_head[v]; x != NIL; x = g.e_next_out[x]) {
        uint32_t w = g.e_end[x];
        for (uint32_t y = g.in_head[w]; y != NIL; y = g.e_next_in[y])
            if (g.e_begin[y] != v) s.score[g.e_begin[y]] = -1;
    }
    int64_t max_score = 0;
    uint32_t max_id = 0;
    for (uint32_t i = rank + 1; i < N; ++i) {
        uint32_t u = g.rank2node[i];
        s.score[u] = -1; s.pred[u] = -1;
        for (uint32_t y = g.in_head[u]; y != NIL; y = g.e_next_in[y]) {
            uint32_t b = g.e_begin[y];
            if (s.score[b] == -1) continue;
            int64_t w = g.e_w[y];
            if (s.score[u] < w || (s.score[u] == w && s.score[s.pred[u]] <= s.score[b])) { s.score[u] = w; s.pred[u] = (int32_t)b; }
        }
        if (s.pred[u] != -1) s.score[u] += s.score[s.pred[u]];
        if (max_score < s.score[u]) { max_score = s.score[u]; max_id = u; }
    }
    return max_id;
}

// first half of SPOA's traverse_heaviest_bundle: best in-edge / path score of every node in rank order; returns the
// first node (rank order) with the maximal score
HGPU_HD uint32_t g_consensus_scores(GraphView& g, GraphScratch& s) {
    const uint32_t N = *g.n_nodes;
    for (uint32_t i = 0; i < N; ++i) { s.score[i] = -1; s.pred[i] = -1; }
    uint32_t max_id = 0;
    for (uint32_t r = 0; r < N; ++r) {
        uint32_t u = g.rank2node[r];
        for (uint32_t y = g.in_head[u]; y != NIL; y = g.e_next_in[y]) {
            uint32_t b = g.e_begin[y];
            int64_t w = g.e_w[y];
            if (s.score[u] < w || (s.score[u] == w && s.score[s.pred[u]] <= s.score[b])) { s.score[u] = w; s.pred[u] = (int32_t)b; }
        }
        if (s.pred[u] != -1) s.score[u] += s.score[s.pred[u]];
        if (s.score[max_id] < s.score[u]) max_id = u;
    }
    return max_id;
}

// second half: branch completion until the path ends in a sink, then the backtrack; node ids of the path land in `out`
HGPU_HD uint32_t g_consensus_finish(GraphView& g, GraphScratch& s, uint32_t max_id, uint32_t* out) {
    while (g.out_head[max_id] != NIL) max_id = g_branch_completion(g, s, g.node2rank[max_id]);
    uint32_t n = 0;
    while (s.pred[max_id] != -1) { out[n++] = max_id; max_id = (uint32_t)s.pred[max_id]; }
    out[n++] = max_id;
    for (uint32_t a = 0, b = n - 1; a < b; ++a, --b) { uint32_t t = out[a]; out[a] = out[b]; out[b] = t; }
    return n;
}

HGPU_HD uint32_t g_consensus(GraphView& g, GraphScratch& s, uint32_t* out) {
    if (*g.n_nodes == 0) return 0;
    return g_consensus_finish(g, s, g_consensus_scores(g, s), out);
}

}  // namespace hgpu
