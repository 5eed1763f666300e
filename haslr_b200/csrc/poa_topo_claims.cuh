// w_toposort_claims: SPOA Graph::topological_sort (the order the reference's consensus depends on, reached through
// spoa::Graph::add_alignment, Assemble.cpp:540), exact, with the walks of its outer loop running side by side.
//
// Included from poa_device.cuh. Why: the sort was the largest serial piece of a deep edge's chain of alignments (14 % of
// k_poa_pool's warp-cycles, a third of the time between two fills of a 9,000-node graph): one DFS after the other, every
// visit a dependent round trip to L2. The order itself allows more:
//   (1) SPOA starts a walk from every node id in ascending order that is still unmarked. A walk marks everything it reaches
//       over in-edges and aligned links, so node u ends up in the walk of the SMALLEST id that reaches it:
//       claim[u] = min(u, claim[v] for every v that lists u as in-edge tail or aligned node). That is a min-propagation over
//       the DAG, one lane per node, swept until a sweep changes nothing (branches off the backbone are short: 3-4 sweeps).
//   (2) The nodes with claim[i] == i are exactly the roots of SPOA's outer loop, and a root's walk only ever enters nodes it
//       claims (whatever a smaller root reaches is emitted before it starts; a node pushed as somebody's aligned sibling
//       reaches that somebody back, so both have the same claim). The walks therefore touch disjoint node sets: one walk per
//       lane, 32 at a time, each with SPOA's own visiting order, marks and "do not check aligned" flags.
//   (3) A root emits exactly the nodes it claims, so the first rank of every root is an exclusive scan of the claim counts
//       in id order; a root that claims only itself (most of a POA graph) is ranked by that scan alone.
// tests/native/graph_host_check.cpp holds the same three steps on the CPU and compares them with the serial walk on every
// graph the CPU tests build; on the GPU the parity tests compare consensus, graph and ranks with the oracle bit for bit.
// Measured (profiles/r2N_toposort_claims_ab.log): bit-exact in all 46 POA parity tests on the GPU, and NOT faster - 592 deep edges
// 200.4 ms against 200.7-206 ms, config 2 K3 337-338 against 333-338 ms; in the shallow kernel (six reads, ~1,800 nodes) 10 %
// slower (1,536 against 1,706 GCUPS): 32 diverged walks still wait for their own dependent loads one after the other, and the
// sweeps add four passes of atomics over a graph in which one node in sixteen needs a walk at all. It is therefore compiled
// out by default (HGPU_TOPO_CLAIMS=0: the binary is instruction for instruction the one without this file) and kept as the
// starting point for a version that gives each walk to a lane GROUP and batches the claim sweeps with the record pass. It has
// no node limit, unlike w_toposort's shared-memory bitmaps (18,432 nodes, then one lane).
// Returns 1 = ranks written, 0 = not done (tiny graph, scratch too small, a walk outgrew its stack): the caller runs w_toposort.
#pragma once

namespace hgpu {

#ifndef HGPU_TOPO_CLAIMS
#define HGPU_TOPO_CLAIMS 0            // 1: the deep kernels sort with w_toposort_claims (exact, measured no faster: see the header comment)
#endif
#ifndef HGPU_TOPO_CLAIMS_SHALLOW
#define HGPU_TOPO_CLAIMS_SHALLOW 0    // 1: the shallow kernel (k_poa_edges, ~1,800-node graphs of six reads) too
#endif
static constexpr uint32_t TOPO_CLAIMS_MIN_NODES = 96;     // below this the batched walk of w_toposort is a handful of passes
static constexpr uint32_t TOPO_WALK_STACK = 5;            // stack words per claimed node of a walk (a node can be pushed by several parents)

__device__ __noinline__ int w_toposort_claims(GraphView& g, GraphScratch& s, const TopoRec* rec, int lane) {
    const uint32_t N = *g.n_nodes;
    if (N < TOPO_CLAIMS_MIN_NODES || (uint64_t)TOPO_WALK_STACK * N > s.stack_cap || N > g.ncap) return 0;
    // scratch that is dead between add_alignment and w_build_meta: the consensus arrays, the per-rank records
    uint32_t* const r2n = g.rank2node; uint32_t* const n2r = g.node2rank;
    const uint32_t* const e_begin = g.e_begin; const uint32_t* const e_next_in = g.e_next_in;
    uint32_t* const claim = reinterpret_cast<uint32_t*>(s.score);     // [N]   (score is int64[ncap])
    uint32_t* const cnt = claim + N;                                  // [N]   nodes claimed by root i
    uint32_t* const base = g.pred_off;                                // [N]   first rank of root i
    uint32_t* const roots = g.meta0;                                  // roots that claim more than themselves, ascending
    uint8_t* const mark = s.mark; uint8_t* const check = s.check;
    uint32_t* const stk_all = s.stack;
    const uint4* const rec4 = reinterpret_cast<const uint4*>(rec);

    for (uint32_t v = lane; v < N; v += 32) { claim[v] = v; cnt[v] = 0u; mark[v] = 0; check[v] = 1; }
    __syncwarp();

    // ---- (1) claims. Atomics work in L2: claim[] and cnt[] are only ever read around L1 (__ldcg).
    constexpr int K = 4;
    {
        bool settled = false;
        for (uint32_t sweep = 0; sweep < 256u && !settled; ++sweep) {
            bool changed = false;
            for (uint32_t v0 = 0; v0 < N; v0 += 32 * K) {
                uint32_t c[K]; uint4 ra[K], rb[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t v = v0 + k * 32 + lane;
                    c[k] = NIL; ra[k] = make_uint4(NIL, NIL, NIL, NIL); rb[k] = make_uint4(NIL, NIL, NIL, NIL);
                    if (v < N) { c[k] = __ldcg(claim + v); ra[k] = rec4[2 * (size_t)v]; rb[k] = rec4[2 * (size_t)v + 1]; }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (c[k] == NIL) continue;
                    const uint32_t d[7] = {ra[k].x, ra[k].y, ra[k].z, ra[k].w, rb[k].x, rb[k].y, rb[k].z};
                    // claim[d] <= d: only a dependency with a larger id than this node's claim can be lowered by it
#pragma unroll
                    for (int q = 0; q < 7; ++q)
                        if (d[q] != NIL && d[q] > c[k] && atomicMin(claim + d[q], c[k]) > c[k]) changed = true;
                    for (uint32_t x = rb[k].w; x != NIL; x = e_next_in[x]) {              // fifth and later in-edges (rare)
                        const uint32_t b = e_begin[x];
                        if (b > c[k] && atomicMin(claim + b, c[k]) > c[k]) changed = true;
                    }
                }
            }
            __syncwarp();
            settled = !__any_sync(FULL, changed);
        }
        if (!settled) return 0;
    }

    // ---- (3) counts, first ranks, the roots that need no walk, the list of those that do
    for (uint32_t u = lane; u < N; u += 32) atomicAdd(cnt + __ldcg(claim + u), 1u);
    __syncwarp();
    uint32_t running = 0, nroots = 0;
    for (uint32_t i0 = 0; i0 < N; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t c = i < N ? __ldcg(cnt + i) : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t b = running + incl - c;
        if (c != 0u) base[i] = b;
        if (c == 1u) { r2n[b] = i; n2r[i] = b; }                       // claims only itself: no aligned nodes, every in-edge tail emitted earlier
        const unsigned m = __ballot_sync(FULL, c > 1u);
        if (c > 1u) roots[nroots + __popc(m & ((1u << lane) - 1u))] = i;
        nroots += __popc(m);
        running += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    if (running != N) return 0;

    // ---- (2) the walks, one per lane: SPOA's DFS (in-edges in list order, then the aligned nodes; the last one pushed is visited
    //      first; a node whose dependencies are all emitted is final and, unless it was pushed as an aligned sibling, emits itself
    //      and its aligned nodes), looking only at the root's own nodes
    bool ok = true;
    for (uint32_t r0 = 0; r0 < nroots; r0 += 32) {
        const uint32_t ridx = r0 + lane;
        if (ridx < nroots) {
            const uint32_t i = roots[ridx];
            const uint32_t b = base[i], n_own = __ldcg(cnt + i);
            uint32_t* const stk = stk_all + (size_t)TOPO_WALK_STACK * b;
            const uint32_t cap = TOPO_WALK_STACK * n_own, limit = 64u * n_own + 64u;
            uint32_t sp = 1, k = 0, guard = 0;
            stk[0] = i;
            while (sp > 0 && ok) {
                if (++guard > limit) { ok = false; break; }
                const uint32_t v = stk[sp - 1];
                if (mark[v] == 2) { --sp; continue; }
                const uint4 ra = rec4[2 * (size_t)v], rb = rec4[2 * (size_t)v + 1];
                const uint32_t d[7] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z};
                uint32_t cl[7];
#pragma unroll
                for (int q = 0; q < 7; ++q) cl[q] = d[q] != NIL ? __ldcg(claim + d[q]) : 0u;      // all round trips of this visit together
                const bool chk = check[v] != 0;
                bool valid = true;
                auto dep = [&](uint32_t u, uint32_t cu, bool aligned_link) {
                    if (cu < i || mark[u] == 2) return;                 // emitted by an earlier root, or by this walk
                    if (cu != i || sp >= cap) { ok = false; return; }   // cannot happen (the walk stays inside its claim) / stack full: serial fallback
                    stk[sp++] = u; valid = false;
                    if (aligned_link) check[u] = 0;
                };
#pragma unroll
                for (int q = 0; q < 4; ++q) if (d[q] != NIL) dep(d[q], cl[q], false);
                for (uint32_t x = rb.w; x != NIL && ok; x = e_next_in[x]) { const uint32_t u = e_begin[x]; dep(u, __ldcg(claim + u), false); }
                if (chk) {
#pragma unroll
                    for (int q = 4; q < 7; ++q) if (d[q] != NIL) dep(d[q], cl[q], true);
                }
                if (!ok) break;
                if (!valid) {
                    if (mark[v] == 1) { ok = false; break; }            // not a DAG
                    mark[v] = 1;
                    continue;
                }
                mark[v] = 2;
                if (chk) {
                    const uint32_t na = (d[4] != NIL) + (d[5] != NIL) + (d[6] != NIL);         // NIL-terminated
                    if (k + 1u + na > n_own) { ok = false; break; }
                    r2n[b + k] = v; n2r[v] = b + k; ++k;
#pragma unroll
                    for (int q = 4; q < 7; ++q) if (d[q] != NIL) { r2n[b + k] = d[q]; n2r[d[q]] = b + k; ++k; }
                }
                --sp;
            }
            if (k != n_own) ok = false;
        }
        __syncwarp();
    }
    return __all_sync(FULL, ok) ? 1 : 0;
}

}  // namespace hgpu
