// k_poa_pool: the deep-edge kernel. A persistent block of POOL_WARPS warps works on up to POOL_MAX_CTX backbone edges at once
// ("contexts": a POA graph + a score-matrix slot each); its warps pull TASKS from the block's contexts:
//
//   stripe task   one 512-column stripe of the current alignment of a context (rel_stripe<true>): the stripes of an alignment
//                 run on different warps, pipelined 32 rows apart through the boundary columns, as in the team kernel;
//   graph task    everything between two fills of a context - traceback, add_alignment, topological sort, DP records, plan -
//                 and, at the end of an edge, consensus + publish + the next edge from the device queue.
//
// Why: a deep edge (tens of reads over kilobases: where a real 25x dataset spends its DP cells) leaves ~30 % of a lone warp's
// time in serial graph work, and 592-6000 such edges do not fill 148 SMs with one warp each. Warp-per-edge (k_poa_edges_deep)
// idles the machine, block-per-edge (k_poa_edges_team) idles the block during the graph work and when the stripes do not
// divide by the team size. Here the stripes of one edge overlap with the graph work of the others on the same SM, whatever
// the shapes are: any free warp takes whatever task exists. Replaces the pthread edge queue + SPOA engine per thread of
// asm_cal_cns_seq_MT (reference Assemble.cpp:365-434,479-605). Results are identical to the other kernels (same device
// functions; the order of the stripes does not change a cell).
#pragma once

namespace hgpu {

#ifndef HGPU_POOL_WARPS
#define HGPU_POOL_WARPS 16            // one block of 16 warps per SM: an edge whose new alignment needs seven warps at once finds them sooner among 16 than
                                      // among 8 (config 2 K3 355 -> 347 ms against two blocks of 8; profiles/r2E_w16.log, r2G_pool_shape_crit_k1_ab.log)
#endif
static constexpr int POOL_WARPS = HGPU_POOL_WARPS;
#ifndef HGPU_POOL_MAX_CTX
#define HGPU_POOL_MAX_CTX 32          // contexts (edges in flight) per block, two per warp: A/B on config 2 with blocks of 8 warps, 16 vs 8 contexts: 530 vs 552 ms
                                      // (profiles/r2p_ab.log); at most 32: one lane looks at one context when a warp claims a task
#endif
static constexpr int POOL_MAX_CTX = HGPU_POOL_MAX_CTX;
static constexpr uint32_t POOL_MAX_STRIPES = 32;          // alignments with more stripes are filled by one warp, stripe after stripe
enum : uint32_t { PS_IDLE = 0, PS_BUSY = 1, PS_FILL = 2, PS_GRAPH_READY = 3, PS_DONE = 4 };
enum : int { PT_NONE = 0, PT_STRIPE = 1, PT_GRAPH = 2, PT_NEW = 3, PT_EXIT = 4 };

// One size class of deep edges: its own queue, slot size and workspace layout. Classes are ordered by slot size, largest
// first; a context whose own queue has run dry takes edges of the classes after its own (they fit its slot and workspace).
static constexpr int POOL_MAX_CLASSES = 6;
struct PoolClass {
    const uint32_t* items; uint32_t n_items; uint32_t* counter;
    uint8_t* ws; WsLayout wl; uint8_t* arena; uint64_t slot_bytes;
};
#if HGPU_PHASE_CLOCKS
#define POOL_CLK_DECL long long pk_t = clock64();
#define POOL_CLK(a, idx) { const long long pk_n = clock64(); if (lane == 0 && (a).phase_clk) atomicAdd((a).phase_clk + (idx), (unsigned long long)(pk_n - pk_t)); pk_t = pk_n; }
#else
#define POOL_CLK_DECL
#define POOL_CLK(a, idx)
#endif
// developer clocks of the pool (warp-cycles): 0 looking for a task / idle, 1 opening a context, 2 stripe tasks, 3 traceback, 4 add_alignment,
// 5 records + topological sort, 6 DP records + plan, 7 next alignment / consensus / publish / next edge

struct PoolArgs {
    PoaArgs a;                           // everything that is not per class (a.items / a.ws / a.arena / a.wl / a.slot_bytes are unused)
    PoolClass cls[POOL_MAX_CLASSES];
    uint32_t n_cls;
    const uint8_t* ctx_class;            // [blocks * n_ctx] class of every context (0xFF: none, the context stays idle)
    const uint32_t* ctx_slot;            // [blocks * n_ctx] its slot / workspace index inside the class
    const uint32_t* ctx_first;           // [blocks * n_ctx] the edge the context starts with (the host deals the heaviest edges of every class so
                                         // that the blocks' summed work is even), 0xFFFFFFFF: from the class queue like every later edge
};

struct PoolCtx {
    uint32_t state;
    uint32_t edge, k, R, s0;             // edge id, index of the segment being aligned, segments of the edge, its first segment
    uint32_t V, L, NS, mode;             // current alignment
    uint32_t n_tasks;                    // stripe tasks of the current alignment (NS, or 1 when one warp fills all of it)
    uint32_t claim, done;                // (generation << 8) | stripe tasks claimed; finished stripe tasks. The generation changes with
                                         // every alignment, so a claim (a CAS on the whole word) can never cross into the next one
    uint32_t sync_fail;
    uint32_t prio;                       // what is left of the edge's critical path, ~ (alignments to go) x (graph nodes): the claim order
    uint32_t vprog[POOL_MAX_STRIPES];    // rows done + 1 of every stripe (TeamSync)
    unsigned long long cells, padded, aln, aln32, bases;   // of the current edge; added to the block's totals when it completes
};

struct PoolShared {
    PoolCtx ctx[POOL_MAX_CTX];
    unsigned long long tot[5];           // cells, padded, alignments, int32 / REL16 alignments, bases of completed edges
};

__device__ __forceinline__ uint32_t vld(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void vst(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

struct PoolEnv {                         // per-context global storage, bound on demand
    GraphView gv; GraphScratch gs; uint32_t* hdr; uint32_t* plan; uint32_t* tbp; TopoRec* trec; uint8_t* slot; uint64_t slot_bytes; uint32_t cls, gctx;
};
__device__ __forceinline__ PoolEnv pool_env(const PoolArgs& pa, uint32_t gctx) {
    PoolEnv e;
    e.gctx = gctx;
    e.cls = pa.ctx_class[gctx];
    const PoolClass& c = pa.cls[e.cls];
    const uint32_t idx = pa.ctx_slot[gctx];
    uint8_t* wsb = c.ws + (uint64_t)idx * c.wl.bytes;
    e.gv = bind_graph(wsb, c.wl);
    e.gs = bind_scratch(wsb, c.wl);
    e.hdr = reinterpret_cast<uint32_t*>(wsb + c.wl.o_hdr);
    e.plan = reinterpret_cast<uint32_t*>(wsb + c.wl.o_plan);
    e.trec = reinterpret_cast<TopoRec*>(wsb + c.wl.o_trec);
    e.tbp = reinterpret_cast<uint32_t*>(wsb + c.wl.o_tbp);
    e.slot = c.arena + (uint64_t)idx * c.slot_bytes;
    e.slot_bytes = c.slot_bytes;
    return e;
}

// consensus of a finished edge + publish; lane-uniform
__device__ __noinline__ void pool_finish_edge(const PoaArgs& a, PoolEnv& E, PoolCtx* C, PoolShared* sh, uint32_t st, int lane) {
    uint32_t n_cons = 0;
    GraphView& gv = E.gv;
    if (st == ST_OK && vld(&C->R) != 0) {
        if (*gv.n_nodes <= (1u << 20)) {
            const uint32_t max_id = w_consensus_scores(gv, E.gs, lane);
            if (gv.out_head[max_id] == NIL) n_cons = w_consensus_backtrack(gv, E.gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank), lane);
            else if (lane == 0) n_cons = g_consensus_finish(gv, E.gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank));
        } else if (lane == 0) n_cons = g_consensus(gv, E.gs, reinterpret_cast<uint32_t*>(gv.aln_rank));
        n_cons = __shfl_sync(FULL, n_cons, 0);
        __syncwarp();
    }
    unsigned long long pos = 0;
    if (lane == 0 && n_cons > 0) {
        pos = atomicAdd(a.pool_cursor, (unsigned long long)n_cons);
        if (pos + n_cons > a.pool_cap) st = ST_POOL;
    }
    pos = __shfl_sync(FULL, pos, 0);
    st = __shfl_sync(FULL, st, 0);
    if (st == ST_OK && n_cons > 0) {
        const uint32_t* ids = reinterpret_cast<const uint32_t*>(gv.aln_rank);
        for (uint32_t i = lane; i < n_cons; i += 32) a.pool[pos + i] = (uint8_t)"ACGT"[gv.code[ids[i]]];
    }
    if (lane == 0) {
        const uint32_t e = vld(&C->edge);
        a.status[e] = st;
        a.cons_len[e] = (st == ST_OK) ? n_cons : 0;
        a.cons_pos[e] = (uint64_t)(uintptr_t)(a.pool + pos);
        if (a.out_nodes) a.out_nodes[e] = *gv.n_nodes;
        if (a.edge_clk) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.edge_clk[2 * e + 1] = t; }
        if (st == ST_OK) {
            atomicAdd(&sh->tot[0], C->cells); atomicAdd(&sh->tot[1], C->padded); atomicAdd(&sh->tot[2], C->aln);
            atomicAdd(&sh->tot[3], C->aln32); atomicAdd(&sh->tot[4], C->bases);
        }
    }
    __syncwarp();
}

// Bring context C to its next fill: set up alignment k of the current edge, or finish the edge and start the next one from the
// queue (several times over if edges have a single segment). Leaves state = PS_FILL or PS_DONE. Lane-uniform.
__device__ __noinline__ void pool_advance(const PoolArgs& pa, PoolEnv& E, PoolCtx* C, PoolShared* sh, uint32_t st, bool have_edge, int lane) {
    const PoaArgs& a = pa.a;
    GraphView& gv = E.gv;
    while (true) {
        if (have_edge && st == ST_OK && vld(&C->k) < vld(&C->R)) {
            // ---- alignment k against the current graph
            const uint32_t k = vld(&C->k), s0 = vld(&C->s0);
            const uint32_t V = *gv.n_nodes, NE = *gv.n_edges, L = a.seg_len[s0 + k];
            const int mode = dp_mode_deep(a.sc, a.force_i32);
            if ((uint64_t)V + L > gv.ncap || (uint64_t)NE + L + 1 > gv.ecap) st = ST_CAPACITY;
            else if (dp_slot_bytes(V, L, mode) > E.slot_bytes) st = ST_TOO_LARGE;
            else {
                const uint32_t NS = mode == DPM_REL16 ? Geo<DP_NW16, true>::stripes(L) : Geo<DP_NW32, false>::stripes(L);
                if (lane == 0) {
                    C->V = V; C->L = L; C->NS = NS; C->mode = (uint32_t)mode;
                    C->n_tasks = (mode == DPM_REL16 && NS <= POOL_MAX_STRIPES) ? NS : 1u;
                    C->done = 0; C->sync_fail = 0;
                    C->prio = (uint32_t)min((unsigned long long)(vld(&C->R) - k) * ((V >> 4) + 1u), 0xFFFFFFull);
                }
                for (uint32_t s = lane; s < POOL_MAX_STRIPES; s += 32) C->vprog[s] = 0;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    vst(&C->claim, ((vld(&C->claim) >> 8) + 1u) << 8);      // opens the new generation: everything above is visible by now
                    __threadfence_block();
                    vst(&C->state, PS_FILL);
                }
                return;
            }
        }
        // ---- the edge is complete (or failed): consensus, publish, next edge
        if (have_edge) pool_finish_edge(a, E, C, sh, st, lane);
        // next edge: the context's own class first, then the classes with smaller slots (their edges fit)
        uint32_t e = 0xFFFFFFFFu;
        if (lane == 0) {
            if (!have_edge) e = pa.ctx_first[E.gctx];           // the context's first edge was dealt by the host
            for (uint32_t k = E.cls; k < pa.n_cls && e == 0xFFFFFFFFu; ++k) {
                const PoolClass& c = pa.cls[k];
                if (*reinterpret_cast<volatile uint32_t*>(c.counter) >= c.n_items) continue;
                const uint32_t item = atomicAdd(c.counter, 1u);
                if (item < c.n_items) e = c.items[item];
            }
        }
        e = __shfl_sync(FULL, e, 0);
        if (e == 0xFFFFFFFFu) {
            __threadfence_block();
            if (lane == 0) vst(&C->state, PS_DONE);
            return;
        }
        const uint32_t s0 = a.e_seg_off[e], R = a.e_seg_off[e + 1] - s0;
        if (lane == 0) { C->edge = e; C->k = 1; C->R = R; C->s0 = s0; C->cells = 0; C->padded = 0; C->aln = 0; C->aln32 = 0; C->bases = 0; }
        if (lane == 0 && a.edge_clk) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.edge_clk[2 * e] = t; }
        __syncwarp();
        st = ST_OK; have_edge = true;
        if (R == 0) {
            if (lane == 0) { *gv.n_nodes = 0; *gv.n_edges = 0; *gv.aln_len = 0; }
            __syncwarp();
        } else {
            const uint32_t L0 = a.seg_len[s0];
            if (L0 > gv.ncap || L0 > gv.ecap) st = ST_CAPACITY;
            else {
                w_init_chain(gv, a.bases + a.seg_ptr[s0], L0, lane);
                w_build_plan(gv, E.plan, E.tbp, lane);
                if (lane == 0) C->bases = L0;
                __syncwarp();
            }
        }
    }
}

// everything between two fills of a context
__device__ __noinline__ void pool_graph_task(const PoolArgs& pa, PoolEnv& E, PoolCtx* C, PoolShared* sh, uint8_t* wsm, int lane) {
    const PoaArgs& a = pa.a;
    GraphView& gv = E.gv;
    const uint32_t V = vld(&C->V), L = vld(&C->L), k = vld(&C->k), s0 = vld(&C->s0);
    const int mode = (int)vld(&C->mode);
    const uint8_t* seq = a.bases + a.seg_ptr[s0 + k];
    uint32_t st = ST_OK;
    POOL_CLK_DECL
    if (vld(&C->sync_fail)) st = ST_SYNC;
    else {
        const bool ok = mode == DPM_REL16 ? dp_traceback<DP_NW16, true, true, true>(gv, E.slot, wsm, seq, V, L, a.sc, lane, E.tbp)
                                          : dp_traceback<DP_NW32, false, false, true>(gv, E.slot, wsm, seq, V, L, a.sc, lane, E.tbp);
        POOL_CLK(a, 3)
        if (lane == 0) {
            E.hdr[HDR_LAST_P16] = (uint32_t)mode; E.hdr[HDR_LAST_V] = V; E.hdr[HDR_LAST_L] = L; E.hdr[HDR_LAST_BIAS] = 0;
            C->cells += (unsigned long long)(V + 1) * (L + 1);
            C->padded += (unsigned long long)(V + 1) * (mode == DPM_REL16 ? Geo<DP_NW16, true>::stripes(L) * Geo<DP_NW16, true>::SW
                                                                          : Geo<DP_NW32, false>::stripes(L) * Geo<DP_NW32, false>::SW);
            C->aln += 1; C->aln32 += mode == DPM_I32 ? 1ull : (1ull << 32); C->bases += L;
        }
        if (!ok) st = ST_TRACEBACK;
        else {
            uint32_t ust = w_add_alignment(gv, E.gs, seq, L, lane);
            if (ust == 0xFFFFFFFFu) {
                ust = ST_OK;
                if (lane == 0 && !g_add_alignment(gv, seq, L)) ust = ST_CAPACITY;
                ust = __shfl_sync(FULL, ust, 0);
                __syncwarp();
            }
            POOL_CLK(a, 4)
            if (ust == ST_OK) {
                w_build_trec(gv, E.trec, lane);
                if (!(HGPU_TOPO_CLAIMS && w_toposort_claims(gv, E.gs, E.trec, lane)) && !w_toposort(gv, E.trec, wsm, lane)) {
                    if (lane == 0 && !g_toposort(gv, E.gs)) ust = ST_TOPOSORT;
                    ust = __shfl_sync(FULL, ust, 0);
                    __syncwarp();
                }
            }
            POOL_CLK(a, 5)
            if (ust == ST_OK) { w_build_meta(gv, lane); w_build_plan(gv, E.plan, E.tbp, lane); }
            POOL_CLK(a, 6)
            st = ust;
        }
    }
    if (lane == 0) C->k = k + 1;
    __syncwarp();
    pool_advance(pa, E, C, sh, st, true, lane);
    POOL_CLK(a, 7)
}

#ifndef HGPU_POOL_POLICY
#define HGPU_POOL_POLICY 0
#endif
#ifndef HGPU_POOL_BLOCKS_PER_SM
#define HGPU_POOL_BLOCKS_PER_SM 1
#endif
__global__ void __launch_bounds__(32 * POOL_WARPS, HGPU_POOL_BLOCKS_PER_SM) k_poa_pool(const __grid_constant__ PoolArgs pa, uint32_t n_ctx) {
    const PoaArgs& a = pa.a;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = lane_id();
    const uint32_t wib = threadIdx.x >> 5;
    uint8_t* wsm = smem_raw + (size_t)wib * DP_SMEM_PER_WARP_DEEP;
    PoolShared* sh = reinterpret_cast<PoolShared*>(smem_raw + (size_t)POOL_WARPS * DP_SMEM_PER_WARP_DEEP);
    static_assert(POOL_MAX_CTX <= 32, "one lane looks at one context when a warp claims a task");
    if (threadIdx.x < POOL_MAX_CTX) {
        const bool live = threadIdx.x < n_ctx && pa.ctx_class[blockIdx.x * n_ctx + threadIdx.x] != 0xFFu;
        sh->ctx[threadIdx.x].state = live ? PS_IDLE : PS_DONE; sh->ctx[threadIdx.x].claim = 0; sh->ctx[threadIdx.x].n_tasks = 0;
    }
    if (threadIdx.x < 5) sh->tot[threadIdx.x] = 0;
    __syncthreads();
    uint32_t idle_spins = 0;
    POOL_CLK_DECL
    while (true) {
        // ---- claim a task. Every lane looks at one context and offers what it has: a context that was never opened (first: an
        //      edge can only be ranked once its first alignment is set up), else its graph task or its next stripe, ranked by what
        //      is left of the edge's critical path. The longest edges of the block get the warps first and run at their own speed;
        //      the short ones fill what is left and make the tail of the kernel (first come first served left the long ones
        //      sharing warps eight ways, and the kernel waited 60 % of its time for them).
        int kind = PT_NONE; uint32_t c = 0, s = 0;
        {
            uint32_t key = 0, w = 0, st_ = PS_DONE;
            if ((uint32_t)lane < n_ctx) {
                PoolCtx* Cq = &sh->ctx[lane];
                st_ = vld(&Cq->state);
#if HGPU_POOL_POLICY == 3
                const uint32_t uniq = 31u - (uint32_t)lane;
#else
                const uint32_t uniq = (uint32_t)(lane + wib) & 31u;
#endif
                if (st_ == PS_IDLE) key = 0x80000000u | uniq;
#if HGPU_POOL_POLICY == 1
                else if (st_ == PS_GRAPH_READY) key = 0x40000000u | (vld(&Cq->prio) << 6) | uniq;
#elif HGPU_POOL_POLICY == 2
                else if (st_ == PS_GRAPH_READY) key = ((vld(&Cq->prio) >> 4) << 7) | 0x40u | uniq;
#else
                else if (st_ == PS_GRAPH_READY) key = (vld(&Cq->prio) << 7) | 0x40u | uniq;
#endif
                else if (st_ == PS_FILL) {
                    // read the claim word, then the task count of that generation, then claim by CAS on the whole word: if the
                    // alignment changed in between, the generation differs and the CAS fails
                    w = vld(&Cq->claim);
                    __threadfence_block();
#if HGPU_POOL_POLICY == 1
                    if ((w & 0xFFu) < vld(&Cq->n_tasks)) key = (vld(&Cq->prio) << 6) | uniq;
#elif HGPU_POOL_POLICY == 2
                    if ((w & 0xFFu) < vld(&Cq->n_tasks)) key = ((vld(&Cq->prio) >> 4) << 7) | uniq;
#else
                    if ((w & 0xFFu) < vld(&Cq->n_tasks)) key = (vld(&Cq->prio) << 7) | uniq;
#endif
                }
            }
            const uint32_t best = __reduce_max_sync(FULL, key);
            if (best == 0) {
                if (__all_sync(FULL, st_ == PS_DONE)) kind = PT_EXIT;
            } else {
                const int winner = __ffs(__ballot_sync(FULL, key == best)) - 1;
                int got = PT_NONE;
                if (lane == winner) {
                    PoolCtx* Cq = &sh->ctx[lane];
                    if (st_ == PS_IDLE) { if (atomicCAS(&Cq->state, PS_IDLE, PS_BUSY) == PS_IDLE) got = PT_NEW; }
                    else if (st_ == PS_GRAPH_READY) { if (atomicCAS(&Cq->state, PS_GRAPH_READY, PS_BUSY) == PS_GRAPH_READY) got = PT_GRAPH; }
                    else if (atomicCAS(&Cq->claim, w, w + 1u) == w) got = PT_STRIPE;
                }
                kind = __shfl_sync(FULL, got, winner);
                if (kind == PT_NONE) continue;                  // another warp was faster: look again
                c = (uint32_t)winner; s = __shfl_sync(FULL, w, winner) & 0xFFu;
            }
        }
        if (kind == PT_EXIT) break;
        if (kind == PT_NONE) {
            __nanosleep(idle_spins < 64 ? 200 : 2000);
            if (++idle_spins > (1u << 26)) {                  // minutes without a task while edges are open: never hang the GPU, tell the host
                if (lane == 0 && a.stats) atomicExch(a.stats + 7, 1ull);
                break;
            }
            continue;
        }
        idle_spins = 0;
        POOL_CLK(a, 0)
        __threadfence_block();
        PoolCtx* C = &sh->ctx[c];
        PoolEnv E = pool_env(pa, blockIdx.x * n_ctx + c);
        if (kind == PT_NEW) {
            pool_advance(pa, E, C, sh, ST_OK, false, lane);
            POOL_CLK(a, 1)
        } else if (kind == PT_GRAPH) {
            pool_graph_task(pa, E, C, sh, wsm, lane);
#if HGPU_PHASE_CLOCKS
            pk_t = clock64();
#endif
        } else {
            const uint32_t V = vld(&C->V), L = vld(&C->L), k = vld(&C->k), s0 = vld(&C->s0), NS = vld(&C->NS);
            const int mode = (int)vld(&C->mode);
            const uint8_t* seq = a.bases + a.seg_ptr[s0 + k];
            bool ok = true;
            if (mode == DPM_REL16 && vld(&C->n_tasks) == NS) {
                uint32_t* prof = reinterpret_cast<uint32_t*>(wsm);
                RelFrame* frame = reinterpret_cast<RelFrame*>(wsm + Geo<DP_NW16, true>::PROF_BYTES);
                __syncwarp();
                if (lane == 0) rel_frame_init(frame, E.gv, E.plan, E.slot, seq, V, L, a.sc.sm, a.sc.sx);
                RelState S;
                rel_state_init(S, prof, frame, NS, a.sc.g, lane);
                const TeamSync ts{C->vprog, s, s > 0 ? s - 1 : 0u, 0u, 0u};
                ok = rel_stripe<true>(S, prof, frame, s, lane, ts);
            } else if (mode == DPM_REL16) {
                ok = dp_fill_rel<false>(E.gv, E.plan, E.slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, lane, 0, 1, nullptr);
            } else {
                ok = dp_fill<DP_NW32, false>(E.gv, E.slot, wsm, seq, V, L, a.sc, lane, 0, 1, nullptr);
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                if (!ok) vst(&C->sync_fail, 1u);
                const uint32_t d = atomicAdd(&C->done, 1u) + 1u;
                if (d == vld(&C->n_tasks)) { __threadfence_block(); vst(&C->state, PS_GRAPH_READY); }
            }
            __syncwarp();
            POOL_CLK(a, 2)
        }
    }
    POOL_CLK(a, 0)
    __syncthreads();
    if (threadIdx.x == 0 && a.stats) {
        atomicAdd(a.stats + 0, sh->tot[0]); atomicAdd(a.stats + 1, sh->tot[1]); atomicAdd(a.stats + 2, sh->tot[2]);
        atomicAdd(a.stats + 3, sh->tot[3]); atomicAdd(a.stats + 4, sh->tot[4]);
    }
}

}  // namespace hgpu
