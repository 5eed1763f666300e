// Batched partial-order-alignment kernel for sm_100a.
//
// Replaces SPOA's graph-NW engine + graph update + consensus as driven by asm_calc_single_cns_seq
// (reference src/haslr_assemble/src/Assemble.cpp:499-554) and the pthread edge queue around it
// (Assemble.cpp:365-434,562-605). ONE persistent kernel, k_poa_edges: every warp pulls backbone edges from
// a device-side queue and takes one edge from its first supporting segment to its consensus string:
//
//   chain graph of segment 0  ->  for each further segment { graph-NW fill, traceback, add_alignment,
//   topological sort, per-rank DP records }  ->  heaviest-bundle consensus  ->  bytes into the output pool
//
// so warps in their (serial, latency-bound) graph-update phase overlap with warps in the (throughput-bound)
// score-matrix fill, and no host round trip sits between the R alignments of an edge.
//
// Score matrix ("H") layout and arithmetic of the fill — see DESIGN.md:
//  * rows = nodes in topological order (+ virtual row 0), columns = sequence positions; every lane owns CPL
//    consecutive columns of a stripe of SW = 32*CPL columns and keeps the previous row in registers, so the
//    common case (single predecessor = previous rank) touches no memory for its inputs;
//  * cells are stored in "gap-hat" space  Hhat[i][j] = H[i][j] - j*gap (+ bias): the horizontal gap recurrence
//    becomes a plain prefix maximum, done per lane in registers then across lanes with warp shuffles;
//  * int16 cells packed two per 32-bit register, updated with the packed-int16 DPX instructions
//    (VIADD.16x2 / VIADDMNMX.S16x2 / VIMNMX.S16x2). A lane's 16 columns are two runs of 8: word k packs column k
//    (low half) with column k+8 (high half), so the one-column shift of the diagonal move is "the previous word"
//    (no byte permutes) and the in-lane prefix maximum is ONE 7-step chain that scans both runs at once;
//    an int32 instantiation takes alignments whose score range does not fit (SPOA makes the same switch);
//  * every finished row is streamed to the warp's slot of the HBM arena as 512-byte lane-interleaved
//    units (coalesced 16-byte stores); non-adjacent predecessors and the traceback read it back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "poa_graph.cuh"

namespace hgpu {

enum : uint32_t { ST_OK = 0, ST_CAPACITY = 1, ST_TOPOSORT = 2, ST_TRACEBACK = 3, ST_TOO_LARGE = 4, ST_POOL = 5, ST_SYNC = 6 };

// Scores in gap-hat space (host computes them once per call).
struct DpScores {
    int32_t sm;      // match - gap      (diagonal step, bases equal)
    int32_t sx;      // mismatch - gap   (diagonal step, bases differ)
    int32_t g;       // gap              (vertical step); horizontal steps cost 0 in hat space
    int32_t lo_step; // most negative change of Hhat per consumed row (>= 0, magnitude)
    int32_t hi_step; // most positive change of Hhat per consumed column (>= 0)
};

// ---------------------------------------------------------------------------------------------------------
// Per-warp workspace: the POA graph of the edge the warp currently owns + update/consensus scratch.
// ---------------------------------------------------------------------------------------------------------
struct WsLayout {
    uint32_t ncap, ecap, scap;
    uint64_t o_hdr, o_code, o_in_head, o_in_tail, o_out_head, o_aligned, o_e_begin, o_e_end, o_e_w, o_e_next_in, o_e_next_out,
        o_rank2node, o_node2rank, o_meta0, o_pred_off, o_pred_rank, o_sinks, o_aln_rank, o_aln_pos, o_mark, o_check, o_stack,
        o_score, o_pred, o_plan, o_trec, o_tbp;
    uint64_t bytes;
};

__host__ __device__ inline WsLayout ws_layout(uint32_t ncap, uint32_t ecap) {
    WsLayout w;
    w.ncap = ncap; w.ecap = ecap; w.scap = ecap + 4 * ncap + 8;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { uint64_t r = o; o += (bytes + 15) / 16 * 16; return r; };
    const uint64_t n = ncap, e = ecap;
    w.o_hdr = take(64);
    w.o_code = take(n);
    w.o_in_head = take(4 * n); w.o_in_tail = take(4 * n); w.o_out_head = take(4 * n);
    w.o_aligned = take(12 * n);
    w.o_e_begin = take(4 * e); w.o_e_end = take(4 * e); w.o_e_w = take(4 * e); w.o_e_next_in = take(4 * e); w.o_e_next_out = take(4 * e);
    w.o_rank2node = take(4 * n); w.o_node2rank = take(4 * n);
    w.o_meta0 = take(4 * n);
    w.o_pred_off = take(4 * (n + 1)); w.o_pred_rank = take(4 * e); w.o_sinks = take(4 * n);
    w.o_aln_rank = take(4 * n); w.o_aln_pos = take(4 * n);
    w.o_mark = take(n); w.o_check = take(n);
    w.o_stack = take(4 * (uint64_t)w.scap);
    w.o_score = take(8 * n); w.o_pred = take(4 * n);
    w.o_plan = take(4 * n);                                   // deep kernels: per-rank predecessor plan (poa_fill_rel.cuh)
    w.o_trec = take(32 * n);                                  // per-node record of the topological sort (w_build_trec)
    w.o_tbp = take(4 * n);                                    // deep kernels: per-rank predecessor distances in in-edge order for the traceback (w_build_plan)
    w.bytes = (o + 127) / 128 * 128;
    return w;
}

// header words of a workspace
enum : int { HDR_N_NODES = 0, HDR_N_EDGES = 1, HDR_ALN_LEN = 2, HDR_LAST_P16 = 3, HDR_LAST_V = 4, HDR_LAST_L = 5, HDR_LAST_BIAS = 6, HDR_N_SINKS = 7 };

__host__ __device__ inline GraphView bind_graph(uint8_t* base, const WsLayout& w) {
    GraphView g;
    g.ncap = w.ncap; g.ecap = w.ecap;
    uint32_t* hdr = reinterpret_cast<uint32_t*>(base + w.o_hdr);
    g.n_nodes = hdr + HDR_N_NODES; g.n_edges = hdr + HDR_N_EDGES; g.aln_len = hdr + HDR_ALN_LEN; g.n_sinks = hdr + HDR_N_SINKS;
    g.code = base + w.o_code;
    g.in_head = reinterpret_cast<uint32_t*>(base + w.o_in_head);
    g.in_tail = reinterpret_cast<uint32_t*>(base + w.o_in_tail);
    g.out_head = reinterpret_cast<uint32_t*>(base + w.o_out_head);
    g.aligned = reinterpret_cast<uint32_t*>(base + w.o_aligned);
    g.e_begin = reinterpret_cast<uint32_t*>(base + w.o_e_begin);
    g.e_end = reinterpret_cast<uint32_t*>(base + w.o_e_end);
    g.e_w = reinterpret_cast<uint32_t*>(base + w.o_e_w);
    g.e_next_in = reinterpret_cast<uint32_t*>(base + w.o_e_next_in);
    g.e_next_out = reinterpret_cast<uint32_t*>(base + w.o_e_next_out);
    g.rank2node = reinterpret_cast<uint32_t*>(base + w.o_rank2node);
    g.node2rank = reinterpret_cast<uint32_t*>(base + w.o_node2rank);
    g.meta0 = reinterpret_cast<uint32_t*>(base + w.o_meta0);
    g.pred_off = reinterpret_cast<uint32_t*>(base + w.o_pred_off);
    g.pred_rank = reinterpret_cast<uint32_t*>(base + w.o_pred_rank);
    g.sinks = reinterpret_cast<uint32_t*>(base + w.o_sinks);
    g.aln_rank = reinterpret_cast<int32_t*>(base + w.o_aln_rank);
    g.aln_pos = reinterpret_cast<int32_t*>(base + w.o_aln_pos);
    return g;
}

__host__ __device__ inline GraphScratch bind_scratch(uint8_t* base, const WsLayout& w) {
    GraphScratch s;
    s.mark = base + w.o_mark; s.check = base + w.o_check;
    s.stack = reinterpret_cast<uint32_t*>(base + w.o_stack); s.stack_cap = w.scap;
    s.score = reinterpret_cast<int64_t*>(base + w.o_score);
    s.pred = reinterpret_cast<int32_t*>(base + w.o_pred);
    return s;
}

static constexpr unsigned FULL = 0xFFFFFFFFu;
#ifndef HGPU_STCS
#define HGPU_STCS 1
#endif
__device__ __forceinline__ void st_row(uint4* p, uint4 v) {
#if HGPU_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}
static constexpr int TB_ROWS = 32, TB_COLS = 32;
static constexpr uint32_t TBP_GENERIC = 0xC0000000u;       // traceback predecessor word: the row takes the one-lane generic step
#ifndef HGPU_TB_CPASYNC
#define HGPU_TB_CPASYNC 1            // traceback tiles as asynchronous global -> shared copies (LDGSTS); 0 = LDG + STS (A/B: +0.4 % on config 3,
#endif                               // +1.4 % on the deep shape, profiles/r2n_ab_cpasync.log)
#ifndef HGPU_SHALLOW_TREC
#define HGPU_SHALLOW_TREC 1          // the shallow kernel also runs the record-based topological sort (0: the in-list walking one; A/B on config 3 with
                                     // the warp-cooperative walk: 1,506 vs 1,461 GCUPS, profiles/r2o_ab_strec.log)
#endif

// Geometry of one alignment inside a slot. P16: two int16 cells per word; I32: one int32 cell per word.
template <int NW, bool P16>
struct Geo {
    static constexpr int CPL = P16 ? 2 * NW : NW;   // columns per lane per stripe
    static constexpr int SW = 32 * CPL;             // columns per stripe
    static constexpr int UNITS = NW / 4;            // 16-byte units per lane per stripe row
    static constexpr int NEGV = P16 ? (-32768 + 256) : -(1 << 29);
    static constexpr int PROF_BYTES = 4 * NW * 32 * 4;
    __host__ __device__ static uint32_t stripes(uint32_t L) { return (L + 1 + SW - 1) / SW; }
    // bytes of a slot for V nodes and L columns: (V+1) rows * NS stripes * NW*128 B, then NS*(V+1) int32 boundary column
    __host__ __device__ static uint64_t slot_bytes(uint32_t V, uint32_t L) {
        uint64_t ns = stripes(L);
        uint64_t h = (uint64_t)(V + 1) * ns * NW * 128;
        uint64_t bc = ((uint64_t)(V + 1) * ns * 4 + 15) / 16 * 16;
        return h + bc;
    }
    __host__ __device__ static int32_t bias(uint32_t V, const DpScores& sc) {
        return P16 ? (-32768 + 1024 + (int32_t)(V + 2) * sc.lo_step) : 0;
    }
};

// does the int16 range hold for this alignment? (stored = Hhat + bias must stay inside [-32768+1024, 32767-512])
__host__ __device__ inline bool dp_fits16(uint32_t V, uint32_t L, const DpScores& sc) {
    int64_t up = (int64_t)(L + 1) * sc.hi_step;
    return up + (int64_t)(V + 2) * sc.lo_step + 1024 + 512 < 65536;
}

static constexpr int DP_NW16 = 8;   // int16 kernel: 16 columns per lane, 512-column stripes
static constexpr int DP_NW32 = 8;   // int32 kernel:  8 columns per lane, 256-column stripes
#ifndef HGPU_RING
#define HGPU_RING 0
#endif
static constexpr int DP_SMEM_PER_WARP = 6272;  // int16 fill: profile 4 KB + frame 128 B + two parked rows 2 KB; reused by the traceback tile (4384 B) and the sort bitmaps
// Deep edges (tens of supporting reads) have graphs several times wider than the gap: most ranks read predecessor rows 2-8
// ranks back. Their kernel (k_poa_edges_deep, and the team kernel) parks EVERY row in a ring of DP_RING_DEEP rows in shared
// memory, so those reads never leave the SM; a lone warp otherwise waits a full L2 round trip per predecessor row.
#ifndef HGPU_RING_DEEP
#define HGPU_RING_DEEP 8
#endif
static constexpr int DP_RING_DEEP = HGPU_RING_DEEP;
#ifndef HGPU_REL_SMEM_BASES
#define HGPU_REL_SMEM_BASES 1       // the REL fill keeps the row bases of the current and the previous batch (64 words) and the batch's plan words (32) in
                                    // shared memory instead of batch registers read through shuffles: poa_fill_rel.cuh
#endif
static constexpr int REL_BASES_BYTES = HGPU_REL_SMEM_BASES ? 64 * 4 + 32 * 4 : 0;
static constexpr int DP_SMEM_PER_WARP_DEEP = 4096 + 128 + DP_RING_DEEP * 1024 + REL_BASES_BYTES;
static constexpr int DP_WARPS_PER_BLOCK = 4;

// Cell encodings of a stored score matrix. ABS16: int16 Hhat + bias, when the whole range 13(L+1) + 8(V+2) fits.
// REL16: int16 relative to the row's own value at the left edge of the stripe (kept as int32 in the boundary-column
// arrays) - Hhat never decreases along a row, so the in-stripe range is [0, (hi_step + lo_step) * 512] whatever V and L are; rows
// are re-based by (base of predecessor row - base of this row) <= -gap when they are combined. I32: plain int32, for
// scores the packed arithmetic cannot hold (and on request, for tests).
enum : int { DPM_I32 = 0, DPM_ABS16 = 1, DPM_REL16 = 2 };
static constexpr int REL_CLAMP = -16000;             // a re-base below this can never win against the row's own cells (>= 0)
__host__ __device__ inline bool dp_rel_ok(const DpScores& sc) {
    // a row spans at most (hi_step + lo_step) per column inside a stripe: its left-edge cell may sit a vertical gap per
    // column below the cells further right (e.g. Hhat[i][0] = -8 i against Hhat[i][i] = 13 i)
    return sc.hi_step < (1 << 19) && ((int64_t)sc.hi_step + sc.lo_step) * (32 * 2 * DP_NW16) + (int64_t)sc.hi_step + 1024 < -(int64_t)REL_CLAMP;
}
// force: 0 = pick, 1 = int32, 2 = REL16 even where ABS16 would fit (tests)
__host__ __device__ inline int dp_mode(uint32_t V, uint32_t L, const DpScores& sc, int force) {
    if (force == 1 || sc.hi_step >= (1 << 19)) return DPM_I32;
    if (force == 2 && dp_rel_ok(sc)) return DPM_REL16;
    if (dp_fits16(V, L, sc)) return DPM_ABS16;
    return dp_rel_ok(sc) ? DPM_REL16 : DPM_I32;
}
// the deep kernels run every alignment in REL16 (one row body, poa_fill_rel.cuh); int32 only for scores it cannot hold
__host__ __device__ inline int dp_mode_deep(const DpScores& sc, int force) {
    return (force == 1 || sc.hi_step >= (1 << 19) || !dp_rel_ok(sc)) ? DPM_I32 : DPM_REL16;
}
__host__ __device__ inline uint64_t dp_slot_bytes(uint32_t V, uint32_t L, int mode) {
    if (mode == DPM_I32) return Geo<DP_NW32, false>::slot_bytes(V, L);
    // REL16 keeps one more boundary array: the bases of stripe 0 (column 0 of every row)
    return Geo<DP_NW16, true>::slot_bytes(V, L) + (mode == DPM_REL16 ? ((uint64_t)(V + 1) * 4 + 15) / 16 * 16 : 0);
}

struct PoaArgs {
    // input segments (non-empty ones only), CSR by edge
    const uint8_t* bases;
    const uint64_t* seg_ptr;     // offset into bases
    const uint32_t* seg_len;
    const uint32_t* e_seg_off;   // [n_edges+1]
    // work queue
    const uint32_t* items;       // edge ids in processing order
    uint32_t n_items;
    uint32_t* counter;           // cursor, zeroed before the launch
    // per-edge results
    uint32_t* status;
    uint32_t* cons_len;
    uint64_t* cons_pos;          // device address of the edge's consensus bytes (inside some pass's pool)
    uint32_t* out_nodes;         // final node count (may be null)
    uint8_t* pool; uint64_t pool_cap; unsigned long long* pool_cursor;
    // per-warp storage
    uint8_t* ws; WsLayout wl;
    uint8_t* arena; uint64_t slot_bytes;
    DpScores sc;
    unsigned long long* stats;   // [0] cells [1] cells computed incl. padding [2] alignments [3] int32 alignments [4] bases in
    uint32_t stop_round;         // debug: stop after the fill+traceback of this round (0xFFFFFFFF = run to consensus)
    int force_i32;
    unsigned long long* phase_clk; // developer build (-DHGPU_PHASE_CLOCKS=1): per-phase SM cycles summed over warps, [16]
    unsigned long long* edge_clk;  // HGPU_VERBOSE=2: [2 x edges] globaltimer ns at which k_poa_pool started / finished each edge (else null)
    uint32_t probe, probe_round; // developer timing probes (HGPU_PROBE=phase, HGPU_PROBE_ROUND=k): the edge ends in round k after 1 fill, 2 traceback,
                                 // 3 add_alignment, 4 topological sort, 5 DP records; results are invalid
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

template <int NW, bool P16, bool REL = false>
struct SlotView {
    using G = Geo<NW, P16>;
    uint32_t* H;       // score words
    int32_t* bcol;     // [NS][V+1] last column of each stripe (stored space); REL: [NS+1][V+1] row bases per stripe (Hhat, int32)
    uint32_t V, L, NS;
    __device__ __forceinline__ void bind(uint8_t* slot, uint32_t V_, uint32_t L_) {
        V = V_; L = L_; NS = G::stripes(L_);
        H = reinterpret_cast<uint32_t*>(slot);
        bcol = reinterpret_cast<int32_t*>(slot + (uint64_t)(V + 1) * NS * NW * 128);
    }
    __device__ __forceinline__ uint4* row_units(uint32_t i, uint32_t s, int lane) const {
        return reinterpret_cast<uint4*>(H) + ((uint64_t)i * NS + s) * G::UNITS * 32 + lane;
    }
    // one cell, stored space
    __device__ __forceinline__ int load(uint32_t i, uint32_t j) const {
        uint32_t s = j / G::SW, jj = j - s * G::SW, ln = jj / G::CPL, c = jj - ln * G::CPL;
        uint32_t k = P16 ? (c & (uint32_t)(NW - 1)) : c;      // P16: word k = columns k (low half) and k+NW (high half)
        uint32_t w = H[(((uint64_t)i * NS + s) * G::UNITS + (k >> 2)) * 128 + ln * 4 + (k & 3)];
        if (P16) return ((c >= (uint32_t)NW) ? (int)(int16_t)(w >> 16) : (int)(int16_t)(w & 0xFFFFu)) + (REL ? bcol[(uint64_t)s * (V + 1) + i] : 0);
        return (int)w;
    }
};

template <int NW, bool P16>
struct RowOps {
    using G = Geo<NW, P16>;
    // h := max over {diag from src shifted by one column, vert from src}, src being the row held in h itself
    __device__ __forceinline__ static void from_regs(uint32_t (&h)[NW], uint32_t left, const uint32_t* pf, uint32_t g2, int g) {
        if (P16) {
#pragma unroll
            for (int k = NW - 1; k >= 0; --k) {
                uint32_t hs = __byte_perm(k ? h[k > 0 ? k - 1 : 0] : left, h[k], 0x5432);
                uint32_t d = __vadd2(hs, pf[k * 32]);
                h[k] = __viaddmax_s16x2(h[k], g2, d);
            }
        } else {
#pragma unroll
            for (int k = NW - 1; k >= 0; --k) {
                int hs = (int)(k ? h[k > 0 ? k - 1 : 0] : left);
                int d = hs + (int)pf[k * 32];
                h[k] = (uint32_t)max((int)h[k] + g, d);
            }
        }
    }
    // t := max(t, diag/vert from one source word w at position k with its left neighbour `prev`)
    __device__ __forceinline__ static uint32_t acc(uint32_t t, uint32_t prev, uint32_t w, uint32_t pfw, uint32_t g2, int g) {
        if (P16) {
            uint32_t hs = __byte_perm(prev, w, 0x5432);
            return __vimax3_s16x2(t, __vadd2(hs, pfw), __vadd2(w, g2));
        } else {
            return (uint32_t)max((int)t, max((int)prev + (int)pfw, (int)w + g));
        }
    }
    // in-lane inclusive prefix maximum; returns the lane total (max of all the lane's cells)
    __device__ __forceinline__ static int scan(uint32_t (&h)[NW]) {
        if (P16) {
            const uint32_t neg2 = pack2(G::NEGV);
            uint32_t run = neg2;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                uint32_t tsh = __byte_perm(h[k], neg2, 0x1054);   // (lo = NEG, hi = h.lo)
                uint32_t rb = __byte_perm(run, 0, 0x3232);        // broadcast run.hi
                run = __vimax3_s16x2(h[k], tsh, rb);
                h[k] = run;
            }
            return (int)(int16_t)(run >> 16);
        } else {
            int run = G::NEGV;
#pragma unroll
            for (int k = 0; k < NW; ++k) { run = max(run, (int)h[k]); h[k] = (uint32_t)run; }
            return run;
        }
    }
    __device__ __forceinline__ static void apply_carry(uint32_t (&h)[NW], int c) {
        if (P16) {
            uint32_t c2 = pack2(c);
#pragma unroll
            for (int k = 0; k < NW; ++k) h[k] = __vmaxs2(h[k], c2);
        } else {
#pragma unroll
            for (int k = 0; k < NW; ++k) h[k] = (uint32_t)max((int)h[k], c);
        }
    }
    __device__ __forceinline__ static int last_cell(const uint32_t (&h)[NW]) {
        return P16 ? (int)(int16_t)(h[NW - 1] >> 16) : (int)h[NW - 1];
    }
    // the word a lane hands to its right neighbour as diagonal source: only the LAST cell matters
    __device__ __forceinline__ static uint32_t left_from_cell(int v) { return P16 ? ((uint32_t)v << 16) : (uint32_t)v; }
};

// ---------------------------------------------------------------------------------------------------------
// Graph-NW of one segment against the warp's graph: fill the slot, trace back into gv.aln_rank/aln_pos
// (traceback order, graph side as ranks). Returns false if the traceback found no predecessor.
// ---------------------------------------------------------------------------------------------------------
#ifndef HGPU_INLINE_DP
#define HGPU_INLINE_DP 1
#endif
#if HGPU_INLINE_DP
#define DP_INLINE __forceinline__
#else
#define DP_INLINE __noinline__
#endif
// Team mode (k_poa_edges_team): the warps of a team share one alignment; warp `trank` of `tsize` fills stripes
// trank, trank+tsize, ... and stripe s can start a batch of 32 rows once stripe s-1 has published those rows'
// boundary column. vprog[w] = (stripe * (V+1) + rows done) of warp w, monotone, in shared memory.
__device__ __forceinline__ void team_publish(volatile uint32_t* vprog, uint32_t trank, uint32_t value, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) vprog[trank] = value;
}
__device__ __forceinline__ bool team_wait(volatile uint32_t* vprog, uint32_t who, uint32_t need, int lane) {
    int ok = 1;
    if (lane == 0) {
        uint32_t spins = 0;
        while (vprog[who] < need) {
            __nanosleep(64);
            if (++spins > (1u << 25)) { ok = 0; break; }      // never hang the GPU: give up, the edge reports ST_SYNC
        }
    }
    ok = __shfl_sync(FULL, ok, 0);
    __threadfence_block();
    return ok != 0;
}

template <int NW, bool P16>
__device__ DP_INLINE bool dp_fill(const GraphView& gv, uint8_t* slot, uint8_t* wsm, const uint8_t* seq,
                                  uint32_t V, uint32_t L, const DpScores sc, int lane,
                                  uint32_t trank, uint32_t tsize, volatile uint32_t* vprog) {
    using G = Geo<NW, P16>;
    using R = RowOps<NW, P16>;
    bool sync_ok = true;
    uint32_t* prof = reinterpret_cast<uint32_t*>(wsm);                 // [4][NW][32]
    uint4* ring = reinterpret_cast<uint4*>(wsm + G::PROF_BYTES);      // [2][UNITS][32]: rows i-2 / i-3 of the current stripe
    const int g = sc.g;
    const uint32_t g2 = P16 ? pack2(g) : (uint32_t)g;
    SlotView<NW, P16> sv;
    sv.bind(slot, V, L);
    const int bias = G::bias(V, sc);
    const uint32_t NS = sv.NS;

    // =============================== fill ===============================
    for (uint32_t s = trank; s < NS; s += tsize) {
        __syncwarp();
        // --- sequence profile of this stripe: prof[code][k][lane] = hat-score(s) of the lane's k-th word
        {
            const uint32_t j0 = s * G::SW + lane * G::CPL;   // first column owned by this lane
#pragma unroll 4
            for (int k = 0; k < NW; ++k) {
                uint32_t w[4];
                if (P16) {
                    uint32_t ja = j0 + 2 * k, jb = ja + 1;
                    int ca = (ja >= 1 && ja <= L) ? (int)base_code(seq[ja - 1]) : -1;
                    int cb = (jb >= 1 && jb <= L) ? (int)base_code(seq[jb - 1]) : -1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        int va = ca < 0 ? 0 : (ca == c ? sc.sm : sc.sx);
                        int vb = cb < 0 ? 0 : (cb == c ? sc.sm : sc.sx);
                        w[c] = ((uint32_t)va & 0xFFFFu) | ((uint32_t)vb << 16);
                    }
                } else {
                    uint32_t ja = j0 + k;
                    int ca = (ja >= 1 && ja <= L) ? (int)base_code(seq[ja - 1]) : -1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) w[c] = (uint32_t)(ca < 0 ? 0 : (ca == c ? sc.sm : sc.sx));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) prof[(c * NW + k) * 32 + lane] = w[c];
            }
        }
        __syncwarp();
        const int32_t* bc_prev = (s > 0) ? sv.bcol + (uint64_t)(s - 1) * (V + 1) : nullptr;
        int32_t* bc_cur = sv.bcol + (uint64_t)s * (V + 1);

        // --- row 0: Hhat = 0 everywhere
        uint32_t h[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) h[k] = P16 ? pack2(bias) : (uint32_t)bias;
        {
            uint4* dst = sv.row_units(0, s, lane);
#pragma unroll
            for (int u = 0; u < G::UNITS; ++u) dst[u * 32] = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
            if (lane == 31) bc_cur[0] = bias;
        }
        if (tsize > 1) team_publish(vprog, trank, s * (V + 1) + 1, lane);

        const bool has_next_stripe = s + 1 < NS;
        for (uint32_t r0 = 0; r0 < V; r0 += 32) {
            // batched per-rank records: lane q holds rank r0+q
            const uint32_t rr = r0 + lane;
            uint32_t mm0 = 0;
            int bcd = G::NEGV, bcc = G::NEGV;
            if (tsize > 1 && s > 0) {       // rows r0 .. r0+32 of the stripe to the left must be complete
                const uint32_t need_rows = (r0 + 33 < V + 1) ? r0 + 33 : V + 1;
                if (!team_wait(vprog, (trank + tsize - 1) % tsize, (s - 1) * (V + 1) + need_rows, lane)) sync_ok = false;
            }
            if (rr < V) {
                mm0 = gv.meta0[rr];
                if (s > 0) { bcd = bc_prev[rr]; bcc = bc_prev[rr + 1]; }
            }
            const int nb = (V - r0) < 32u ? (int)(V - r0) : 32;
            for (int q = 0; q < nb; ++q) {
                const uint32_t i = r0 + q + 1;
                const uint32_t m0 = __shfl_sync(FULL, mm0, q);
                int diag_in = G::NEGV, carry_in = G::NEGV;
                if (s > 0) {
                    diag_in = __shfl_sync(FULL, bcd, q);      // Hhat[i-1][first col of stripe - 1]
                    carry_in = __shfl_sync(FULL, bcc, q);     // Hhat[i][first col of stripe - 1]
                }
                const uint32_t* pf = prof + (m0 & 3u) * (NW * 32) + lane;
                // ring slot of row i-1 (the row in h); it replaces row i-3 there, so ring reads come first
                uint4* ring_wr = ring + ((i - 1) & 1u) * (G::UNITS * 32) + lane;
                if ((m0 & META_FAST) != 0) {                  // single predecessor = previous rank: registers only
                    if (HGPU_RING) {
#pragma unroll
                        for (int u = 0; u < G::UNITS; ++u) ring_wr[u * 32] = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                    }
                    uint32_t left = __shfl_up_sync(FULL, h[NW - 1], 1);
                    if (lane == 0) left = R::left_from_cell(diag_in);
                    R::from_regs(h, left, pf, g2, g);
                } else {
                    const uint32_t npc = (m0 >> 3) & 3u, d0 = meta_d0(m0);
                    uint32_t t[NW];
#pragma unroll
                    for (int k = 0; k < NW; ++k) t[k] = P16 ? pack2(G::NEGV) : (uint32_t)G::NEGV;
                    uint32_t np = npc, cs = 0, d1 = 0;
                    if (npc == 0) np = 1;
                    if (npc == 2) d1 = meta_d1(m0);
                    if (npc == 3) { cs = gv.pred_off[i - 1]; np = gv.pred_off[i] - cs; }
                    for (uint32_t x = 0; x < np; ++x) {
                        uint32_t dist;                         // rank distance to this predecessor row
                        if (npc == 0) dist = i;
                        else if (npc == 3) dist = i - (gv.pred_rank[cs + x] + 1);
                        else dist = x == 0 ? d0 : d1;
                        if (dist == 1) {
                            uint32_t left = __shfl_up_sync(FULL, h[NW - 1], 1);
                            if (lane == 0) left = R::left_from_cell(diag_in);
#pragma unroll
                            for (int k = 0; k < NW; ++k) t[k] = R::acc(t[k], k ? h[k > 0 ? k - 1 : 0] : left, h[k], pf[k * 32], g2, g);
                        } else {
                            const uint32_t prow = i - dist;
                            // rows i-2 and i-3 live in the shared-memory ring, older ones are re-read from the slot
                            const uint4* src = (HGPU_RING && dist <= 3) ? ring + (prow & 1u) * (G::UNITS * 32) + lane : sv.row_units(prow, s, lane);
                            const uint32_t ustride = 32;
                            uint4 v[G::UNITS];
#pragma unroll
                            for (int u = 0; u < G::UNITS; ++u) v[u] = src[u * ustride];
                            uint32_t left = __shfl_up_sync(FULL, v[G::UNITS - 1].w, 1);
                            int bl = G::NEGV;                  // Hhat[prow][first col of stripe - 1]
                            if (s > 0) {
                                const int ql = q + 1 - (int)dist;
                                bl = ql >= 0 ? __shfl_sync(FULL, bcd, ql & 31) : bc_prev[prow];
                            }
                            if (lane == 0) left = R::left_from_cell(bl);
                            uint32_t prev = left;
#pragma unroll
                            for (int u = 0; u < G::UNITS; ++u) {
                                t[4 * u] = R::acc(t[4 * u], prev, v[u].x, pf[(4 * u) * 32], g2, g);
                                t[4 * u + 1] = R::acc(t[4 * u + 1], v[u].x, v[u].y, pf[(4 * u + 1) * 32], g2, g);
                                t[4 * u + 2] = R::acc(t[4 * u + 2], v[u].y, v[u].z, pf[(4 * u + 2) * 32], g2, g);
                                t[4 * u + 3] = R::acc(t[4 * u + 3], v[u].z, v[u].w, pf[(4 * u + 3) * 32], g2, g);
                                prev = v[u].w;
                            }
                        }
                    }
                    if (HGPU_RING) {
#pragma unroll
                        for (int u = 0; u < G::UNITS; ++u) ring_wr[u * 32] = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                    }
#pragma unroll
                    for (int k = 0; k < NW; ++k) h[k] = t[k];
                }
                // horizontal gaps = prefix maximum in hat space
                // Across lanes the exclusive prefix is needed. Lane totals almost always rise up to the lane holding the
                // row maximum and everything right of it inherits that maximum, so: neighbour total (1 SHFL) + row maximum
                // (1 REDUX) + two ballots decide the common case exactly; any violation falls back to the 5-step scan.
                const int tot = R::scan(h);
                const int nbv = __shfl_up_sync(FULL, tot, 1);
                const int rowmax = __reduce_max_sync(FULL, tot);
                const int prevv = (lane == 0) ? carry_in : nbv;
                const int amax = __ffs(__ballot_sync(FULL, tot == rowmax)) - 1;     // first lane holding the maximum
                int excl;
                if (__ballot_sync(FULL, lane <= amax && tot < prevv) == 0) {
                    excl = (lane <= amax) ? prevv : rowmax;
                } else {
                    int incl = tot;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        int v = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl = max(incl, v);
                    }
                    excl = __shfl_up_sync(FULL, incl, 1);
                    excl = (lane == 0) ? carry_in : max(excl, carry_in);
                }
                R::apply_carry(h, excl);
                // stream the row out
                uint4* dst = sv.row_units(i, s, lane);
#pragma unroll
                for (int u = 0; u < G::UNITS; ++u) st_row(dst + u * 32, make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]));
                if (has_next_stripe && lane == 31) bc_cur[i] = R::last_cell(h);
            }
            if (tsize > 1) team_publish(vprog, trank, s * (V + 1) + ((r0 + 32 < V) ? r0 + 32 : V) + 1, lane);
        }
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();
    return sync_ok;
}

// ---------------------------------------------------------------------------------------------------------
// int16 fill. One warp, stripes of 512 columns, 16 columns per lane as 8 words; word k = column k (low half) and
// column k+8 (high half). Rows i-1 and i-2 of the current stripe stay in registers (A/B, swapping roles every
// row), which serves every fast row and ~85 % of the other rows (a second predecessor two ranks back is what a
// substitution or an indel bubble produces) without touching memory; older predecessor rows are re-read from the
// slot. Shared-memory and slot addresses are carried as plain integers through inline PTX so that the row loop
// holds no address arithmetic beyond one add per pointer.
// ---------------------------------------------------------------------------------------------------------
template <int OFF>
__device__ __forceinline__ uint32_t lds_off(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void stg_cs_v4(uint64_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void stg_cs_v4_512(uint64_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.global.cs.v4.u32 [%0+512], {%1, %2, %3, %4};" :: "l"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void stg_u32(uint64_t addr, uint32_t v) {
    asm volatile("st.global.u32 [%0], %1;" :: "l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldg_w(uint64_t addr, int off) {
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(addr + (uint64_t)off) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg_v4(uint64_t addr, int off512) {
    uint4 v;
    if (off512) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4+512];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
    else asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
    return v;
}

struct Fill16 {
    static constexpr int NW = DP_NW16;
    using G = Geo<DP_NW16, true>;
    static constexpr uint32_t NEG2 = ((uint32_t)(G::NEGV & 0xFFFF)) * 0x10001u;

    // A[k] := max(A[k], diagonal from src shifted one column, vertical from src); hs0 = the word left of src[0]
    __device__ __forceinline__ static void acc(uint32_t (&A)[8], const uint4 s0, const uint4 s1, uint32_t hs0, const uint4 p0, const uint4 p1, uint32_t g2) {
        A[7] = __vimax3_s16x2(A[7], __vadd2(s1.z, p1.w), __vadd2(s1.w, g2));
        A[6] = __vimax3_s16x2(A[6], __vadd2(s1.y, p1.z), __vadd2(s1.z, g2));
        A[5] = __vimax3_s16x2(A[5], __vadd2(s1.x, p1.y), __vadd2(s1.y, g2));
        A[4] = __vimax3_s16x2(A[4], __vadd2(s0.w, p1.x), __vadd2(s1.x, g2));
        A[3] = __vimax3_s16x2(A[3], __vadd2(s0.z, p0.w), __vadd2(s0.w, g2));
        A[2] = __vimax3_s16x2(A[2], __vadd2(s0.y, p0.z), __vadd2(s0.z, g2));
        A[1] = __vimax3_s16x2(A[1], __vadd2(s0.x, p0.y), __vadd2(s0.y, g2));
        A[0] = __vimax3_s16x2(A[0], __vadd2(hs0, p0.x), __vadd2(s0.x, g2));
    }
    // the word "one column to the left" of src[0]: low half = column 15 of the lane to the left (or the stripe
    // boundary value for lane 0), high half = this lane's column 7
    __device__ __forceinline__ static uint32_t left_word(const uint32_t (&src)[8], int boundary, int lane) {
        uint32_t left = __shfl_up_sync(FULL, src[7], 1);
        if (lane == 0) left = (uint32_t)boundary << 16;
        return __byte_perm(left, src[7], 0x5432);
    }
};

// Cold per-alignment state of the int16 fill lives in the warp's shared memory (behind the 4 KB profile) instead of
// registers; the row loop reads it only on its rare paths (batch header, >= 3 predecessors, far predecessor rows).
struct FillFrame16 {
    unsigned long long meta0, pred_off, pred_rank, bc_prev, bc_cur, seq, H, bcol;
    uint32_t V, L, NS; int32_t sm, sx, bias;
};
__device__ __forceinline__ uint32_t lds_u32v(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u32(unsigned long long base, uint32_t index) {
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(base + 4ull * index) : "memory");
    return v;
}
#define FRAME16(field) ((uint32_t)offsetof(FillFrame16, field))

// One row, in place: A = row i-1 on entry and row i on exit (stored to the slot as well). Rows are "parked" in the warp's
// shared memory (two slots, by row parity) when a later row reads them two ranks back: `save` says that row i+1 does, so
// row i-1 is parked before A is overwritten. Predecessor rows other than i-1 stream through two registers, from the
// parked copy (two ranks back, ~85 % of such rows) or from the slot, and are folded into A in place - no second row of
// registers. The body is one small code path on purpose: the kernel is instruction-cache bound otherwise, and a spill
// reload in this loop costs an L2 round trip.
#ifndef HGPU_FILL16_BCO_STORE
#define HGPU_FILL16_BCO_STORE 1   // +2.5 % on config 3 (1,682 -> 1,725 GCUPS at 50k edges, profiles/r2I_shallow_bco_ab.log)
#endif
struct Row16State {
    uint32_t pf_lane;        // shared address of this lane's 16 bytes of prof[0][0]: prof[code][half][lane][4 words]
    uint32_t frame;          // shared address of the FillFrame16; parked row r is at frame + 128 + (r & 1) * 1024 as [half][lane][4 words]
    uint32_t g2;             // packed gap
    uint32_t row_bytes;      // distance between consecutive rows of the stripe in the slot
    uint64_t dst;            // slot address of this lane's first unit of row i
    uint32_t bcx;            // batch register: lane q holds the boundary value of row r0+q+1 in the stripe to the left
    uint32_t bco;            // batch register: lane q collects the last cell of row r0+q+1 (boundary for the next stripe)
    uint32_t npk, pk0, pk1;  // RING > 2, batch registers: lane q holds the predecessor count (0xFF: walk the CSR) and up to four rank
                             // distances (16 bits each) of a row r0+q+1 whose record says "walk the CSR", fetched for 32 rows at once
    int lane; bool has_prev, has_next;
};
static constexpr uint32_t FILL16_PARKED = 128;      // byte offset of the parked rows behind the frame

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// REL (DPM_REL16): cells are relative to the row's own base; diag_in / carry_in then carry the BASES (int32 Hhat) of
// row i-1 and row i, predecessor rows are re-based on the fly and the left boundary of a row is its re-base itself.
// RING: rows parked in shared memory. 2 = rows i-1 / i-2 by parity, parked only when `save` says a later row needs them;
// more = a ring of the last RING rows, every row parked.
template <bool REL, int RING>
__device__ __forceinline__ void row16(uint32_t (&A)[8], Row16State& S, uint32_t m0, int q, uint32_t i,
                                      int diag_in, int carry_in, bool save) {
    using F = Fill16;
    const int lane = S.lane;
    const uint32_t g2 = S.g2;
    const uint32_t pf = S.pf_lane + (m0 & 3u) * (uint32_t)(F::NW * 32 * 4);
    const uint32_t parked = S.frame + FILL16_PARKED + (uint32_t)lane * 16u;
    const uint32_t npc = (m0 >> 3) & 3u;
    const bool fast = (m0 & META_FAST) != 0;
    // ---- predecessor list; is row i-1 one of them?
    uint32_t np = npc == 0 ? 1u : npc, cs = 0;
    unsigned long long prank = 0;
    bool has1 = fast || (npc != 3 && npc != 0 && (meta_d0(m0) == 1 || (npc == 2 && meta_d1(m0) == 1)));
    uint32_t pd01 = 0, pd23 = 0;
    bool packed = false;                                            // RING > 2: the row's CSR distances came with the batch
    if (npc == 3) {
        if (RING > 2) {
            const uint32_t k = __shfl_sync(FULL, S.npk, q);
            if (k != 0xFFu) {
                packed = true; np = k;
                pd01 = __shfl_sync(FULL, S.pk0, q); pd23 = __shfl_sync(FULL, S.pk1, q);
                has1 = (pd01 & 0xFFFFu) == 1u || (np > 1 && (pd01 >> 16) == 1u) || (np > 2 && (pd23 & 0xFFFFu) == 1u) || (np > 3 && (pd23 >> 16) == 1u);
            }
        }
        if (!packed) {
            const unsigned long long poff = lds_u64(S.frame + FRAME16(pred_off));
            prank = lds_u64(S.frame + FRAME16(pred_rank));
            cs = ldg_u32(poff, i - 1);
            np = ldg_u32(poff, i) - cs;
            for (uint32_t x = 0; x < np; ++x) has1 = has1 || ldg_u32(prank, cs + x) + 2 == i;
        }
    }
    auto dist_of = [&](uint32_t x) -> uint32_t {                    // rank distance to the x-th predecessor row
        if (npc == 3) {
            if (RING > 2 && packed) return ((x < 2 ? pd01 : pd23) >> ((x & 1u) * 16u)) & 0xFFFFu;
            return i - (ldg_u32(prank, cs + x) + 1);
        }
        if (npc == 0) return i;                                     // no predecessor: the virtual row 0
        return x == 0 ? meta_d0(m0) : meta_d1(m0);
    };
    if (REL && !S.has_prev) {
        // stripe 0: the row's base is its own column 0, Hhat[i][0] = gap + max over the predecessors' bases; it is laid down
        // here, row by row (lane q of bcx keeps it for the rows of this batch, the batch end stores it for later ones)
        int best = diag_in;
        if (!fast) {
            best = INT32_MIN;
            for (uint32_t x = 0; x < np; ++x) {
                const uint32_t dist = dist_of(x);
                int b = diag_in;
                if (dist != 1) {
                    const int ql = q - (int)dist;
                    b = ql >= 0 ? __shfl_sync(FULL, (int)S.bcx, ql & 31) : (int)ldg_u32(lds_u64(S.frame + FRAME16(bc_prev)), i - dist);
                }
                best = max(best, b);
            }
        }
        carry_in = best + (int)(int16_t)(g2 & 0xFFFFu);
        if (lane == q) S.bcx = (uint32_t)carry_in;
    }
    uint32_t hs0 = 0;
    if (!REL) hs0 = F::left_word(A, diag_in, lane);
    if (RING > 2 || save) {
        sts_v4(parked + ((i - 1) & (uint32_t)(RING - 1)) * 1024u, A[0], A[1], A[2], A[3]);
        sts_v4(parked + ((i - 1) & (uint32_t)(RING - 1)) * 1024u + 512u, A[4], A[5], A[6], A[7]);
    }
    if (REL && has1) {                                              // row i-1 into this row's frame
        const int d1 = max(diag_in - carry_in, REL_CLAMP);
        const uint32_t d2 = pack2(d1);
#pragma unroll
        for (int k = 0; k < 8; ++k) A[k] = __vadd2(A[k], d2);
        hs0 = F::left_word(A, S.has_prev ? d1 : F::G::NEGV, lane);
    }
    const uint4 p0 = lds_v4(pf), p1 = lds_v4(pf + 512u);           // the row's profile: scores of the node base against the lane's 16 columns
    // ---- row i-1 (registers): diagonal and vertical moves, in place
    if (has1) {
        A[7] = __viaddmax_s16x2(A[7], g2, __vadd2(A[6], p1.w));
        A[6] = __viaddmax_s16x2(A[6], g2, __vadd2(A[5], p1.z));
        A[5] = __viaddmax_s16x2(A[5], g2, __vadd2(A[4], p1.y));
        A[4] = __viaddmax_s16x2(A[4], g2, __vadd2(A[3], p1.x));
        A[3] = __viaddmax_s16x2(A[3], g2, __vadd2(A[2], p0.w));
        A[2] = __viaddmax_s16x2(A[2], g2, __vadd2(A[1], p0.z));
        A[1] = __viaddmax_s16x2(A[1], g2, __vadd2(A[0], p0.y));
        A[0] = __viaddmax_s16x2(A[0], g2, __vadd2(hs0, p0.x));
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) A[k] = F::NEG2;
    }
    // ---- every other predecessor row streams through two registers and is folded into A
    if (!fast) {
#pragma unroll 1
        for (uint32_t x = 0; x < np; ++x) {
            const uint32_t dist = dist_of(x);
            if (dist == 1) continue;
            int bl = F::G::NEGV;                                    // Hhat[i - dist][first column of the stripe - 1]
            if (REL || S.has_prev) {
                const int ql = q - (int)dist;                       // bcx lane that holds row i - dist
                bl = ql >= 0 ? __shfl_sync(FULL, (int)S.bcx, ql & 31) : (int)ldg_u32(lds_u64(S.frame + FRAME16(bc_prev)), i - dist);
            }
            const bool near = RING > 2 ? dist <= (uint32_t)RING : dist == 2;
            uint4 s0, s1;
            if (near) {                                             // a parked row (RING 2: row i-2)
                const uint32_t pa = parked + ((i - dist) & (uint32_t)(RING - 1)) * 1024u;
                s0 = lds_v4(pa); s1 = lds_v4(pa + 512u);
            } else {                                                // a far row (3.6 % of rows): back from the slot
                const uint64_t src = S.dst - (uint64_t)dist * S.row_bytes;
                s0 = ldg_v4(src, 0); s1 = ldg_v4(src, 1);
            }
            if (REL) {                                              // bl is that row's base: re-base its cells, its boundary is the re-base
                const int dp = max(bl - carry_in, REL_CLAMP);
                const uint32_t d2 = pack2(dp);
                s0.x = __vadd2(s0.x, d2); s0.y = __vadd2(s0.y, d2); s0.z = __vadd2(s0.z, d2); s0.w = __vadd2(s0.w, d2);
                s1.x = __vadd2(s1.x, d2); s1.y = __vadd2(s1.y, d2); s1.z = __vadd2(s1.z, d2); s1.w = __vadd2(s1.w, d2);
                bl = S.has_prev ? dp : F::G::NEGV;
            }
            uint32_t left = __shfl_up_sync(FULL, s1.w, 1);
            if (lane == 0) left = (uint32_t)bl << 16;
            F::acc(A, s0, s1, __byte_perm(left, s1.w, 0x5432), p0, p1, g2);
        }
    }
    // horizontal gaps = prefix maximum in hat space. In the lane: one chain scans columns 0..7 (low halves) and 8..15
    // (high halves) together; the high run then also takes the low run's total.
#pragma unroll
    for (int k = 1; k < 8; ++k) A[k] = __vmaxs2(A[k], A[k - 1]);
    const uint32_t both = __vmaxs2(A[7], __byte_perm(A[7], 0, 0x1032));
    const int tot = (int)(int16_t)(both & 0xFFFFu);                // the lane's maximum
    // Across lanes the exclusive prefix is needed. Lane totals almost always rise up to the lane holding the row
    // maximum and everything right of it inherits that maximum, so: neighbour total (1 SHFL) + row maximum (1 REDUX)
    // + two ballots decide the common case exactly; any violation falls back to the 5-step scan.
    const int row_base = carry_in;                                  // REL only
    if (REL) carry_in = S.has_prev ? 0 : F::G::NEGV;                // the cell left of the stripe is the base itself
    const int nbv = __shfl_up_sync(FULL, tot, 1);
    const int rowmax = __reduce_max_sync(FULL, tot);
    const int prevv = (lane == 0) ? carry_in : nbv;
    const int amax = __ffs(__ballot_sync(FULL, tot == rowmax)) - 1;   // first lane holding the maximum
    int excl;
    if (__ballot_sync(FULL, lane <= amax && tot < prevv) == 0) {
        excl = (lane <= amax) ? prevv : rowmax;
    } else {
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl = max(incl, v);
        }
        excl = __shfl_up_sync(FULL, incl, 1);
        excl = (lane == 0) ? carry_in : max(excl, carry_in);
    }
    const uint32_t mid = __byte_perm(A[7], F::NEG2, 0x1054);        // (low = NEG, high = total of the low run)
    const uint32_t cc = __vmaxs2(mid, __byte_perm((uint32_t)excl, 0, 0x1010));
#pragma unroll
    for (int k = 0; k < 8; ++k) A[k] = __vmaxs2(A[k], cc);
    // stream the row out; its last cell (lane 31) is parked in lane q of bco until the batch ends
    stg_cs_v4(S.dst, A[0], A[1], A[2], A[3]);
    stg_cs_v4_512(S.dst, A[4], A[5], A[6], A[7]);
    S.dst += S.row_bytes;
#if HGPU_FILL16_BCO_STORE
    // lane 31 stores the row's boundary for the next stripe itself (as the REL fill does, poa_fill_rel.cuh) instead of shuffling it
    // into a batch register that is stored every 32 rows
    if (S.has_next && lane == 31) {
        const unsigned long long bc = lds_u64(S.frame + FRAME16(bc_cur));
        asm volatile("st.global.u32 [%0], %1;" :: "l"(bc + 4ull * i), "r"((uint32_t)(((int32_t)A[7] >> 16) + (REL ? row_base : 0))) : "memory");
    }
#else
    if (S.has_next) {
        const uint32_t last = __shfl_sync(FULL, A[7], 31);
        if (lane == q) S.bco = (uint32_t)(((int32_t)last >> 16) + (REL ? row_base : 0));
    }
#endif
}

// does the rank with record m0 read the row two ranks back? (class 3 walks the CSR: assume yes)
__device__ __forceinline__ bool meta_reads_two_back(uint32_t m0) {
    const uint32_t npc = (m0 >> 3) & 3u;
    if (npc == 0) return (m0 & META_FAST) == 0;      // a later rank without predecessors reads the virtual row 0
    return npc == 3 || meta_d0(m0) == 2 || (npc == 2 && meta_d1(m0) == 2);
}

// sequence profile of stripe s: prof[code][k / 4][lane][k % 4] = hat scores of the lane's k-th word (columns k, k+8)
__device__ __noinline__ void fill16_profile(uint32_t* prof, const FillFrame16* frame, uint32_t s, int lane) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    const uint8_t* seq = reinterpret_cast<const uint8_t*>((uintptr_t)frame->seq);
    const uint32_t L = frame->L;
    const int sm = frame->sm, sx = frame->sx;
    const uint32_t j0 = s * G::SW + lane * G::CPL;   // first column owned by this lane
#pragma unroll 2
    for (int k = 0; k < NW; ++k) {
        const uint32_t ja = j0 + k, jb = ja + NW;
        const int ca = (ja >= 1 && ja <= L) ? (int)base_code(seq[ja - 1]) : -1;
        const int cb = (jb >= 1 && jb <= L) ? (int)base_code(seq[jb - 1]) : -1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int va = ca < 0 ? 0 : (ca == c ? sm : sx);
            const int vb = cb < 0 ? 0 : (cb == c ? sm : sx);
            prof[((c * 2 + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = ((uint32_t)va & 0xFFFFu) | ((uint32_t)vb << 16);
        }
    }
}

// The fill of one alignment. Everything that is not needed row by row (graph arrays, slot geometry, the sequence, the
// scores) is parked in the shared-memory frame first, so the row loop keeps its working set in registers.
template <bool TEAM, bool REL = false, int RING = 2>
__device__ __noinline__ bool dp_fill16(const uint32_t* meta0, const uint32_t* pred_off, const uint32_t* pred_rank, uint8_t* slot, uint8_t* wsm,
                                       const uint8_t* seq, uint32_t V, uint32_t L, int sm, int sx, int gap, int bias, int lane,
                                       uint32_t trank, uint32_t tsize, volatile uint32_t* vprog) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    bool sync_ok = true;
    uint32_t* prof = reinterpret_cast<uint32_t*>(wsm);                 // [4][NW][32]
    FillFrame16* frame = reinterpret_cast<FillFrame16*>(wsm + G::PROF_BYTES);
    const uint32_t NS = G::stripes(L);
    __syncwarp();
    if (lane == 0) {
        frame->meta0 = (unsigned long long)(uintptr_t)meta0;
        frame->pred_off = (unsigned long long)(uintptr_t)pred_off;
        frame->pred_rank = (unsigned long long)(uintptr_t)pred_rank;
        frame->seq = (unsigned long long)(uintptr_t)seq;
        frame->H = (unsigned long long)(uintptr_t)slot;
        frame->bcol = (unsigned long long)(uintptr_t)(slot + (uint64_t)(V + 1) * NS * NW * 128);
        frame->V = V; frame->L = L; frame->NS = NS; frame->sm = sm; frame->sx = sx; frame->bias = bias;
    }
    Row16State S;
    S.lane = lane;
    S.g2 = pack2(gap);
    S.pf_lane = (uint32_t)__cvta_generic_to_shared(prof + lane * 4);
    S.frame = (uint32_t)__cvta_generic_to_shared(frame);
    S.row_bytes = NS * (uint32_t)(G::UNITS * 32 * 16);
    S.bcx = 0; S.bco = 0;
    asm volatile("" : "+r"(S.pf_lane), "+r"(S.frame), "+r"(S.g2), "+r"(S.row_bytes));   // kept in registers, never recomputed in the row loop

    for (uint32_t s = TEAM ? trank : 0u; s < lds_u32v(S.frame + FRAME16(NS)); s += TEAM ? tsize : 1u) {
        __syncwarp();
        fill16_profile(prof, frame, s, lane);
        const uint32_t Vs = lds_u32v(S.frame + FRAME16(V));
        {
            const unsigned long long bcol = lds_u64(S.frame + FRAME16(bcol));
            if (lane == 0) {
                if (REL) {                                          // array s = bases of stripe s, array s+1 = bases of the next one
                    frame->bc_prev = bcol + 4ull * s * (Vs + 1);
                    frame->bc_cur = bcol + 4ull * (s + 1) * (Vs + 1);
                } else {
                    frame->bc_prev = s > 0 ? bcol + 4ull * (s - 1) * (Vs + 1) : 0ull;
                    frame->bc_cur = bcol + 4ull * s * (Vs + 1);
                }
            }
        }
        __syncwarp();
        asm volatile("" : "+r"(S.pf_lane) :: "memory");              // profile loads below may not move above / be merged across this point
        S.has_prev = s > 0;
        S.has_next = s + 1 < lds_u32v(S.frame + FRAME16(NS));

        // --- row 0: Hhat = 0 everywhere
        uint32_t A[NW];
        {
            const uint32_t b2 = pack2((int)lds_u32v(S.frame + FRAME16(bias)));
#pragma unroll
            for (int k = 0; k < NW; ++k) A[k] = b2;
            S.dst = lds_u64(S.frame + FRAME16(H)) + ((uint64_t)s * (G::UNITS * 32) + lane) * 16;       // row 0 of this stripe
            stg_cs_v4(S.dst, b2, b2, b2, b2);
            stg_cs_v4_512(S.dst, b2, b2, b2, b2);
            S.dst += S.row_bytes;
            if (lane == 31) asm volatile("st.global.u32 [%0], %1;" :: "l"(lds_u64(S.frame + FRAME16(bc_cur))), "r"(lds_u32v(S.frame + FRAME16(bias))) : "memory");
            if (REL && s == 0 && lane == 0) asm volatile("st.global.u32 [%0], %1;" :: "l"(lds_u64(S.frame + FRAME16(bc_prev))), "r"(0) : "memory");   // base of row 0
        }
        if (TEAM) team_publish(vprog, trank, s * (Vs + 1) + 1, lane);

        // per-rank records of the first batch; later batches are fetched one batch ahead
        uint32_t nm0 = ((uint32_t)lane < Vs) ? ldg_u32(lds_u64(S.frame + FRAME16(meta0)), lane) : 0u;
        int diag = G::NEGV;                                           // boundary value of row i-1: starts at row 0
        if (s > 0) {
            if (TEAM && !team_wait(vprog, (trank + tsize - 1) % tsize, (s - 1) * (Vs + 1) + 1, lane)) sync_ok = false;
            diag = (int)ldg_u32(lds_u64(S.frame + FRAME16(bc_prev)), 0);
        }
        if (REL) diag = 0;                                            // base of row 0
#pragma unroll 1
        for (uint32_t r0 = 0; r0 < Vs; r0 += 32) {
            const uint32_t rr = r0 + lane;
            const uint32_t mm0 = nm0;
            if (rr + 32 < Vs) nm0 = ldg_u32(lds_u64(S.frame + FRAME16(meta0)), rr + 32);
            if (s > 0) {
                if (TEAM) {                 // rows r0 .. r0+32 of the stripe to the left must be complete
                    const uint32_t need_rows = (r0 + 33 < Vs + 1) ? r0 + 33 : Vs + 1;
                    if (!team_wait(vprog, (trank + tsize - 1) % tsize, (s - 1) * (Vs + 1) + need_rows, lane)) sync_ok = false;
                }
                S.bcx = (rr < Vs) ? ldg_u32(lds_u64(S.frame + FRAME16(bc_prev)), rr + 1) : 0u;
            } else if (REL) {
                S.bcx = 0u;                 // stripe 0: the bases are produced by the rows themselves
            }
            if (RING > 2) {                 // rows that "walk the CSR": count and distances of up to four predecessors, 32 rows at once
                S.npk = 0xFFu; S.pk0 = 0u; S.pk1 = 0u;
                if (rr < Vs && ((mm0 >> 3) & 3u) == 3u) {
                    const unsigned long long poff = lds_u64(S.frame + FRAME16(pred_off)), prk = lds_u64(S.frame + FRAME16(pred_rank));
                    const uint32_t c0 = ldg_u32(poff, rr), n = ldg_u32(poff, rr + 1) - c0;
                    if (n <= 4u) {
                        uint32_t d[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (uint32_t x = 0; x < 4u; ++x) if (x < n) d[x] = rr - ldg_u32(prk, c0 + x);
                        if ((d[0] | d[1] | d[2] | d[3]) <= 0xFFFFu) { S.npk = n; S.pk0 = d[0] | (d[1] << 16); S.pk1 = d[2] | (d[3] << 16); }
                    }
                }
            }
            const int nb = (Vs - r0) < 32u ? (int)(Vs - r0) : 32;
            // bit q of save_mask: row r0+q+1 parks row r0+q in B before overwriting it, because row r0+q+2 reads two ranks back
            const uint32_t save_mask = RING == 2 ? __ballot_sync(FULL, meta_reads_two_back(mm0)) >> 1 : 0u;
#pragma unroll 1
            for (int q = 0; q < nb; ++q) {
                const uint32_t m0 = __shfl_sync(FULL, mm0, q);
                int carry = G::NEGV;
                if (REL || S.has_prev) carry = __shfl_sync(FULL, (int)S.bcx, q);
                bool save = ((save_mask >> q) & 1u) != 0;
                if (RING == 2 && q == 31) save = (__ballot_sync(FULL, rr + 32 < Vs && meta_reads_two_back(nm0)) & 1u) != 0;   // first row of the next batch
                row16<REL, RING>(A, S, m0, q, r0 + q + 1, diag, carry, save);
                diag = (REL && !S.has_prev) ? __shfl_sync(FULL, (int)S.bcx, q) : carry;      // base of the row just done
            }
            if (REL && !S.has_prev && lane < nb) {
                const unsigned long long bp = lds_u64(S.frame + FRAME16(bc_prev));
                asm volatile("st.global.u32 [%0], %1;" :: "l"(bp + 4ull * (rr + 1)), "r"(S.bcx) : "memory");
            }
            if (REL && !S.has_prev) __syncwarp();                     // later rows read these through other lanes' loads
#if !HGPU_FILL16_BCO_STORE
            if (S.has_next && lane < nb) {
                const unsigned long long bc = lds_u64(S.frame + FRAME16(bc_cur));
                asm volatile("st.global.u32 [%0], %1;" :: "l"(bc + 4ull * (rr + 1)), "r"(S.bco) : "memory");
            }
#endif
            if (TEAM) team_publish(vprog, trank, s * (Vs + 1) + ((r0 + 32 < Vs) ? r0 + 32 : Vs) + 1, lane);
        }
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();
    return sync_ok;
}

}  // namespace hgpu
#include "poa_fill_rel.cuh"
namespace hgpu {

// ---------------------------------------------------------------------------------------------------------
// Traceback. The stored matrix is read in tiles of 32 rows x 32 columns with the current cell in the corner. A
// tile is kept as the raw 16-byte units of the slot (6 or 10 coalesced LDG.128 + STS.128 per lane, no unpacking),
// and the rows the walk will most likely need next are prefetched into L2 while the current tile is walked.
// Inside a tile every lane t decides the move of "its" cell (ci - t, cj - t) in SPOA's preference order (diagonal
// over the predecessors in in-edge order, then vertical, then horizontal): as long as the lanes before it all moved
// to the previous rank diagonally, its cell is on the path, so one ballot consumes the whole run of such moves
// plus the first move of another kind. Only rows with three or more predecessors, or predecessors outside the
// tile, take the one-lane generic step.
// ---------------------------------------------------------------------------------------------------------
#if HGPU_PHASE_CLOCKS
#define TB_COUNT(x) ++(x)
#else
#define TB_COUNT(x)
#endif
#if HGPU_PHASE_CLOCKS
__device__ unsigned long long* g_tb_counters = nullptr;   // developer build: [4] tiles loaded, walk iterations, generic steps, path length
#endif
template <int NW, bool P16, bool REL = false, bool TBP = false>
struct TbTile {
    using G = Geo<NW, P16>;
    static constexpr int NG = (TB_COLS + G::CPL - 2) / G::CPL + 1;        // column groups a 32-column window can touch (3 / 5)
    static constexpr int LDW = NG * NW + 4;                               // row stride in words (16-byte multiple, spreads banks)
    static constexpr int BYTES = TB_ROWS * LDW * 4 + TB_ROWS * 4 + TB_COLS + (REL ? 2 * TB_ROWS * 4 : 0) + (TBP ? TB_ROWS * 4 : 0);   // REL: row bases of the (at most two) stripes a tile touches; TBP: the rows' predecessor words
    static_assert(BYTES <= DP_SMEM_PER_WARP, "traceback tile does not fit the warp's shared memory");
};

template <int NW, bool P16, bool REL = false, bool TBP = false>
__device__ DP_INLINE bool dp_traceback(const GraphView& gv, uint8_t* slot, uint8_t* wsm, const uint8_t* seq,
                                       uint32_t V, uint32_t L, const DpScores sc, int lane, const uint32_t* tbp = nullptr) {
    using G = Geo<NW, P16>;
    using T = TbTile<NW, P16, REL, TBP>;
    const int g = sc.g;
    // the view's address escapes to the graph functions, so the compiler re-reads its fields from local memory around every
    // store: keep what the walk uses in registers
    const uint32_t* const T_meta0 = gv.meta0; const uint32_t* const T_pred_off = gv.pred_off; const uint32_t* const T_pred_rank = gv.pred_rank;
    const uint32_t* const T_sinks = gv.sinks; int32_t* const T_aln_rank = gv.aln_rank; int32_t* const T_aln_pos = gv.aln_pos;
    const uint32_t T_ncap = gv.ncap;
    SlotView<NW, P16, REL> sv;
    sv.bind(slot, V, L);
    // end cell: best Hhat[i][L] over sink nodes, first maximum in rank order (SPOA kNW)
    int best = INT32_MIN; uint32_t best_i = 0;
    const uint32_t n_sinks = *gv.n_sinks;
    for (uint32_t x = lane; x < n_sinks; x += 32) {                      // the sink list is built with the DP records
        const uint32_t r = T_sinks[x];
        const int v = sv.load(r + 1, L);
        if (v > best) { best = v; best_i = r + 1; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int ov = __shfl_xor_sync(FULL, best, d);
        uint32_t oi = __shfl_xor_sync(FULL, best_i, d);
        if (oi != 0 && (best_i == 0 || ov > best || (ov == best && oi < best_i))) { best = ov; best_i = oi; }
    }
    uint32_t* tile = reinterpret_cast<uint32_t*>(wsm);                    // [TB_ROWS][LDW] raw words: row it - a, groups g0 ..
    uint32_t* tm0 = tile + TB_ROWS * T::LDW;
    uint8_t* tseq = reinterpret_cast<uint8_t*>(tm0 + TB_ROWS);
    int32_t* tbase = reinterpret_cast<int32_t*>(tseq + TB_COLS);         // REL: [2][TB_ROWS] bases of the tile's rows in its first / second stripe
    uint32_t* ttp = reinterpret_cast<uint32_t*>(tbase + (REL ? 2 * TB_ROWS : 0));   // TBP: predecessor words of the tile's rows
    uint32_t J0 = TB_ROWS, J1 = TB_ROWS, J2 = TB_ROWS, J3 = TB_ROWS, J4 = TB_ROWS;  // TBP: lane a holds the tile row 2^k first-predecessor hops from row a (32: outside)
    const uint4* Hu = reinterpret_cast<const uint4*>(sv.H);
    const uint32_t NS = sv.NS;
    uint32_t ci = best_i, cj = L, n_out = 0;
    bool bad = (best_i == 0);
#if HGPU_PHASE_CLOCKS
    uint32_t tbc_tiles = 0, tbc_iters = 0, tbc_generic = 0;
#endif
    while (!bad && !(ci == 0 && cj == 0)) {
        TB_COUNT(tbc_tiles);
        // ---- tile with (ci, cj) in its corner: rows it-31 .. it, columns jt-31 .. jt
        const uint32_t it = ci, jt = cj;
        const uint32_t g0 = (jt >= (uint32_t)(TB_COLS - 1) ? jt - (TB_COLS - 1) : 0u) / G::CPL;   // first column group of the tile
        __syncwarp();
        if (it >= (uint32_t)lane) {
            const uint32_t row = it - lane;
#pragma unroll
            for (int gi = 0; gi < T::NG; ++gi) {
                const uint32_t gg = g0 + gi;
                if (gg * G::CPL <= jt) {
                    const uint4* src = Hu + ((uint64_t)row * NS + gg / 32u) * (G::UNITS * 32) + (gg & 31u);
                    uint4* dst = reinterpret_cast<uint4*>(tile + lane * T::LDW + gi * NW);
#if HGPU_TB_CPASYNC
                    // A/B: the same 16-byte pieces as asynchronous global -> shared copies (LDGSTS), no register staging
#pragma unroll
                    for (int u = 0; u < G::UNITS; ++u)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst + u)), "l"(src + u * 32) : "memory");
#else
#pragma unroll
                    for (int u = 0; u < G::UNITS; ++u) dst[u] = src[u * 32];
#endif
                }
            }
            if (row >= 1) tm0[lane] = T_meta0[row - 1];
            if (TBP) {
                const uint32_t tp = row >= 1 ? tbp[row - 1] : TBP_GENERIC;
                ttp[lane] = tp;
                J0 = (tp >> 30) ? (uint32_t)TB_ROWS : min((uint32_t)lane + (tp & 31u), (uint32_t)TB_ROWS);
            }
            if (REL) {
                const uint32_t sA = g0 / 32u;
                tbase[lane] = sv.bcol[(uint64_t)sA * (V + 1) + row];
                tbase[TB_ROWS + lane] = sv.bcol[(uint64_t)(sA + 1) * (V + 1) + row];
            }
        }
        if (TBP) {
            if (it < (uint32_t)lane) J0 = TB_ROWS;
            // hop tables by doubling: J(k+1)[a] = Jk[Jk[a]]
            uint32_t y;
            y = __shfl_sync(FULL, J0, J0 & 31u); J1 = J0 >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : y;
            y = __shfl_sync(FULL, J1, J1 & 31u); J2 = J1 >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : y;
            y = __shfl_sync(FULL, J2, J2 & 31u); J3 = J2 >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : y;
            y = __shfl_sync(FULL, J3, J3 & 31u); J4 = J3 >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : y;
        }
#if HGPU_TB_CPASYNC
        asm volatile("cp.async.wait_all;" ::: "memory");
#endif
        if (jt >= (uint32_t)lane + 1) tseq[lane] = (uint8_t)base_code(seq[jt - lane - 1]);
        // prefetch what the walk will most likely read next: the rows above the tile, one tile to the left
        if (it >= (uint32_t)(TB_ROWS - 1) + lane && jt >= (uint32_t)(TB_COLS - 1)) {
            const uint32_t prow = it - (TB_ROWS - 1) - lane, pj = jt - (TB_COLS - 1);
            const uint32_t pg0 = (pj >= (uint32_t)(TB_COLS - 1) ? pj - (TB_COLS - 1) : 0u) / G::CPL;
#pragma unroll
            for (int gi = 0; gi < T::NG; ++gi) {
                const uint32_t gg = pg0 + gi;
                if (gg * G::CPL <= pj) {
                    const uint4* src = Hu + ((uint64_t)prow * NS + gg / 32u) * (G::UNITS * 32) + (gg & 31u);
#pragma unroll
                    for (int u = 0; u < G::UNITS; ++u) asm volatile("prefetch.global.L2 [%0];" :: "l"(src + u * 32));
                }
            }
        }
        __syncwarp();
        // stored value of cell (row it - a, column j) of the tile
        auto cell = [&](uint32_t a, uint32_t j) -> int {
            const uint32_t gi = j / G::CPL - g0, c = j % G::CPL;
            const uint32_t w = tile[a * T::LDW + gi * NW + (P16 ? (c & (uint32_t)(NW - 1)) : c)];
            if (P16) return ((c >= (uint32_t)NW) ? (int)(int16_t)(w >> 16) : (int)(int16_t)(w & 0xFFFFu))
                            + (REL ? tbase[((gi + g0) / 32u - g0 / 32u) * TB_ROWS + a] : 0);
            return (int)w;
        };
        while (true) {
            if (ci == 0 && cj == 0) break;
            const uint32_t li = it - ci, lj = jt - cj;
            // ---- every lane decides the move of cell (ci - lane, cj - lane)
            //      kind: 0 diagonal to the previous rank, 1 another move (di rows up, dj columns left), 2 needs the generic step,
            //            3 reload the tile here, 4 no move reproduces the cell (cannot happen), 5 the walk is complete
            int kind = 3; uint32_t di = 0, dj = 0;
            uint32_t my_i = 0;                                            // TBP: the row of this lane's cell
            if (TBP) {
                // lane t looks at the cell t first-predecessor hops up and t columns left of (ci, cj): the path runs through it as long
                // as every lane before it moved diagonally to its first predecessor (SPOA's first preference), whatever that
                // predecessor's rank distance is. Rows with up to six predecessors within the tile are decided here.
                uint32_t a_ = li, nx;
                nx = __shfl_sync(FULL, J0, a_ & 31u); if (lane & 1) a_ = a_ >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : nx;
                nx = __shfl_sync(FULL, J1, a_ & 31u); if (lane & 2) a_ = a_ >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : nx;
                nx = __shfl_sync(FULL, J2, a_ & 31u); if (lane & 4) a_ = a_ >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : nx;
                nx = __shfl_sync(FULL, J3, a_ & 31u); if (lane & 8) a_ = a_ >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : nx;
                nx = __shfl_sync(FULL, J4, a_ & 31u); if (lane & 16) a_ = a_ >= (uint32_t)TB_ROWS ? (uint32_t)TB_ROWS : nx;
                if (a_ < (uint32_t)TB_ROWS && (uint32_t)lane <= cj) {
                    const uint32_t i = it - a_, j = cj - lane, b_ = lj + lane;
                    my_i = i;
                    if (i == 0 && j == 0) kind = 5;
                    else if (a_ + 1 < (uint32_t)TB_ROWS && b_ + 1 < (uint32_t)TB_COLS) {
                        const int val = cell(a_, j);
                        if (i == 0) { kind = 1; di = 0; dj = 1; }          // row 0: only horizontal moves are left
                        else {
                            const uint32_t tp = ttp[a_];
                            if (tp >> 30) kind = 2;
                            else {
                                bool far = false;
                                for (uint32_t q = tp; q; q >>= 5) far = far || a_ + (q & 31u) >= (uint32_t)TB_ROWS;
                                if (far) kind = a_ == 0 ? 2 : 3;            // a predecessor row is outside the tile
                                else {
                                    kind = 4;
                                    if (j != 0) {
                                        const int dsc = (tseq[b_] == (tm0[a_] & 3u)) ? sc.sm : sc.sx;
                                        for (uint32_t q = tp; q; q >>= 5)
                                            if (val == cell(a_ + (q & 31u), j - 1) + dsc) { kind = q == tp ? 0 : 1; di = q & 31u; dj = 1; break; }
                                    }
                                    if (kind == 4) {
                                        for (uint32_t q = tp; q; q >>= 5)
                                            if (val == cell(a_ + (q & 31u), j) + g) { kind = 1; di = q & 31u; dj = 0; break; }
                                        if (kind == 4 && j != 0 && val == cell(a_, j - 1)) { kind = 1; di = 0; dj = 1; }
                                    }
                                }
                            }
                        }
                    }
                }
            } else
            if ((uint32_t)lane <= ci && (uint32_t)lane <= cj) {
                const uint32_t i = ci - lane, j = cj - lane, a_ = li + lane, b_ = lj + lane;
                if (i == 0 && j == 0) kind = 5;
                else if (a_ + 1 < (uint32_t)TB_ROWS && b_ + 1 < (uint32_t)TB_COLS) {
                    const int val = cell(a_, j);
                    if (i == 0) { kind = 1; di = 0; dj = 1; }              // row 0: only horizontal moves are left
                    else {
                        const uint32_t m0 = tm0[a_];
                        const uint32_t npc = (m0 >> 3) & 3u;
                        if (npc == 3) kind = 2;
                        else {
                            const uint32_t np = npc == 0 ? 1u : npc;
                            const uint32_t d0 = npc == 0 ? i : meta_d0(m0), d1 = meta_d1(m0);
                            const uint32_t far = (a_ + d0 >= (uint32_t)TB_ROWS) || (np == 2 && a_ + d1 >= (uint32_t)TB_ROWS);
                            if (far) kind = a_ == 0 ? 2 : 3;                // a predecessor row is outside the tile
                            else {
                                kind = 4;
                                if (j != 0) {
                                    const int dsc = (tseq[b_] == (m0 & 3u)) ? sc.sm : sc.sx;
                                    if (val == cell(a_ + d0, j - 1) + dsc) { kind = d0 == 1 ? 0 : 1; di = d0; dj = 1; }
                                    else if (np == 2 && val == cell(a_ + d1, j - 1) + dsc) { kind = d1 == 1 ? 0 : 1; di = d1; dj = 1; }
                                }
                                if (kind == 4) {
                                    if (val == cell(a_ + d0, j) + g) { kind = 1; di = d0; dj = 0; }
                                    else if (np == 2 && val == cell(a_ + d1, j) + g) { kind = 1; di = d1; dj = 0; }
                                    else if (j != 0 && val == cell(a_, j - 1)) { kind = 1; di = 0; dj = 1; }
                                }
                            }
                        }
                    }
                }
            }
            TB_COUNT(tbc_iters);
            const unsigned fm = __ballot_sync(FULL, kind != 0);           // never 0: the tile's edge cells say "reload"
            const uint32_t run = (uint32_t)(__ffs(fm) - 1);
            const int kr = __shfl_sync(FULL, kind, run);
            const uint32_t extra = kr == 1 ? 1u : 0u;
            if ((uint64_t)n_out + run + extra > T_ncap) { bad = true; break; }
            const uint32_t row_i = TBP ? my_i : ci - lane;                 // the row of this lane's cell
            if ((uint32_t)lane < run) {
                T_aln_rank[n_out + lane] = (int32_t)(row_i - 1);
                T_aln_pos[n_out + lane] = (int32_t)(cj - lane - 1);
            } else if ((uint32_t)lane == run && extra) {
                T_aln_rank[n_out + run] = di == 0 ? -1 : (int32_t)(row_i - 1);
                T_aln_pos[n_out + run] = dj == 0 ? -1 : (int32_t)(cj - run - 1);
            }
            const uint32_t xdi = __shfl_sync(FULL, di, run), xdj = __shfl_sync(FULL, dj, run);
            n_out += run + extra;
            if (TBP) { const uint32_t after = __shfl_sync(FULL, row_i - di, (run + 31u) & 31u); if (run) ci = after; }   // where the last diagonal move of the run went
            else ci -= run;
            cj -= run;
            if (extra) { ci -= xdi; cj -= xdj; continue; }
            if (kr == 3) break;                                           // reload the tile around (ci, cj)
            if (kr == 5) continue;                                        // (0, 0) reached: the loop head ends the walk
            if (kr == 4) { bad = true; break; }
            // ---- kr == 2: one generic step on one lane (three or more predecessors, or predecessor rows outside the tile)
            TB_COUNT(tbc_generic);
            if (lane == 0) {
                const uint32_t i = ci, j = cj;
                auto getH = [&](uint32_t ii, uint32_t jj) -> int {
                    const uint32_t a_ = it - ii, b_ = jt - jj;   // ii <= it, jj <= jt always hold on a walk up/left
                    if (a_ < (uint32_t)TB_ROWS && b_ < (uint32_t)TB_COLS) return cell(a_, jj);
                    return sv.load(ii, jj);
                };
                const int val = getH(i, j);
                uint32_t pi = i, pj = j;
                bool found = false;
                if (i != 0) {
                    const uint32_t m0 = T_meta0[i - 1];
                    const uint32_t code = m0 & 3u, npc = (m0 >> 3) & 3u, d0 = meta_d0(m0), m1 = meta_d1(m0);
                    uint32_t np = npc, cs = 0;
                    if (npc == 0) np = 1;
                    if (npc == 3) { cs = T_pred_off[i - 1]; np = T_pred_off[i] - cs; }
                    if (j != 0) {
                        const int dsc = ((uint32_t)base_code(seq[j - 1]) == code) ? sc.sm : sc.sx;
                        for (uint32_t x = 0; x < np && !found; ++x) {
                            uint32_t prow = npc == 0 ? 0 : npc == 3 ? T_pred_rank[cs + x] + 1 : i - (x == 0 ? d0 : m1);
                            if (val == getH(prow, j - 1) + dsc) { pi = prow; pj = j - 1; found = true; }
                        }
                    }
                    for (uint32_t x = 0; x < np && !found; ++x) {
                        uint32_t prow = npc == 0 ? 0 : npc == 3 ? T_pred_rank[cs + x] + 1 : i - (x == 0 ? d0 : m1);
                        if (val == getH(prow, j) + g) { pi = prow; pj = j; found = true; }
                    }
                }
                if (!found && j != 0 && val == getH(i, j - 1)) { pi = i; pj = j - 1; found = true; }
                if (!found || n_out >= T_ncap) { bad = true; }
                else {
                    T_aln_rank[n_out] = (pi == i) ? -1 : (int32_t)(i - 1);
                    T_aln_pos[n_out] = (pj == j) ? -1 : (int32_t)(j - 1);
                    ++n_out;
                    ci = pi; cj = pj;
                }
            }
            ci = __shfl_sync(FULL, ci, 0); cj = __shfl_sync(FULL, cj, 0);
            n_out = __shfl_sync(FULL, n_out, 0);
            bad = __shfl_sync(FULL, (int)bad, 0) != 0;
            if (bad) break;
            if (it - ci >= (uint32_t)TB_ROWS - 1 || jt - cj >= (uint32_t)TB_COLS - 1) break;     // left the tile: reload
        }
    }
    if (lane == 0) *gv.aln_len = bad ? 0 : n_out;
#if HGPU_PHASE_CLOCKS
    if (lane == 0 && g_tb_counters) {
        atomicAdd(g_tb_counters + 0, (unsigned long long)tbc_tiles); atomicAdd(g_tb_counters + 1, (unsigned long long)tbc_iters);
        atomicAdd(g_tb_counters + 2, (unsigned long long)tbc_generic); atomicAdd(g_tb_counters + 3, (unsigned long long)n_out);
    }
#endif
    __syncwarp();
    return !bad;
}

// ---------------------------------------------------------------------------------------------------------
// Lane-parallel pieces of the graph update.
// ---------------------------------------------------------------------------------------------------------
// chain graph of the first segment: ranks = node ids
__device__ __forceinline__ void w_init_chain(GraphView& g, const uint8_t* seq, uint32_t L, int lane) {
    for (uint32_t i = lane; i < L; i += 32) {
        uint32_t c = base_code(seq[i]);
        g.code[i] = (uint8_t)c;
        g.in_head[i] = g.in_tail[i] = (i > 0) ? i - 1 : NIL;
        g.out_head[i] = (i + 1 < L) ? i : NIL;
        g.aligned[3 * i] = g.aligned[3 * i + 1] = g.aligned[3 * i + 2] = NIL;
        if (i + 1 < L) { g.e_begin[i] = i; g.e_end[i] = i + 1; g.e_w[i] = 2; g.e_next_in[i] = NIL; g.e_next_out[i] = NIL; }
        g.rank2node[i] = i; g.node2rank[i] = i;
        g.meta0[i] = meta_pack(c | ((i + 1 < L) ? 0u : META_SINK), i, i > 0 ? 1u : 0u, 1u, 0u);
        g.pred_off[i] = (i > 0) ? i - 1 : 0;
        if (i > 0) g.pred_rank[i - 1] = i - 1;
    }
    if (lane == 0) {
        g.pred_off[L] = L - 1;
        *g.n_nodes = L; *g.n_edges = L - 1; *g.aln_len = 0;
        g.sinks[0] = L - 1; *g.n_sinks = 1;
    }
    __syncwarp();
}

// per-rank DP records (meta0 / pred CSR) from rank2node/node2rank + in-lists; same content as g_build_meta
// Every access below is a dependent global load (rank -> node -> in-list -> edge -> rank of its tail), so one rank per
// lane leaves the warp waiting on ~6 round trips per 32 ranks. Each lane therefore carries four ranks (128 per
// iteration) through the same stages together: the round trips of the four overlap.
__device__ __noinline__ void w_build_meta(GraphView& g, int lane) {
    const uint32_t N = *g.n_nodes;
    // the view lives in the caller's frame: keep the arrays this function walks in registers
    auto* const L_rank2node = g.rank2node; auto* const L_in_head = g.in_head; auto* const L_code = g.code; auto* const L_out_head = g.out_head; auto* const L_e_begin = g.e_begin; auto* const L_e_next_in = g.e_next_in; auto* const L_node2rank = g.node2rank; auto* const L_pred_off = g.pred_off; auto* const L_pred_rank = g.pred_rank; auto* const L_meta0 = g.meta0; auto* const L_sinks = g.sinks;
    constexpr int K = 4;
    uint32_t running = 0, n_sinks = 0;
    for (uint32_t r0 = 0; r0 < N; r0 += 32 * K) {
        uint32_t v[K], ih[K], cd[K], oh[K], b0[K], b1[K], n1[K], n2[K], q0[K], q1[K], deg[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { const uint32_t r = r0 + k * 32 + lane; v[k] = r < N ? L_rank2node[r] : NIL; }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ih[k] = NIL; cd[k] = 0; oh[k] = NIL;
            if (v[k] != NIL) { ih[k] = L_in_head[v[k]]; cd[k] = L_code[v[k]]; oh[k] = L_out_head[v[k]]; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) { b0[k] = NIL; n1[k] = NIL; if (ih[k] != NIL) { b0[k] = L_e_begin[ih[k]]; n1[k] = L_e_next_in[ih[k]]; } }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            q0[k] = 0; b1[k] = NIL; n2[k] = NIL;
            if (b0[k] != NIL) q0[k] = L_node2rank[b0[k]];
            if (n1[k] != NIL) { b1[k] = L_e_begin[n1[k]]; n2[k] = L_e_next_in[n1[k]]; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            q1[k] = 0;
            if (b1[k] != NIL) q1[k] = L_node2rank[b1[k]];
            deg[k] = (ih[k] != NIL) + (n1[k] != NIL);
            for (uint32_t x = n2[k]; x != NIL; x = L_e_next_in[x]) ++deg[k];       // three or more in-edges: rare
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t r = r0 + k * 32 + lane;
            uint32_t incl = deg[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t off = running + incl - deg[k];
            bool is_sink = false;
            if (v[k] != NIL) {
                const uint32_t base = cd[k] | (oh[k] == NIL ? META_SINK : 0u);
                L_pred_off[r] = off;
                uint32_t d0 = 0, d1 = 0;
                if (deg[k] >= 1) { L_pred_rank[off] = q0[k]; d0 = r - q0[k]; }
                if (deg[k] >= 2) { L_pred_rank[off + 1] = q1[k]; d1 = r - q1[k]; }
                uint32_t np = 2;
                for (uint32_t x = n2[k]; x != NIL; x = L_e_next_in[x]) L_pred_rank[off + np++] = L_node2rank[L_e_begin[x]];
                L_meta0[r] = meta_pack(base, r, deg[k], d0, d1);
                is_sink = (base & META_SINK) != 0;
            }
            const unsigned sm_ = __ballot_sync(FULL, is_sink);
            if (is_sink) L_sinks[n_sinks + __popc(sm_ & ((1u << lane) - 1))] = r;
            n_sinks += __popc(sm_);
            running += __shfl_sync(FULL, incl, 31);
        }
    }
    if (lane == 0) { L_pred_off[N] = running; *g.n_sinks = n_sinks; }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// w_add_alignment: SPOA Graph::add_alignment (unit weights), lane-parallel.
// A global (kNW) alignment consumes every sequence position exactly once, so the work is indexed by sequence
// position j: (A) resolve the node j lands on (existing node with the same base, an aligned sibling with that
// base, or a new node), (B) number the new nodes in alignment order (= SPOA's creation order) with a warp scan,
// initialise them and cross-link aligned sets, (C) add or bump the edge (node[j-1] -> node[j]). A path visits a
// node at most once and at most one member of an aligned set, so lanes never touch the same list; an in-list
// receives at most one new edge per alignment, so in-edge ORDER (what fixes every tie-break downstream) is the
// same as in the serial version; edge ids differ, which nothing observes.
// Returns ST_OK, or ST_CAPACITY; 0xFFFFFFFF = "not a full-coverage alignment, use the serial version".
// ---------------------------------------------------------------------------------------------------------
static constexpr uint32_t ABSENT = 0xFFFFFFFEu;

__device__ __noinline__ uint32_t w_add_alignment(GraphView& g, GraphScratch& s, const uint8_t* seq, uint32_t L, int lane) {
    const uint32_t n = *g.aln_len, N0 = *g.n_nodes, E0 = *g.n_edges;
    // the view lives in the caller's frame: keep the arrays this function walks in registers
    auto* const L_aln_pos = g.aln_pos; auto* const L_aln_rank = g.aln_rank; auto* const L_rank2node = g.rank2node; auto* const L_code = g.code; auto* const L_aligned = g.aligned; auto* const L_in_head = g.in_head; auto* const L_in_tail = g.in_tail; auto* const L_out_head = g.out_head; auto* const L_e_begin = g.e_begin; auto* const L_e_end = g.e_end; auto* const L_e_w = g.e_w; auto* const L_e_next_in = g.e_next_in; auto* const L_e_next_out = g.e_next_out;
    if (n == 0 || 3ull * L > s.stack_cap) return 0xFFFFFFFFu;
    if ((uint64_t)N0 + L > g.ncap || (uint64_t)E0 + L + 1 > g.ecap) return ST_CAPACITY;
    uint32_t* const by_j = s.stack; uint32_t* const nid = by_j + L; uint32_t* const alto = nid + L;
    for (uint32_t j = lane; j < L; j += 32) by_j[j] = ABSENT;
    __syncwarp();
    for (uint32_t t = lane; t < n; t += 32) {
        const int32_t pos = L_aln_pos[t];
        if (pos != -1) by_j[pos] = (uint32_t)L_aln_rank[t];     // rank, or 0xFFFFFFFF for an insertion
    }
    __syncwarp();
    // (A)+(B): node of every position; new nodes numbered in alignment order. Four positions per lane (j0 + 32k + lane)
    // go through the dependent loads (rank -> node -> base / aligned triple -> bases of the aligned nodes) together.
    uint32_t n_new = 0;
    bool full = true;
    constexpr int K = 4;
    for (uint32_t j0 = 0; j0 < L; j0 += 32 * K) {
        uint32_t rk[K], c[K], an[K], o0[K], o1[K], o2[K], ca[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t j = j0 + k * 32 + lane;
            rk[k] = ABSENT; c[k] = 0;
            if (j < L) { rk[k] = by_j[j]; c[k] = base_code(seq[j]); if (rk[k] == ABSENT) full = false; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) an[k] = (rk[k] < 0xFFFFFFFEu) ? L_rank2node[rk[k]] : NIL;      // neither ABSENT nor an insertion
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ca[k] = 0; o0[k] = NIL; o1[k] = NIL; o2[k] = NIL;
            if (an[k] != NIL) { ca[k] = L_code[an[k]]; o0[k] = L_aligned[3 * an[k]]; o1[k] = L_aligned[3 * an[k] + 1]; o2[k] = L_aligned[3 * an[k] + 2]; }
        }
        uint32_t c0[K], c1[K], c2[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (o0[k] == NIL) { o1[k] = NIL; o2[k] = NIL; } else if (o1[k] == NIL) o2[k] = NIL;
            const bool look = an[k] != NIL && ca[k] != c[k];                                            // the aligned nodes matter only on a mismatch
            c0[k] = (look && o0[k] != NIL) ? L_code[o0[k]] : 4u;
            c1[k] = (look && o1[k] != NIL) ? L_code[o1[k]] : 4u;
            c2[k] = (look && o2[k] != NIL) ? L_code[o2[k]] : 4u;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t j = j0 + k * 32 + lane;
            bool isnew = false; uint32_t node = NIL, al = NIL;
            if (j < L && rk[k] != ABSENT) {
                if (rk[k] == 0xFFFFFFFFu) isnew = true;
                else if (ca[k] == c[k]) node = an[k];
                else if (c0[k] == c[k]) node = o0[k];
                else if (c1[k] == c[k]) node = o1[k];
                else if (c2[k] == c[k]) node = o2[k];
                else { isnew = true; al = an[k]; }
            }
            const unsigned m = __ballot_sync(FULL, isnew);
            if (isnew) node = N0 + n_new + __popc(m & ((1u << lane) - 1));
            n_new += __popc(m);
            if (j < L) { nid[j] = node; alto[j] = al; }
        }
    }
    if (__any_sync(FULL, !full)) return 0xFFFFFFFFu;             // nothing modified yet
    __syncwarp();
    // (B2): initialise new nodes, cross-link aligned sets (SPOA order: existing members first, then the node aligned to)
    for (uint32_t j = lane; j < L; j += 32) {
        const uint32_t v = nid[j];
        if (v < N0) continue;
        L_code[v] = (uint8_t)base_code(seq[j]);
        L_in_head[v] = NIL; L_in_tail[v] = NIL; L_out_head[v] = NIL;
        uint32_t a3[3] = {NIL, NIL, NIL};
        const uint32_t a = alto[j];
        if (a != NIL) {
            int cnt = 0;
            for (int q = 0; q < 3; ++q) {
                const uint32_t o = L_aligned[3 * a + q];
                if (o == NIL) break;
                a3[cnt++] = o;
                for (int z = 0; z < 3; ++z) if (L_aligned[3 * o + z] == NIL) { L_aligned[3 * o + z] = v; break; }
            }
            a3[cnt] = a;
            for (int z = 0; z < 3; ++z) if (L_aligned[3 * a + z] == NIL) { L_aligned[3 * a + z] = v; break; }
        }
        L_aligned[3 * v] = a3[0]; L_aligned[3 * v + 1] = a3[1]; L_aligned[3 * v + 2] = a3[2];
    }
    __syncwarp();
    // (C): edges between consecutive positions, weight 2 (both bases contribute 1); four positions per lane again
    uint32_t e_new = 0;
    for (uint32_t j0 = 1; j0 < L; j0 += 32 * K) {
        uint32_t b[K], e[K], x[K], xe[K], xn[K], tl[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t j = j0 + k * 32 + lane;
            b[k] = NIL; e[k] = NIL;
            if (j < L) { b[k] = nid[j - 1]; e[k] = nid[j]; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) { x[k] = NIL; tl[k] = NIL; if (b[k] != NIL) { x[k] = L_out_head[b[k]]; tl[k] = L_in_tail[e[k]]; } }
        const uint32_t oh0 = x[0], oh1 = x[1], oh2 = x[2], oh3 = x[3];
        const uint32_t oh[K] = {oh0, oh1, oh2, oh3};
        bool hit[K] = {false, false, false, false};
        // walk the out-lists of the four tails in step: first two edges with overlapped loads, longer lists one by one
#pragma unroll
        for (int step = 0; step < 2; ++step) {
#pragma unroll
            for (int k = 0; k < K; ++k) { xe[k] = NIL; xn[k] = NIL; if (x[k] != NIL && !hit[k]) { xe[k] = L_e_end[x[k]]; xn[k] = L_e_next_out[x[k]]; } }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (x[k] != NIL && !hit[k]) {
                    if (xe[k] == e[k]) hit[k] = true; else x[k] = xn[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            while (x[k] != NIL && !hit[k]) { if (L_e_end[x[k]] == e[k]) hit[k] = true; else x[k] = L_e_next_out[x[k]]; }
            if (hit[k]) L_e_w[x[k]] += 2;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool need = b[k] != NIL && !hit[k];
            const unsigned m = __ballot_sync(FULL, need);
            if (need) {
                const uint32_t nx = E0 + e_new + __popc(m & ((1u << lane) - 1));
                L_e_begin[nx] = b[k]; L_e_end[nx] = e[k]; L_e_w[nx] = 2;
                L_e_next_in[nx] = NIL;
                L_e_next_out[nx] = oh[k];
                L_out_head[b[k]] = nx;
                if (tl[k] == NIL) L_in_head[e[k]] = nx; else L_e_next_in[tl[k]] = nx;
                L_in_tail[e[k]] = nx;
            }
            e_new += __popc(m);
        }
    }
    if (lane == 0) { *g.n_nodes = N0 + n_new; *g.n_edges = E0 + e_new; }
    __syncwarp();
    return ST_OK;
}

// ---------------------------------------------------------------------------------------------------------
// w_toposort: SPOA Graph::topological_sort, same order, mostly lane-parallel.
// SPOA walks roots i = 0..N-1 in id order and runs a DFS over in-edges / aligned nodes from every unmarked one.
// A root whose predecessors are all emitted and that has no aligned nodes is emitted on the spot — that is the
// bulk of a POA graph, and it is decided here for 32 consecutive ids at a time from registers and two shared-memory
// bitmaps (emitted, "do not check aligned"): the longest prefix of the batch whose nodes are emitted already or
// emit-on-the-spot is ranked with one ballot. Only the first node that breaks the prefix (an aligned set, a
// branch whose nodes have larger ids) runs SPOA's DFS verbatim on one lane, then the batch resumes behind it.
// Returns 1 ok, 0 failed (stack overflow / step guard): the caller falls back to the serial g_toposort.
// ---------------------------------------------------------------------------------------------------------
static constexpr uint32_t TOPO_BM_WORDS = HGPU_RING ? 576 : 400;     // nodes per bitmap = 32x
static constexpr uint32_t TOPO_STACK = (DP_SMEM_PER_WARP - 2 * TOPO_BM_WORDS * 4) / 4;

// Per-node record of the sort: everything a DFS visit needs in two 16-byte loads issued together, instead of the chain
// in_head -> e_begin / e_next_in -> ... -> aligned (4-6 dependent round trips per node, which is what the serial part of
// the sort spent its time on: 24-28 % of the warp-cycles of a deep edge, profiles/r2c_*).
//   p[0..3] = tails of the first four in-edges in in-edge order (NIL-padded), a[0..2] = aligned nodes, more = id of the fifth in-edge
struct __align__(16) TopoRec { uint32_t p[4]; uint32_t a[3]; uint32_t more; };
static_assert(sizeof(TopoRec) == 32, "record layout");

// all records, lane-parallel, four nodes per lane in flight (the chain of an in-list is still dependent loads, but 128 of them overlap)
__device__ __noinline__ void w_build_trec(const GraphView& g, TopoRec* rec, int lane) {
    const uint32_t N = *g.n_nodes;
    // the view lives in the caller's frame: keep the arrays this function walks in registers
    auto* const L_in_head = g.in_head; auto* const L_aligned = g.aligned; auto* const L_e_begin = g.e_begin; auto* const L_e_next_in = g.e_next_in;
    constexpr int K = 4;
    for (uint32_t v0 = 0; v0 < N; v0 += 32 * K) {
        uint32_t x[K];
        TopoRec r[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t v = v0 + k * 32 + lane;
            x[k] = NIL;
#pragma unroll
            for (int q = 0; q < 4; ++q) r[k].p[q] = NIL;
            r[k].a[0] = r[k].a[1] = r[k].a[2] = NIL; r[k].more = NIL;
            if (v < N) { x[k] = L_in_head[v]; r[k].a[0] = L_aligned[3 * v]; r[k].a[1] = L_aligned[3 * v + 1]; r[k].a[2] = L_aligned[3 * v + 2]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (x[k] != NIL) { r[k].p[q] = L_e_begin[x[k]]; x[k] = L_e_next_in[x[k]]; }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t v = v0 + k * 32 + lane;
            if (v < N) {
                if (r[k].a[0] == NIL) { r[k].a[1] = NIL; r[k].a[2] = NIL; } else if (r[k].a[1] == NIL) r[k].a[2] = NIL;
                r[k].more = x[k];
                uint4* dst = reinterpret_cast<uint4*>(rec + v);
                dst[0] = make_uint4(r[k].p[0], r[k].p[1], r[k].p[2], r[k].p[3]);
                dst[1] = make_uint4(r[k].a[0], r[k].a[1], r[k].a[2], r[k].more);
            }
        }
    }
    __syncwarp();
}

__device__ __noinline__ int w_toposort(GraphView& g, const TopoRec* rec, uint8_t* wsm, int lane) {
    const uint32_t N = *g.n_nodes;
    if (N > TOPO_BM_WORDS * 32) return 0;
    // the view lives in the caller's frame (local memory): keep what the loops store through in registers
    uint32_t* const r2n = g.rank2node; uint32_t* const n2r = g.node2rank;
    const uint32_t* const e_begin = g.e_begin; const uint32_t* const e_next_in = g.e_next_in;
    const uint32_t step_limit = 16u * (N + *g.n_edges) + 1024u;
    uint32_t* perm = reinterpret_cast<uint32_t*>(wsm);
    uint32_t* nochk = perm + TOPO_BM_WORDS;
    uint32_t* stk = nochk + TOPO_BM_WORDS;
    for (uint32_t w = lane; w < (N + 31) / 32; w += 32) { perm[w] = 0; nochk[w] = 0; }
    __syncwarp();
    auto is_perm = [&](uint32_t v) -> bool { return (perm[v >> 5] >> (v & 31)) & 1u; };
    const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
    uint32_t nr = 0;
    int okflag = 1;
#if HGPU_PHASE_CLOCKS
    uint32_t tpc_roots = 0, tpc_visits = 0, tpc_passes = 0;
#endif
    for (uint32_t i0 = 0; i0 < N && okflag; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < N;
        uint32_t np = 0, p0 = NIL, p1 = NIL;
        bool slow = false;                                        // aligned nodes or more than two predecessors
        uint4 ra = make_uint4(NIL, NIL, NIL, NIL), rb = make_uint4(NIL, NIL, NIL, NIL);
        if (valid) {
            ra = rec4[2 * i]; rb = rec4[2 * i + 1];
            p0 = ra.x; p1 = ra.y;
            np = (p0 != NIL) + (p1 != NIL);
            slow = ra.z != NIL || rb.x != NIL;
            // the DFS below runs on one lane and waits for every record it touches: pull the records of this root's
            // predecessors / aligned nodes with larger ids (the ones a walk from this root may still have to visit) into L1 now,
            // 32 roots at a time
            const uint32_t ch[7] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z};
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (ch[q] != NIL && ch[q] > i) asm volatile("prefetch.global.L1 [%0];" :: "l"(rec4 + 2 * ch[q]));
        }
        if (i + 32 < N) asm volatile("prefetch.global.L1 [%0];" :: "l"(rec4 + 2 * (i + 32)));   // the next batch's roots
        uint32_t pos = 0;
        while (true) {
            TB_COUNT(tpc_passes);
            const uint32_t lo = i0 + pos;
            const bool marked = valid && is_perm(i);
            bool ok = valid && (uint32_t)lane >= pos && !marked && !slow;
            if (ok && np >= 1) ok = is_perm(p0) || (p0 >= lo && p0 < i);
            if (ok && np >= 2) ok = is_perm(p1) || (p1 >= lo && p1 < i);
            const bool pass = (uint32_t)lane < pos || !valid || marked || ok;
            const unsigned failmask = __ballot_sync(FULL, !pass);
            const uint32_t f = failmask ? (uint32_t)(__ffs(failmask) - 1) : 32u;
            const bool emit = ok && (uint32_t)lane < f;
            const unsigned em = __ballot_sync(FULL, emit);
            if (emit) {
                const uint32_t r = nr + __popc(em & ((1u << lane) - 1));
                r2n[r] = i; n2r[i] = r;
            }
            if (lane == 0 && em) perm[i0 >> 5] |= em;
            nr += __popc(em);
            __syncwarp();
            if (f >= 32) break;
            // SPOA's DFS from root i0+f, same visits and the same emission order as the serial code, but every visit is done by the
            // whole warp: lanes 0-7 read the eight words of the node's record in one 32-byte access, lanes 0-6 test "their"
            // predecessor / aligned node against the emitted bitmap at once, one ballot gives the children still to visit, they
            // are pushed in SPOA's order (in-edges first, then aligned nodes) by the lanes that hold them, and their records are
            // pulled into L1 on the spot - so the walk waits for one shared-memory round trip per visit instead of a chain of
            // dependent global loads. A node that pushed children is flagged on the stack: when the walk returns to it every
            // child is emitted (a child is only popped once permanent), so it is finalised without looking at its lists again.
            {
                constexpr uint32_t EXPANDED = 0x80000000u, IDMASK = 0x3FFFFFFFu;
                const uint32_t* recw = reinterpret_cast<const uint32_t*>(rec);
                uint32_t sp = 1, guard = 0;
                const uint32_t limit = step_limit;
                uint32_t top = i0 + f;                                                // the top of the stack stays in a register (uniform)
                if (lane == 0) stk[0] = top;
                __syncwarp();
                TB_COUNT(tpc_roots);
                while (sp > 0) {
                    TB_COUNT(tpc_visits);
                    if (++guard > limit) { okflag = 0; break; }
                    const uint32_t v = top & IDMASK;
                    // everything this visit may need is requested at once: the node's two bitmap words, its record, the entry below it
                    const uint32_t pw = perm[v >> 5], nw = nochk[v >> 5];
                    const uint32_t child = lane < 8 ? recw[8 * (size_t)v + lane] : NIL;   // words 0-3 in-edge tails, 4-6 aligned nodes, 7 fifth in-edge
                    const uint32_t below = sp >= 2 ? stk[sp - 2] : 0u;
                    if ((pw >> (v & 31)) & 1u) { --sp; top = below; continue; }
                    const bool chk = !((nw >> (v & 31)) & 1u);
                    if (!(top & EXPANDED)) {
                        const uint32_t more = __shfl_sync(FULL, child, 7);
                        if (more == NIL) {
                            const bool mine = lane < 4 || (lane < 7 && chk);
                            const bool unem = mine && child != NIL && !is_perm(child);
                            const unsigned m = __ballot_sync(FULL, unem);
                            if (m) {
                                const uint32_t n = (uint32_t)__popc(m);
                                if (sp + n > TOPO_STACK) { okflag = 0; break; }
                                if (unem) {
                                    stk[sp + __popc(m & ((1u << lane) - 1u))] = child;
                                    if (lane >= 4) atomicOr(&nochk[child >> 5], 1u << (child & 31));
                                    asm volatile("prefetch.global.L1 [%0];" :: "l"(rec4 + 2 * (size_t)child));
                                }
                                if (lane == 0) stk[sp - 1] = v | EXPANDED;
                                top = __shfl_sync(FULL, child, 31 - __clz(m));        // the last child pushed is visited first
                                sp += n;
                                __syncwarp();
                                continue;
                            }
                        } else {
                            // five or more in-edges (2 % of the nodes of a deep graph): the in-list beyond the record on one lane
                            uint32_t nsp = sp;
                            if (lane == 0) {
                                const uint4 ra = rec4[2 * (size_t)v], rb = rec4[2 * (size_t)v + 1];
                                const uint32_t pp[4] = {ra.x, ra.y, ra.z, ra.w};
                                for (int q = 0; q < 4 && okflag; ++q)
                                    if (!is_perm(pp[q])) { if (nsp >= TOPO_STACK) okflag = 0; else stk[nsp++] = pp[q]; }
                                for (uint32_t x = rb.w; x != NIL && okflag; ) {
                                    const uint32_t b = e_begin[x], nx = e_next_in[x];
                                    if (!is_perm(b)) { if (nsp >= TOPO_STACK) okflag = 0; else stk[nsp++] = b; }
                                    x = nx;
                                }
                                if (chk) {
                                    const uint32_t al[3] = {rb.x, rb.y, rb.z};
                                    for (int q = 0; q < 3 && okflag; ++q) {
                                        if (al[q] == NIL) break;
                                        if (!is_perm(al[q])) { if (nsp >= TOPO_STACK) okflag = 0; else { stk[nsp++] = al[q]; nochk[al[q] >> 5] |= 1u << (al[q] & 31); } }
                                    }
                                }
                                if (nsp != sp) stk[sp - 1] = v | EXPANDED;
                            }
                            nsp = __shfl_sync(FULL, nsp, 0);
                            okflag = __shfl_sync(FULL, okflag, 0);
                            if (!okflag) break;
                            __syncwarp();
                            if (nsp != sp) { sp = nsp; top = stk[sp - 1]; continue; }
                        }
                    }
                    // every child is emitted: the node (and, if it leads its aligned set, the set) is final
                    if (lane == 0) perm[v >> 5] = pw | (1u << (v & 31));                  // only this walk writes the bitmap: pw is current
                    if (chk) {
                        const bool al = lane >= 4 && lane < 7 && child != NIL;           // NIL-terminated: the aligned nodes sit in lanes 4, 5, 6 in order
                        const uint32_t na = (uint32_t)__popc(__ballot_sync(FULL, al));
                        if (lane == 0) { r2n[nr] = v; n2r[v] = nr; }
                        if (al) { r2n[nr + 1 + (lane - 4)] = child; n2r[child] = nr + 1 + (lane - 4); }
                        nr += 1 + na;
                    }
                    --sp; top = below;
                    __syncwarp();
                }
            }
            __syncwarp();
            if (!okflag) break;
            pos = f + 1;
            if (pos >= 32) break;
        }
    }
#if HGPU_PHASE_CLOCKS
    if (lane == 0 && g_tb_counters) {
        atomicAdd(g_tb_counters + 4, (unsigned long long)N); atomicAdd(g_tb_counters + 5, (unsigned long long)tpc_passes);
        atomicAdd(g_tb_counters + 6, (unsigned long long)tpc_roots); atomicAdd(g_tb_counters + 7, (unsigned long long)tpc_visits);
    }
#endif
    if (okflag && nr != N) okflag = 0;
    return okflag;
}

// The same sort walking the in-lists directly (no records), kept for A/B runs (-DHGPU_SHALLOW_TREC=0). With the serial one-lane DFS
// the record pass cost the shallow kernel 4 % (1,394 vs 1,453 GCUPS on config 3); with the warp-cooperative walk it gains 3 %
// (1,506 vs 1,461), so every kernel runs w_toposort now.
__device__ __noinline__ int w_toposort_chain(GraphView& g, uint8_t* wsm, int lane) {
    const uint32_t N = *g.n_nodes;
    if (N > TOPO_BM_WORDS * 32) return 0;
    uint32_t* perm = reinterpret_cast<uint32_t*>(wsm);
    uint32_t* nochk = perm + TOPO_BM_WORDS;
    uint32_t* stk = nochk + TOPO_BM_WORDS;
    for (uint32_t w = lane; w < (N + 31) / 32; w += 32) { perm[w] = 0; nochk[w] = 0; }
    __syncwarp();
    auto is_perm = [&](uint32_t v) -> bool { return (perm[v >> 5] >> (v & 31)) & 1u; };
    uint32_t nr = 0;
    int okflag = 1;
    for (uint32_t i0 = 0; i0 < N && okflag; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < N;
        uint32_t np = 0, p0 = NIL, p1 = NIL;
        bool slow = false;                                        // aligned nodes or more than two predecessors
        if (valid) {
            uint32_t x = g.in_head[i];
            if (x != NIL) {
                p0 = g.e_begin[x]; np = 1; x = g.e_next_in[x];
                if (x != NIL) { p1 = g.e_begin[x]; np = 2; if (g.e_next_in[x] != NIL) slow = true; }
            }
            if (g.aligned[3 * i] != NIL) slow = true;
        }
        uint32_t pos = 0;
        while (true) {
            const uint32_t lo = i0 + pos;
            const bool marked = valid && is_perm(i);
            bool ok = valid && (uint32_t)lane >= pos && !marked && !slow;
            if (ok && np >= 1) ok = is_perm(p0) || (p0 >= lo && p0 < i);
            if (ok && np >= 2) ok = is_perm(p1) || (p1 >= lo && p1 < i);
            const bool pass = (uint32_t)lane < pos || !valid || marked || ok;
            const unsigned failmask = __ballot_sync(FULL, !pass);
            const uint32_t f = failmask ? (uint32_t)(__ffs(failmask) - 1) : 32u;
            const bool emit = ok && (uint32_t)lane < f;
            const unsigned em = __ballot_sync(FULL, emit);
            if (emit) {
                const uint32_t r = nr + __popc(em & ((1u << lane) - 1));
                g.rank2node[r] = i; g.node2rank[i] = r;
            }
            if (lane == 0 && em) perm[i0 >> 5] |= em;
            nr += __popc(em);
            __syncwarp();
            if (f >= 32) break;
            // SPOA's DFS from root i0+f on one lane: same visits and the same emission order as the serial code, with
            // two changes that only remove memory round trips: (1) the loads of a visit are issued before their first
            // use (node record and aligned triple together, both words of an edge together); (2) a node that pushed
            // children is flagged on the stack - when the walk returns to it every child is emitted (a child is only
            // popped once permanent), so it is finalised on the spot instead of walking its lists a second time.
            if (lane == 0) {
                constexpr uint32_t EXPANDED = 0x80000000u, HAS_ALIGNED = 0x40000000u, IDMASK = 0x3FFFFFFFu;
                uint32_t sp = 0, guard = 0;
                const uint32_t limit = 16u * (N + *g.n_edges) + 1024u;
                stk[sp++] = i0 + f;
                while (sp > 0) {
                    if (++guard > limit) { okflag = 0; break; }
                    const uint32_t self = sp - 1;
                    const uint32_t top = stk[self];
                    const uint32_t v = top & IDMASK;
                    if (is_perm(v)) { --sp; continue; }
                    const bool chk = !((nochk[v >> 5] >> (v & 31)) & 1u);
                    uint32_t al[3] = {NIL, NIL, NIL};
                    bool vvalid = true;
                    if (top & EXPANDED) {
                        if (chk && (top & HAS_ALIGNED)) { al[0] = g.aligned[3 * v]; al[1] = g.aligned[3 * v + 1]; al[2] = g.aligned[3 * v + 2]; }
                    } else {
                        uint32_t x = g.in_head[v];
                        const uint32_t a0 = g.aligned[3 * v], a1 = g.aligned[3 * v + 1], a2 = g.aligned[3 * v + 2];
                        while (x != NIL) {
                            const uint32_t b = g.e_begin[x], nx = g.e_next_in[x];
                            if (!is_perm(b)) {
                                if (sp >= TOPO_STACK) { okflag = 0; break; }
                                stk[sp++] = b; vvalid = false;
                            }
                            x = nx;
                        }
                        if (!okflag) break;
                        if (chk) {
                            al[0] = a0; al[1] = a0 == NIL ? NIL : a1; al[2] = (a0 == NIL || a1 == NIL) ? NIL : a2;
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const uint32_t o = al[q];
                                if (o == NIL) break;
                                if (!is_perm(o)) {
                                    if (sp >= TOPO_STACK) { okflag = 0; break; }
                                    stk[sp++] = o; nochk[o >> 5] |= 1u << (o & 31); vvalid = false;
                                }
                            }
                            if (!okflag) break;
                        }
                        if (!vvalid) stk[self] = v | EXPANDED | (a0 != NIL ? HAS_ALIGNED : 0u);   // children first; finalised on return
                    }
                    if (vvalid) {
                        perm[v >> 5] |= 1u << (v & 31);
                        if (chk) {
                            g.rank2node[nr] = v; g.node2rank[v] = nr; ++nr;
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                if (al[q] == NIL) break;
                                g.rank2node[nr] = al[q]; g.node2rank[al[q]] = nr; ++nr;
                            }
                        }
                        --sp;
                    }
                }
            }
            nr = __shfl_sync(FULL, nr, 0);
            okflag = __shfl_sync(FULL, okflag, 0);
            __syncwarp();
            if (!okflag) break;
            pos = f + 1;
            if (pos >= 32) break;
        }
    }
    if (okflag && nr != N) okflag = 0;
    return okflag;
}

}  // namespace hgpu
#include "poa_topo_claims.cuh"
namespace hgpu {

// ---------------------------------------------------------------------------------------------------------
// w_consensus_scores: the rank-order pass of SPOA's heaviest bundle (g_consensus_scores), 32 ranks at a time.
// The serial version is a chain of dependent global loads per node (rank -> node -> in-list -> edge -> score of the
// tail node): ~5 memory round trips per node on one lane. Here every lane loads the in-list of one rank of the batch
// (all round trips overlap), then the 32 nodes are resolved in rank order with the scores of in-batch predecessors
// passed through shuffles. Scores are kept as int32 (path weight <= 2 * reads * nodes) and stored as SPOA's int64.
// Same score[] / pred[] arrays and the same maximal node as the serial code; branch completion and the backtrack stay
// on one lane (g_consensus_finish).
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t w_consensus_scores(GraphView& g, GraphScratch& s, int lane) {
    const uint32_t N = *g.n_nodes;
    int64_t* const S_score = s.score; int32_t* const S_pred = s.pred; uint32_t* const S_stack = s.stack;
    // the view lives in the caller's frame: keep the arrays this function walks in registers
    auto* const L_rank2node = g.rank2node; auto* const L_in_head = g.in_head; auto* const L_e_begin = g.e_begin; auto* const L_e_w = g.e_w; auto* const L_e_next_in = g.e_next_in; auto* const L_node2rank = g.node2rank;
    constexpr int K = 4;                                                  // in-edges held in registers; longer lists walk memory
    uint32_t max_id = 0; int max_sc = -1;
    for (uint32_t r0 = 0; r0 < N; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool valid = r < N;
        uint32_t u = 0, eb[K], ew[K], rk[K]; int sb[K];
        int deg = 0; bool more = false;
        if (valid) {
            u = L_rank2node[r];
            uint32_t x = L_in_head[u];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                eb[k] = 0; ew[k] = 0; rk[k] = 0; sb[k] = -1;
                if (x != NIL) { eb[k] = L_e_begin[x]; ew[k] = L_e_w[x]; x = L_e_next_in[x]; deg = k + 1; }
            }
            more = x != NIL;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < deg) {
                    rk[k] = L_node2rank[eb[k]];
                    if (rk[k] < r0) sb[k] = (int)S_score[eb[k]];         // resolved in an earlier batch
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) { eb[k] = 0; ew[k] = 0; rk[k] = 0; sb[k] = -1; }
        }
        const int maxdeg = __reduce_max_sync(FULL, deg);
        int myscore = -1; int32_t mypred = -1;
        const int nb = (N - r0) < 32u ? (int)(N - r0) : 32;
        for (int q = 0; q < nb; ++q) {
            int v[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = (k < maxdeg) ? __shfl_sync(FULL, myscore, (int)((rk[k] - r0) & 31u)) : -1;
            if (lane == q) {
                int cur = -1, ps = -1; int32_t cp = -1;
                if (!more) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (k < deg) {
                            const int sbk = rk[k] >= r0 ? v[k] : sb[k];
                            const int w = (int)ew[k];
                            if (cur < w || (cur == w && ps <= sbk)) { cur = w; cp = (int32_t)eb[k]; ps = sbk; }
                        }
                    }
                } else {                                                  // long in-list: scores of earlier ranks are in memory by now
                    for (uint32_t y = L_in_head[u]; y != NIL; y = L_e_next_in[y]) {
                        const uint32_t b = L_e_begin[y];
                        const int w = (int)L_e_w[y];
                        const int sbk = (int)S_score[b];
                        if (cur < w || (cur == w && ps <= sbk)) { cur = w; cp = (int32_t)b; ps = sbk; }
                    }
                }
                if (cp != -1) cur += ps;
                myscore = cur; mypred = cp;
                S_score[u] = (int64_t)cur; S_pred[u] = cp;
            }
            __syncwarp();
        }
        if (valid) S_stack[r] = mypred == -1 ? NIL : L_node2rank[mypred];   // the chosen predecessor by rank, for the backtrack
        // first strict maximum in rank order
        const int bm = __reduce_max_sync(FULL, valid ? myscore : INT32_MIN);
        if (bm > max_sc) {
            const int src = __ffs(__ballot_sync(FULL, valid && myscore == bm)) - 1;
            max_id = __shfl_sync(FULL, u, src);
            max_sc = bm;
        }
    }
    return max_id;
}

// Backtrack of the consensus path from node max_id (a sink, so no branch completion ran) along the predecessor ranks
// that w_consensus_scores left in s.stack. A path runs down the ranks, mostly to the previous one: 32 ranks are
// loaded at a time and the hops inside the batch go through shuffles instead of dependent global loads. Node ids
// land in `out` in path order; returns the path length.
__device__ __noinline__ uint32_t w_consensus_backtrack(GraphView& g, GraphScratch& s, uint32_t max_id, uint32_t* out, int lane) {
    uint32_t cur = g.node2rank[max_id], n = 0;
    while (cur != NIL) {
        const uint32_t rb = cur & ~31u;
        uint32_t pr = NIL, nd = 0;
        if (rb + lane <= cur) { pr = s.stack[rb + lane]; nd = g.rank2node[rb + lane]; }
        uint32_t vis = 0, pos = cur;
        while (pos != NIL && pos >= rb) {
            vis |= 1u << (pos - rb);
            pos = __shfl_sync(FULL, pr, (int)(pos - rb));
        }
        if ((vis >> lane) & 1u) out[n + __popc(vis >> lane) - 1] = nd;   // descending ranks for now
        n += __popc(vis);
        cur = pos;
    }
    __syncwarp();
    for (uint32_t a = lane; a < n / 2; a += 32) { const uint32_t t = out[a]; out[a] = out[n - 1 - a]; out[n - 1 - a] = t; }
    __syncwarp();
    return n;
}

// ---------------------------------------------------------------------------------------------------------
// k_poa_edges: the persistent per-edge kernel.
// ---------------------------------------------------------------------------------------------------------
#ifndef HGPU_MINBLOCKS
#define HGPU_MINBLOCKS 7
#endif
#ifndef HGPU_PHASE_CLOCKS
#define HGPU_PHASE_CLOCKS 0
#endif
#ifndef HGPU_SHALLOW_TBP
#define HGPU_SHALLOW_TBP 0   // A/B: the warp-per-edge kernel's traceback with predecessor words too
#endif
#if HGPU_PHASE_CLOCKS
#define PHASE_CLK_DECL long long pc_t = clock64(); unsigned long long pc_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PHASE_CLK(idx) { const long long pc_n = clock64(); pc_acc[idx] += (unsigned long long)(pc_n - pc_t); pc_t = pc_n; }
#define PHASE_CLK_FLUSH(a) if (lane == 0 && (a).phase_clk) { for (int pc_i = 0; pc_i < 10; ++pc_i) atomicAdd((a).phase_clk + pc_i, pc_acc[pc_i]); }
#else
#define PHASE_CLK_DECL
#define PHASE_CLK(idx)
#define PHASE_CLK_FLUSH(a)
#endif
enum : int { PC_QUEUE = 0, PC_INIT = 1, PC_FILL = 2, PC_TRACEBACK = 3, PC_ADD = 4, PC_TOPO = 5, PC_META = 6, PC_CONSENSUS = 7, PC_PUBLISH = 8 };

__device__ __forceinline__ int lane_id() {       // read once, never rematerialised from S2R inside the hot loops
    const int l = (int)(threadIdx.x & 31u);
    return __shfl_sync(FULL, l, l);                 // a shuffle result is opaque to ptxas, S2R is not
}

template <int RING>
__device__ __forceinline__ void poa_edges_body(const PoaArgs& a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * DP_WARPS_PER_BLOCK + wib;
    uint8_t* wsm = smem_raw + (size_t)wib * (RING > 2 ? DP_SMEM_PER_WARP_DEEP : DP_SMEM_PER_WARP);
    uint8_t* slot = a.arena + (uint64_t)gw * a.slot_bytes;
    uint8_t* wsb = a.ws + (uint64_t)gw * a.wl.bytes;
    GraphView gv = bind_graph(wsb, a.wl);
    GraphScratch gs = bind_scratch(wsb, a.wl);
    uint32_t* hdr = reinterpret_cast<uint32_t*>(wsb + a.wl.o_hdr);
    uint32_t* plan = reinterpret_cast<uint32_t*>(wsb + a.wl.o_plan);
    uint32_t* tbp = reinterpret_cast<uint32_t*>(wsb + a.wl.o_tbp);
    TopoRec* trec = reinterpret_cast<TopoRec*>(wsb + a.wl.o_trec);
    unsigned long long st_cells = 0, st_padded = 0, st_aln = 0, st_aln32 = 0, st_bases = 0;
    PHASE_CLK_DECL

    while (true) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= a.n_items) break;
        const uint32_t e = a.items[item];
        PHASE_CLK(PC_QUEUE)
        const uint32_t s0 = a.e_seg_off[e];
        const uint32_t R = a.e_seg_off[e + 1] - s0;
        uint32_t st = ST_OK;
        uint32_t n_cons = 0;
        bool debug_stop = false;
        unsigned long long e_cells = 0, e_padded = 0, e_aln = 0, e_aln32 = 0, e_bases = 0;   // counted only if the edge completes
        if (R == 0) {
            if (lane == 0) { *gv.n_nodes = 0; *gv.n_edges = 0; *gv.aln_len = 0; }
        } else {
            const uint32_t L0 = a.seg_len[s0];
            if (L0 > gv.ncap || L0 > gv.ecap) st = ST_CAPACITY;
            else { w_init_chain(gv, a.bases + a.seg_ptr[s0], L0, lane); e_bases += L0; if (RING > 2) w_build_plan(gv, plan, tbp, lane); else if (HGPU_SHALLOW_TBP) w_build_tbp(gv, tbp, lane); }
            PHASE_CLK(PC_INIT)
            for (uint32_t k = 1; k < R && st == ST_OK; ++k) {
                const uint32_t V = *gv.n_nodes;
                const uint32_t NE = *gv.n_edges;
                const uint32_t L = a.seg_len[s0 + k];
                const uint8_t* seq = a.bases + a.seg_ptr[s0 + k];
                int mode = RING > 2 ? dp_mode_deep(a.sc, a.force_i32) : dp_mode(V, L, a.sc, a.force_i32);
                // the shallow kernel carries no REL16 code (it is instruction-cache bound): the host sends every edge it expects
                // to leave the plain int16 range to the deep kernel; one that does so unexpectedly runs in int32 cells here
                if (RING == 2 && mode == DPM_REL16) mode = DPM_I32;
                const bool p16 = mode != DPM_I32;
                if ((uint64_t)V + L > gv.ncap || (uint64_t)NE + L + 1 > gv.ecap) { st = ST_CAPACITY; break; }
                if (dp_slot_bytes(V, L, mode) > a.slot_bytes) { st = ST_TOO_LARGE; break; }
                const uint32_t probe = (k == a.probe_round) ? a.probe : 0u;
                if (probe == 1) {
                    if (RING > 2 && mode == DPM_REL16) dp_fill_rel<false>(gv, plan, slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, lane, 0, 1, nullptr);
                    else if (mode == DPM_ABS16) dp_fill16<false, false, 2>(gv.meta0, gv.pred_off, gv.pred_rank, slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, Geo<DP_NW16, true>::bias(V, a.sc), lane, 0, 1, nullptr);
                    e_cells += (unsigned long long)(V + 1) * (L + 1);
                    debug_stop = true; break;
                }
                bool ok;
                if (RING == 2 && mode == DPM_ABS16) {
                    dp_fill16<false, false, 2>(gv.meta0, gv.pred_off, gv.pred_rank, slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, Geo<DP_NW16, true>::bias(V, a.sc), lane, 0, 1, nullptr);
                    PHASE_CLK(PC_FILL)
                    ok = dp_traceback<DP_NW16, true, false, HGPU_SHALLOW_TBP != 0>(gv, slot, wsm, seq, V, L, a.sc, lane, tbp);
                } else if (RING > 2 && mode == DPM_REL16) {
                    dp_fill_rel<false>(gv, plan, slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, lane, 0, 1, nullptr);
                    PHASE_CLK(PC_FILL)
                    ok = dp_traceback<DP_NW16, true, true, true>(gv, slot, wsm, seq, V, L, a.sc, lane, tbp);
                } else {
                    dp_fill<DP_NW32, false>(gv, slot, wsm, seq, V, L, a.sc, lane, 0, 1, nullptr);
                    PHASE_CLK(PC_FILL)
                    ok = dp_traceback<DP_NW32, false, false, (RING > 2) || HGPU_SHALLOW_TBP != 0>(gv, slot, wsm, seq, V, L, a.sc, lane, tbp);
                }
                PHASE_CLK(PC_TRACEBACK)
                if (probe == 2) { e_cells += (unsigned long long)(V + 1) * (L + 1); debug_stop = true; break; }
                if (lane == 0) {
                    hdr[HDR_LAST_P16] = (uint32_t)mode; hdr[HDR_LAST_V] = V; hdr[HDR_LAST_L] = L;
                    hdr[HDR_LAST_BIAS] = (uint32_t)(mode == DPM_ABS16 ? Geo<DP_NW16, true>::bias(V, a.sc) : 0);
                }
                e_cells += (unsigned long long)(V + 1) * (L + 1);
                e_padded += (unsigned long long)(V + 1) *
                             (p16 ? Geo<DP_NW16, true>::stripes(L) * Geo<DP_NW16, true>::SW : Geo<DP_NW32, false>::stripes(L) * Geo<DP_NW32, false>::SW);
                e_aln += 1; e_aln32 += mode == DPM_I32 ? 1ull : (mode == DPM_REL16 ? (1ull << 32) : 0ull);   /* low word int32, high word REL16 */ e_bases += L;
                if (!ok) { st = ST_TRACEBACK; break; }
                if (k == a.stop_round) { debug_stop = true; break; }
                // fold the alignment into the graph, re-sort, rebuild the DP records (same results as SPOA's serial code)
                uint32_t ust = w_add_alignment(gv, gs, seq, L, lane);
                if (ust == 0xFFFFFFFFu) {                 // not a full-coverage alignment: serial restatement
                    ust = ST_OK;
                    if (lane == 0 && !g_add_alignment(gv, seq, L)) ust = ST_CAPACITY;
                    ust = __shfl_sync(FULL, ust, 0);
                    __syncwarp();
                }
                if (ust != ST_OK) { st = ust; break; }
                PHASE_CLK(PC_ADD)
                if (probe == 3) { debug_stop = true; break; }
                constexpr bool USE_TREC = RING > 2 || HGPU_SHALLOW_TREC;
                if (USE_TREC) w_build_trec(gv, trec, lane);
                // deep graphs: the roots' walks side by side (poa_topo_claims.cuh); the batched walk when that declines
                bool sorted = false;
                if (HGPU_TOPO_CLAIMS && USE_TREC && (RING > 2 || HGPU_TOPO_CLAIMS_SHALLOW)) sorted = w_toposort_claims(gv, gs, trec, lane) != 0;
                if (!sorted && !(USE_TREC ? w_toposort(gv, trec, wsm, lane) : w_toposort_chain(gv, wsm, lane))) {    // too large for the shared-memory bitmaps / deep DFS: serial
                    ust = ST_OK;
                    if (lane == 0 && !g_toposort(gv, gs)) ust = ST_TOPOSORT;
                    ust = __shfl_sync(FULL, ust, 0);
                    __syncwarp();
                    if (ust != ST_OK) { st = ust; break; }
                }
                PHASE_CLK(PC_TOPO)
                if (probe == 4) { debug_stop = true; break; }
                w_build_meta(gv, lane);
                if (RING > 2) w_build_plan(gv, plan, tbp, lane);
                else if (HGPU_SHALLOW_TBP) w_build_tbp(gv, tbp, lane);
                PHASE_CLK(PC_META)
                if (probe == 5) { debug_stop = true; break; }
            }
            if (a.stop_round != 0xFFFFFFFFu) debug_stop = true;
            // heaviest-bundle consensus; node ids land in aln_rank
            if (st == ST_OK && !debug_stop) {
                if (*gv.n_nodes <= (1u << 20)) {                          // int32 path scores hold
                    const uint32_t max_id = w_consensus_scores(gv, gs, lane);
                    if (gv.out_head[max_id] == NIL) n_cons = w_consensus_backtrack(gv, gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank), lane);
                    else if (lane == 0) n_cons = g_consensus_finish(gv, gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank));   // branch completion: serial
                } else if (lane == 0) n_cons = g_consensus(gv, gs, reinterpret_cast<uint32_t*>(gv.aln_rank));
                n_cons = __shfl_sync(FULL, n_cons, 0);
                __syncwarp();
            }
        }
        PHASE_CLK(PC_CONSENSUS)
        // publish
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
        if (st == ST_OK) { st_cells += e_cells; st_padded += e_padded; st_aln += e_aln; st_aln32 += e_aln32; st_bases += e_bases; }
        if (lane == 0) {
            a.status[e] = st;
            a.cons_len[e] = (st == ST_OK) ? n_cons : 0;
            a.cons_pos[e] = (uint64_t)(uintptr_t)(a.pool + pos);
            if (a.out_nodes) a.out_nodes[e] = *gv.n_nodes;
        }
        __syncwarp();
        PHASE_CLK(PC_PUBLISH)
    }
    PHASE_CLK_FLUSH(a)
    if (lane == 0 && a.stats) {
        atomicAdd(a.stats + 0, st_cells); atomicAdd(a.stats + 1, st_padded); atomicAdd(a.stats + 2, st_aln);
        atomicAdd(a.stats + 3, st_aln32); atomicAdd(a.stats + 4, st_bases);
    }
}

// shallow edges (a handful of supporting reads): as many resident warps as possible, two parked rows per warp
__global__ void __launch_bounds__(32 * DP_WARPS_PER_BLOCK, HGPU_MINBLOCKS) k_poa_edges(PoaArgs a) { poa_edges_body<2>(a); }
// deep edges: a ring of parked rows per warp (4 blocks of 4 warps per SM by shared memory)
__global__ void __launch_bounds__(32 * DP_WARPS_PER_BLOCK, 4) k_poa_edges_deep(PoaArgs a) { poa_edges_body<DP_RING_DEEP>(a); }

// ---------------------------------------------------------------------------------------------------------
// k_poa_edges_team<TEAM>: one BLOCK of TEAM warps per backbone edge, for edges whose score matrices are so large
// that a single warp would become the tail of the whole batch (deep coverage x long gap). The fill of one
// alignment is spread over the team stripe by stripe (pipelined through the boundary columns, see dp_fill);
// traceback, graph update and consensus run on warp 0 exactly as in k_poa_edges, so results are identical.
// ---------------------------------------------------------------------------------------------------------
template <int TEAM>
__global__ void __launch_bounds__(32 * TEAM, 512 / (32 * TEAM)) k_poa_edges_team(PoaArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = lane_id();
    const uint32_t wib = threadIdx.x >> 5;                  // = rank in the team
    const uint32_t gw = blockIdx.x;                          // one slot / workspace per team
    uint8_t* wsm = smem_raw + (size_t)wib * DP_SMEM_PER_WARP_DEEP;
    volatile uint32_t* vprog = reinterpret_cast<volatile uint32_t*>(smem_raw + (size_t)TEAM * DP_SMEM_PER_WARP_DEEP);   // [TEAM]
    volatile uint32_t* bcast = vprog + TEAM;                 // [4]: item, status, flag, spare
    uint8_t* slot = a.arena + (uint64_t)gw * a.slot_bytes;
    uint8_t* wsb = a.ws + (uint64_t)gw * a.wl.bytes;
    GraphView gv = bind_graph(wsb, a.wl);
    GraphScratch gs = bind_scratch(wsb, a.wl);
    uint32_t* hdr = reinterpret_cast<uint32_t*>(wsb + a.wl.o_hdr);
    uint32_t* plan = reinterpret_cast<uint32_t*>(wsb + a.wl.o_plan);
    uint32_t* tbp = reinterpret_cast<uint32_t*>(wsb + a.wl.o_tbp);
    TopoRec* trec = reinterpret_cast<TopoRec*>(wsb + a.wl.o_trec);
    unsigned long long st_cells = 0, st_padded = 0, st_aln = 0, st_aln32 = 0, st_bases = 0;
    const bool lead = wib == 0;

    while (true) {
        if (threadIdx.x == 0) bcast[0] = atomicAdd(a.counter, 1u);
        __syncthreads();
        const uint32_t item = bcast[0];
        if (item >= a.n_items) break;
        const uint32_t e = a.items[item];
        const uint32_t s0 = a.e_seg_off[e];
        const uint32_t R = a.e_seg_off[e + 1] - s0;
        uint32_t st = ST_OK;
        uint32_t n_cons = 0;
        bool debug_stop = false;
        unsigned long long e_cells = 0, e_padded = 0, e_aln = 0, e_aln32 = 0, e_bases = 0;
        if (R == 0) {
            if (threadIdx.x == 0) { *gv.n_nodes = 0; *gv.n_edges = 0; *gv.aln_len = 0; }
        } else {
            const uint32_t L0 = a.seg_len[s0];
            if (L0 > gv.ncap || L0 > gv.ecap) st = ST_CAPACITY;
            else if (lead) { w_init_chain(gv, a.bases + a.seg_ptr[s0], L0, lane); e_bases += L0; w_build_plan(gv, plan, tbp, lane); }
            for (uint32_t k = 1; k < R && st == ST_OK; ++k) {
                if (threadIdx.x < TEAM) vprog[threadIdx.x] = 0;
                __syncthreads();                              // graph of round k-1 complete, progress words cleared
                const uint32_t V = *gv.n_nodes;
                const uint32_t NE = *gv.n_edges;
                const uint32_t L = a.seg_len[s0 + k];
                const uint8_t* seq = a.bases + a.seg_ptr[s0 + k];
                const int mode = dp_mode_deep(a.sc, a.force_i32);
                const bool p16 = mode != DPM_I32;
                if ((uint64_t)V + L > gv.ncap || (uint64_t)NE + L + 1 > gv.ecap) { st = ST_CAPACITY; break; }
                if (dp_slot_bytes(V, L, mode) > a.slot_bytes) { st = ST_TOO_LARGE; break; }
                const bool fill_ok = mode == DPM_REL16 ? dp_fill_rel<true>(gv, plan, slot, wsm, seq, V, L, a.sc.sm, a.sc.sx, a.sc.g, lane, wib, TEAM, vprog)
                                                       : dp_fill<DP_NW32, false>(gv, slot, wsm, seq, V, L, a.sc, lane, wib, TEAM, vprog);
                const int all_ok = __syncthreads_and(fill_ok ? 1 : 0);    // every stripe stored (and no wait gave up)
                if (!all_ok) { st = ST_SYNC; break; }
                uint32_t rst = ST_OK;
                if (lead) {
                    bool ok = mode == DPM_REL16 ? dp_traceback<DP_NW16, true, true, true>(gv, slot, wsm, seq, V, L, a.sc, lane, tbp)
                                                : dp_traceback<DP_NW32, false, false, true>(gv, slot, wsm, seq, V, L, a.sc, lane, tbp);
                    if (lane == 0) {
                        hdr[HDR_LAST_P16] = (uint32_t)mode; hdr[HDR_LAST_V] = V; hdr[HDR_LAST_L] = L;
                        hdr[HDR_LAST_BIAS] = (uint32_t)(mode == DPM_ABS16 ? Geo<DP_NW16, true>::bias(V, a.sc) : 0);
                    }
                    e_cells += (unsigned long long)(V + 1) * (L + 1);
                    e_padded += (unsigned long long)(V + 1) *
                                 (p16 ? Geo<DP_NW16, true>::stripes(L) * Geo<DP_NW16, true>::SW : Geo<DP_NW32, false>::stripes(L) * Geo<DP_NW32, false>::SW);
                    e_aln += 1; e_aln32 += mode == DPM_I32 ? 1ull : (mode == DPM_REL16 ? (1ull << 32) : 0ull);   /* low word int32, high word REL16 */ e_bases += L;
                    if (!ok) rst = ST_TRACEBACK;
                    else if (k != a.stop_round) {
                        uint32_t ust = w_add_alignment(gv, gs, seq, L, lane);
                        if (ust == 0xFFFFFFFFu) {
                            ust = ST_OK;
                            if (lane == 0 && !g_add_alignment(gv, seq, L)) ust = ST_CAPACITY;
                            ust = __shfl_sync(FULL, ust, 0);
                            __syncwarp();
                        }
                        if (ust == ST_OK) w_build_trec(gv, trec, lane);
                        if (ust == ST_OK && !(HGPU_TOPO_CLAIMS && w_toposort_claims(gv, gs, trec, lane)) && !w_toposort(gv, trec, wsm, lane)) {
                            if (lane == 0 && !g_toposort(gv, gs)) ust = ST_TOPOSORT;
                            ust = __shfl_sync(FULL, ust, 0);
                            __syncwarp();
                        }
                        if (ust == ST_OK) { w_build_meta(gv, lane); w_build_plan(gv, plan, tbp, lane); }
                        rst = ust;
                    }
                    if (lane == 0) bcast[1] = rst;
                }
                __syncthreads();
                rst = bcast[1];
                if (rst != ST_OK) { st = rst; break; }
                if (k == a.stop_round) { debug_stop = true; break; }
            }
            if (a.stop_round != 0xFFFFFFFFu) debug_stop = true;
            if (st == ST_OK && !debug_stop && lead) {
                if (*gv.n_nodes <= (1u << 20)) {                          // int32 path scores hold
                    const uint32_t max_id = w_consensus_scores(gv, gs, lane);
                    if (gv.out_head[max_id] == NIL) n_cons = w_consensus_backtrack(gv, gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank), lane);
                    else if (lane == 0) n_cons = g_consensus_finish(gv, gs, max_id, reinterpret_cast<uint32_t*>(gv.aln_rank));   // branch completion: serial
                } else if (lane == 0) n_cons = g_consensus(gv, gs, reinterpret_cast<uint32_t*>(gv.aln_rank));
                n_cons = __shfl_sync(FULL, n_cons, 0);
                __syncwarp();
            }
        }
        if (lead) {
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
            if (st == ST_OK) { st_cells += e_cells; st_padded += e_padded; st_aln += e_aln; st_aln32 += e_aln32; st_bases += e_bases; }
            if (lane == 0) {
                a.status[e] = st;
                a.cons_len[e] = (st == ST_OK) ? n_cons : 0;
                a.cons_pos[e] = (uint64_t)(uintptr_t)(a.pool + pos);
                if (a.out_nodes) a.out_nodes[e] = *gv.n_nodes;
            }
        }
        __syncthreads();                                      // bcast[0] may be rewritten now
    }
    if (threadIdx.x == 0 && a.stats) {
        atomicAdd(a.stats + 0, st_cells); atomicAdd(a.stats + 1, st_padded); atomicAdd(a.stats + 2, st_aln);
        atomicAdd(a.stats + 3, st_aln32); atomicAdd(a.stats + 4, st_bases);
    }
}

}  // namespace hgpu
#include "poa_pool.cuh"
namespace hgpu {

// out[off[e] .. off[e]+len[e]) = consensus bytes of edge e
__global__ void __launch_bounds__(256) k_poa_gather(const uint64_t* cons_pos, const uint32_t* cons_len,
                                                    const uint64_t* off, uint32_t n_edges, uint8_t* out, uint64_t out_cap) {
    for (uint32_t e = blockIdx.x; e < n_edges; e += gridDim.x) {
        const uint32_t n = cons_len[e];
        const uint8_t* src = reinterpret_cast<const uint8_t*>(cons_pos[e]);
        const uint64_t o = off[e];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) if (o + i < out_cap) out[o + i] = src[i];
    }
}

// Debug: export the score matrix of warp 0's last alignment in the reference's H space (row-major (V+1)*(L+1) int32).
template <int NW, bool P16, bool REL = false>
__global__ void k_poa_dump_H(uint8_t* slot, uint32_t V, uint32_t L, int bias, int gap, int32_t* H) {
    SlotView<NW, P16, REL> sv;
    sv.bind(slot, V, L);
    const uint64_t total = (uint64_t)(V + 1) * (L + 1);
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t i = (uint32_t)(idx / (L + 1)), j = (uint32_t)(idx % (L + 1));
        H[idx] = sv.load(i, j) - bias + (int)j * gap;
    }
}

#endif  // __CUDACC__
}  // namespace hgpu
