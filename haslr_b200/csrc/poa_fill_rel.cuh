// Score-matrix fill of the deep-edge kernels (k_poa_edges_deep, k_poa_edges_team, k_poa_pool): every alignment in
// row-relative int16 cells (DPM_REL16), one lean row body.
//
// Included from poa_device.cuh (uses its Geo / Fill16 / shared-memory helpers). Replaces, for edges with tens of
// supporting reads, the graph-NW engine SPOA runs per read (reference Assemble.cpp:539 -> spoa::align_sequence_with_graph).
//
// What is different from dp_fill16 (the shallow kernel's fill), and why:
//  * A deep POA graph is 2-3x wider than the gap: at depth 27 a rank has 1.8 predecessor rows on average, 1.2 of them not
//    the previous rank, nearly all (99.8 %) within 8 ranks. The row body is therefore a loop over predecessor rows with
//    THREE packed instructions per word and predecessor - x = src[k-1] + profile, y = max(src[k] + gap, x),
//    t[k] = max(t[k], y + rebase) - so re-basing a predecessor row into this row's frame costs nothing extra and the
//    plain-int16 encoding has no advantage left: all alignments run relative, one code path.
//  * The predecessor list of every rank is condensed once per alignment into a one-word PLAN (eight 3-bit rank distances,
//    count, base code; w_build_plan) that the stripes load 32 ranks at a time; rows whose predecessors do not fit the plan
//    (farther than the ring, more than eight, or none but the virtual row) take the generic walk over the CSR.
//  * Row bases are int32 Hhat values of the cell left of the stripe (column 0 itself in stripe 0), kept per stripe in the
//    boundary-column arrays the stripes hand to each other anyway; lane q of two batch registers holds the bases of the
//    current and the previous 32 rows, so a predecessor's base is one shuffle.
#pragma once

namespace hgpu {

// Plan word of a rank: bits 0-23 the rank distances of up to eight predecessor rows, 3 bits each (distance - 1; the previous
// rank first when it is one of them), bits 24-25 the node's base code, bits 26-29 the predecessor count, bit 30 PLAN_SLOW.
static constexpr uint32_t PLAN_SLOW = 1u << 30;
// 1: lane 31 stores every row's base for the next stripe itself instead of shuffling it into a batch register that is stored every 32
// rows (one shuffle + two ALU instructions less per row; +1 % on the pool kernel, profiles/r2C_row_variants_ab.log)
#ifndef HGPU_REL_BCO_STORE
#define HGPU_REL_BCO_STORE 1
#endif
static constexpr int REL_RING = DP_RING_DEEP;            // parked rows per warp; plan distances are 1 .. REL_RING
static_assert(REL_RING == 8 || REL_RING == 4, "plan distances are packed in 3 bits; the ring index is a mask");
__device__ __forceinline__ uint32_t plan_code(uint32_t p) { return (p >> 24) & 3u; }
__device__ __forceinline__ uint32_t plan_np(uint32_t p) { return (p >> 26) & 15u; }

__device__ __forceinline__ uint32_t deep_plan_of(uint32_t m0, uint32_t rr, const uint32_t* pred_off, const uint32_t* pred_rank) {
    const uint32_t code = m0 & 3u, npc = (m0 >> 3) & 3u;
    const uint32_t slow = (code << 24) | PLAN_SLOW;
    auto pack = [&](uint32_t dists, uint32_t n) { return dists | (code << 24) | (n << 26); };
    if (npc == 0) return rr == 0 ? pack(0u, 1u) : slow;                  // first rank: the virtual row 0 is the previous row
    if (npc == 1) {
        const uint32_t d0 = meta_d0(m0);
        return d0 <= (uint32_t)REL_RING ? pack(d0 - 1u, 1u) : slow;
    }
    if (npc == 2) {
        uint32_t d0 = meta_d0(m0), d1 = meta_d1(m0);
        if (d0 > (uint32_t)REL_RING || d1 > (uint32_t)REL_RING) return slow;
        if (d1 == 1u) { d1 = d0; d0 = 1u; }
        return pack((d0 - 1u) | ((d1 - 1u) << 3), 2u);
    }
    const uint32_t c0 = pred_off[rr], n = pred_off[rr + 1] - c0;
    if (n > 8u) return slow;
    uint32_t dists = 0, at = 1;                                           // slot 0 is kept for the previous rank
    bool ok = true, has1 = false;
    for (uint32_t x = 0; x < n; ++x) {
        const uint32_t d = rr - pred_rank[c0 + x];
        ok = ok && d <= (uint32_t)REL_RING;
        if (d == 1u) has1 = true;
        else { dists |= ((d - 1u) & 7u) << (3u * at); ++at; }
    }
    if (!ok) return slow;
    if (!has1) dists >>= 3;                                               // no previous-rank predecessor: close the gap
    return pack(dists, n);
}

// The traceback's word of rank rr: the rank distances of up to six predecessors in in-edge order (SPOA's preference order), five bits
// each, first predecessor in the low bits; TBP_GENERIC when there are more, or one is 32 or more ranks away (it could never be inside
// a 32-row tile). A node without in-edges has the virtual row 0 as its predecessor: distance = its row.
__device__ __forceinline__ uint32_t tb_plan_of(uint32_t m0, uint32_t rr, const uint32_t* pred_off, const uint32_t* pred_rank) {
    const uint32_t npc = (m0 >> 3) & 3u;
    if (npc == 0) return rr + 1u <= 31u ? rr + 1u : TBP_GENERIC;
    if (npc == 1) { const uint32_t d0 = meta_d0(m0); return d0 <= 31u ? d0 : TBP_GENERIC; }
    if (npc == 2) { const uint32_t d0 = meta_d0(m0), d1 = meta_d1(m0); return (d0 <= 31u && d1 <= 31u) ? (d0 | (d1 << 5)) : TBP_GENERIC; }
    const uint32_t c0 = pred_off[rr], n = pred_off[rr + 1] - c0;
    if (n == 0 || n > 6u) return TBP_GENERIC;
    uint32_t w = 0;
    for (uint32_t x = 0; x < n; ++x) {
        const uint32_t d = rr - pred_rank[c0 + x];
        if (d == 0 || d > 31u) return TBP_GENERIC;
        w |= d << (5u * x);
    }
    return w;
}

// plan words of all ranks, lane-parallel; run by the warp that owns the graph, after the per-rank DP records exist
__device__ __noinline__ void w_build_plan(const GraphView& g, uint32_t* plan, uint32_t* tbp, int lane) {
    const uint32_t N = *g.n_nodes;
    const uint32_t* const meta0 = g.meta0; const uint32_t* const pred_off = g.pred_off; const uint32_t* const pred_rank = g.pred_rank;
    for (uint32_t r = lane; r < N; r += 32) {
        const uint32_t m0 = meta0[r];
        plan[r] = deep_plan_of(m0, r, pred_off, pred_rank);
        tbp[r] = tb_plan_of(m0, r, pred_off, pred_rank);
    }
    __syncwarp();
}

// the traceback words alone (the warp-per-edge kernel has no fill plan)
__device__ __noinline__ void w_build_tbp(const GraphView& g, uint32_t* tbp, int lane) {
    const uint32_t N = *g.n_nodes;
    const uint32_t* const meta0 = g.meta0; const uint32_t* const pred_off = g.pred_off; const uint32_t* const pred_rank = g.pred_rank;
    for (uint32_t r = lane; r < N; r += 32) tbp[r] = tb_plan_of(meta0[r], r, pred_off, pred_rank);
    __syncwarp();
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
// shared address of the base of row `row` of the current stripe (rows of this batch and the one before: 64 words behind the ring)
#define REL_BASE_ADDR(S, row) ((S).frame + (uint32_t)(FILL16_PARKED + REL_RING * 1024) + (((row) & 63u) << 2))
#define REL_PLAN_ADDR(S, q) ((S).frame + (uint32_t)(FILL16_PARKED + REL_RING * 1024 + 256) + ((uint32_t)(q) << 2))

struct RelFrame {                                         // cold per-alignment state, in the warp's shared memory behind the profile
    unsigned long long meta0, pred_off, pred_rank, plan, seq, H, bases;
    uint32_t V, L, NS; int32_t sm, sx, pad;
};
static_assert(sizeof(RelFrame) <= FILL16_PARKED, "frame must fit in front of the parked rows");
#define RFRAME(field) ((uint32_t)offsetof(RelFrame, field))

// Progress words of a team (shared memory, monotone): the stripe being filled publishes pub_base + rows done + 1 in
// vprog[pub_idx] and, before a batch of 32 rows, waits for vprog[wait_idx] >= wait_base + rows it needs of the stripe to its left.
struct TeamSync { volatile uint32_t* vprog; uint32_t pub_idx, wait_idx, pub_base, wait_base; };

struct RelState {
    uint32_t pf_lane;        // shared address of this lane's 16 bytes of prof[0][0]
    uint32_t frame;          // shared address of the RelFrame; parked row r is at frame + 128 + (r & 7) * 1024 as [half][lane][4 words]
    uint32_t g2;             // packed gap
    uint32_t row_bytes;      // distance between consecutive rows of the stripe in the slot
    uint64_t dst;            // slot address of this lane's first unit of row i
    uint32_t b, bprev;       // batch registers: lane q holds the base of row r0+q+1 (this batch / the previous batch) in this stripe
    uint32_t bco;            // batch register: lane q collects the last cell (absolute) of row r0+q+1, the base of that row in the next stripe
    uint32_t stripe;         // s
#if HGPU_REL_BCO_STORE
    unsigned long long bnext; // base array of the next stripe
#endif
    int lane; bool has_prev, has_next;
};

// t[k] = max(t[k], max(src[k] + gap, src[k-1] + profile[k]) + rebase); hs = the word left of src[0]
#define REL_FOLD(s0, s1, s2, s3, s4, s5, s6, s7, hs, d2)                                                    \
    do {                                                                                                       \
        t[7] = __viaddmax_s16x2(__viaddmax_s16x2(s7, g2, __vadd2(s6, p1.w)), d2, t[7]);                       \
        t[6] = __viaddmax_s16x2(__viaddmax_s16x2(s6, g2, __vadd2(s5, p1.z)), d2, t[6]);                       \
        t[5] = __viaddmax_s16x2(__viaddmax_s16x2(s5, g2, __vadd2(s4, p1.y)), d2, t[5]);                       \
        t[4] = __viaddmax_s16x2(__viaddmax_s16x2(s4, g2, __vadd2(s3, p1.x)), d2, t[4]);                       \
        t[3] = __viaddmax_s16x2(__viaddmax_s16x2(s3, g2, __vadd2(s2, p0.w)), d2, t[3]);                       \
        t[2] = __viaddmax_s16x2(__viaddmax_s16x2(s2, g2, __vadd2(s1, p0.z)), d2, t[2]);                       \
        t[1] = __viaddmax_s16x2(__viaddmax_s16x2(s1, g2, __vadd2(s0, p0.y)), d2, t[1]);                       \
        t[0] = __viaddmax_s16x2(__viaddmax_s16x2(s0, g2, __vadd2(hs, p0.x)), d2, t[0]);                       \
    } while (0)
// the first predecessor of a row: t[k] = max(src[k] + gap, src[k-1] + profile[k]) + rebase
#define REL_FOLD_FIRST(s0, s1, s2, s3, s4, s5, s6, s7, hs, d2)                                              \
    do {                                                                                                       \
        t[7] = __vadd2(__viaddmax_s16x2(s7, g2, __vadd2(s6, p1.w)), d2);                                      \
        t[6] = __vadd2(__viaddmax_s16x2(s6, g2, __vadd2(s5, p1.z)), d2);                                      \
        t[5] = __vadd2(__viaddmax_s16x2(s5, g2, __vadd2(s4, p1.y)), d2);                                      \
        t[4] = __vadd2(__viaddmax_s16x2(s4, g2, __vadd2(s3, p1.x)), d2);                                      \
        t[3] = __vadd2(__viaddmax_s16x2(s3, g2, __vadd2(s2, p0.w)), d2);                                      \
        t[2] = __vadd2(__viaddmax_s16x2(s2, g2, __vadd2(s1, p0.z)), d2);                                      \
        t[1] = __vadd2(__viaddmax_s16x2(s1, g2, __vadd2(s0, p0.y)), d2);                                      \
        t[0] = __vadd2(__viaddmax_s16x2(s0, g2, __vadd2(hs, p0.x)), d2);                                      \
    } while (0)

// One row, in place: A = row i-1 on entry and row i on exit (stored to the slot and, by the next row, parked in the ring).
// plan: the row's plan word; bi: its base (stripes > 0; stripe 0 derives it from the predecessors' bases here).
__device__ __forceinline__ void row_rel(uint32_t (&A)[8], RelState& S, uint32_t plan, int bi, int q, uint32_t i) {
    using F = Fill16;
    const int lane = S.lane;
    const uint32_t g2 = S.g2;
    const uint32_t parked = S.frame + FILL16_PARKED + (uint32_t)lane * 16u;
    const int gap = (int)(int16_t)(g2 & 0xFFFFu);
    // park row i-1: later rows read it through the ring
    sts_v4(parked + ((i - 1) & (uint32_t)(REL_RING - 1)) * 1024u, A[0], A[1], A[2], A[3]);
    sts_v4(parked + ((i - 1) & (uint32_t)(REL_RING - 1)) * 1024u + 512u, A[4], A[5], A[6], A[7]);
    const uint32_t pf = S.pf_lane + plan_code(plan) * (uint32_t)(F::NW * 32 * 4);
    const uint4 p0 = lds_v4(pf), p1 = lds_v4(pf + 512u);
    const uint32_t blw = (uint32_t)(S.has_prev ? 0 : F::G::NEGV) << 16;   // the cell left of the stripe in a row's own frame: its base, or nothing
    uint32_t t[8];
    // base (int32 Hhat) of the row `dist` ranks back, dist <= 8: lane q - dist of this batch's register, or of the previous batch's
#if HGPU_REL_SMEM_BASES
    auto base_near = [&](uint32_t dist) -> int { return (int)lds_u32v(REL_BASE_ADDR(S, i - dist)); };     // one broadcast load, no shuffle
#else
    auto base_near = [&](uint32_t dist) -> int {
        const int ql = q - (int)dist;
        return __shfl_sync(FULL, (int)(ql >= 0 ? S.b : S.bprev), ql & 31);
    };
#endif
    if ((plan & PLAN_SLOW) == 0) {
        const uint32_t np = plan_np(plan);
        if (!S.has_prev) {                                                // stripe 0: base = column 0 = gap + best predecessor base
            int best = INT32_MIN;
            uint32_t dd = plan;
            for (uint32_t x = 0; x < np; ++x, dd >>= 3) best = max(best, base_near((dd & 7u) + 1u));
            bi = best + gap;
#if HGPU_REL_SMEM_BASES
            sts_u32(REL_BASE_ADDR(S, i), (uint32_t)bi);
#else
            if (lane == q) S.b = (uint32_t)bi;
#endif
        }
        // the first predecessor initialises the row: the previous rank (row i-1, still in registers) if it is one of them
        uint32_t dd = plan;
        {
            const uint32_t dist = (dd & 7u) + 1u;
            const uint32_t d2 = pack2(max(base_near(dist) - bi, REL_CLAMP));
            if (dist == 1u) {
                uint32_t left = __shfl_up_sync(FULL, A[7], 1);
                if (lane == 0) left = blw;
                const uint32_t hs = __byte_perm(left, A[7], 0x5432);
                REL_FOLD_FIRST(A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], hs, d2);
            } else {
                const uint32_t pa_ = parked + ((i - dist) & (uint32_t)(REL_RING - 1)) * 1024u;
                const uint4 s0 = lds_v4(pa_), s1 = lds_v4(pa_ + 512u);
                uint32_t left = __shfl_up_sync(FULL, s1.w, 1);
                if (lane == 0) left = blw;
                const uint32_t hs = __byte_perm(left, s1.w, 0x5432);
                REL_FOLD_FIRST(s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, hs, d2);
            }
        }
        // the others (never the previous rank: the plan lists it first) come from the ring
#pragma unroll 1
        for (uint32_t x = 1; x < np; ++x) {
            dd >>= 3;
            const uint32_t dist = (dd & 7u) + 1u;
            const uint32_t d2 = pack2(max(base_near(dist) - bi, REL_CLAMP));
            const uint32_t pa_ = parked + ((i - dist) & (uint32_t)(REL_RING - 1)) * 1024u;
            const uint4 s0 = lds_v4(pa_), s1 = lds_v4(pa_ + 512u);
            uint32_t left = __shfl_up_sync(FULL, s1.w, 1);
            if (lane == 0) left = blw;
            const uint32_t hs = __byte_perm(left, s1.w, 0x5432);
            REL_FOLD(s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, hs, d2);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = F::NEG2;
        // generic walk: predecessors from the CSR (or the virtual row 0 when there is none), bases of rows outside the two
        // batch registers from the stripe's base array, cells of rows outside the ring from the slot
        const unsigned long long bases = lds_u64(S.frame + RFRAME(bases)) + 4ull * S.stripe * (lds_u32v(S.frame + RFRAME(V)) + 1ull);
        const unsigned long long poff = lds_u64(S.frame + RFRAME(pred_off)), prank = lds_u64(S.frame + RFRAME(pred_rank));
        const uint32_t cs = ldg_u32(poff, i - 1);
        const uint32_t npr = ldg_u32(poff, i) - cs;
        const uint32_t np = npr == 0 ? 1u : npr;
        auto dist_of = [&](uint32_t x) -> uint32_t { return npr == 0 ? i : i - (ldg_u32(prank, cs + x) + 1); };
        auto base_any = [&](uint32_t dist) -> int {
            const int ql = q - (int)dist;
#if HGPU_REL_SMEM_BASES
            if (ql >= -32) return (int)lds_u32v(REL_BASE_ADDR(S, i - dist));
#else
            if (ql >= 0) return __shfl_sync(FULL, (int)S.b, ql);
            if (ql >= -32) return __shfl_sync(FULL, (int)S.bprev, ql + 32);
#endif
            return (int)ldg_u32(bases, i - dist);
        };
        if (!S.has_prev) {
            int best = INT32_MIN;
            for (uint32_t x = 0; x < np; ++x) best = max(best, base_any(dist_of(x)));
            bi = best + gap;
#if HGPU_REL_SMEM_BASES
            sts_u32(REL_BASE_ADDR(S, i), (uint32_t)bi);
#else
            if (lane == q) S.b = (uint32_t)bi;
#endif
        }
#pragma unroll 1
        for (uint32_t x = 0; x < np; ++x) {
            const uint32_t dist = dist_of(x);
            const uint32_t d2 = pack2(max(base_any(dist) - bi, REL_CLAMP));
            uint4 s0, s1;
            if (dist == 1u) {
                s0 = make_uint4(A[0], A[1], A[2], A[3]); s1 = make_uint4(A[4], A[5], A[6], A[7]);
            } else if (dist <= (uint32_t)REL_RING) {
                const uint32_t pa_ = parked + ((i - dist) & (uint32_t)(REL_RING - 1)) * 1024u;
                s0 = lds_v4(pa_); s1 = lds_v4(pa_ + 512u);
            } else {
                const uint64_t src = S.dst - (uint64_t)dist * S.row_bytes;
                s0 = ldg_v4(src, 0); s1 = ldg_v4(src, 1);
            }
            uint32_t left = __shfl_up_sync(FULL, s1.w, 1);
            if (lane == 0) left = blw;
            const uint32_t hs = __byte_perm(left, s1.w, 0x5432);
            REL_FOLD(s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, hs, d2);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) A[k] = t[k];
    // horizontal gaps = prefix maximum in hat space (see row16): in the lane, then across lanes
#pragma unroll
    for (int k = 1; k < 8; ++k) A[k] = __vmaxs2(A[k], A[k - 1]);
    const uint32_t both = __vmaxs2(A[7], __byte_perm(A[7], 0, 0x1032));
    const int tot = (int)(int16_t)(both & 0xFFFFu);
    const int carry_in = S.has_prev ? 0 : F::G::NEGV;               // the cell left of the stripe is the base itself
    const int nbv = __shfl_up_sync(FULL, tot, 1);
    const int rowmax = __reduce_max_sync(FULL, tot);
    const int prevv = (lane == 0) ? carry_in : nbv;
    const int amax = __ffs(__ballot_sync(FULL, tot == rowmax)) - 1;
    int excl;
    if (__ballot_sync(FULL, lane <= amax && tot < prevv) == 0) {
        excl = (lane <= amax) ? prevv : rowmax;
    } else {
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl = max(incl, v);
        }
        excl = __shfl_up_sync(FULL, incl, 1);
        excl = (lane == 0) ? carry_in : max(excl, carry_in);
    }
    const uint32_t mid = __byte_perm(A[7], F::NEG2, 0x1054);
    const uint32_t cc = __vmaxs2(mid, __byte_perm((uint32_t)excl, 0, 0x1010));
#pragma unroll
    for (int k = 0; k < 8; ++k) A[k] = __vmaxs2(A[k], cc);
    stg_cs_v4(S.dst, A[0], A[1], A[2], A[3]);
    stg_cs_v4_512(S.dst, A[4], A[5], A[6], A[7]);
    S.dst += S.row_bytes;
#if HGPU_REL_BCO_STORE
    if (S.has_next && lane == 31) stg_u32(S.bnext + 4ull * i, (uint32_t)(((int32_t)A[7] >> 16) + bi));   // the row's base in the next stripe
#else
    if (S.has_next) {
        const uint32_t last = __shfl_sync(FULL, A[7], 31);
        if (lane == q) S.bco = (uint32_t)(((int32_t)last >> 16) + bi);
    }
#endif
}

// profile of stripe s (same layout as fill16_profile), from the RelFrame
__device__ __noinline__ void rel_profile(uint32_t* prof, const RelFrame* frame, uint32_t s, int lane) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    const uint8_t* seq = reinterpret_cast<const uint8_t*>((uintptr_t)frame->seq);
    const uint32_t L = frame->L;
    const int sm = frame->sm, sx = frame->sx;
    const uint32_t j0 = s * G::SW + lane * G::CPL;
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

// One stripe of one alignment. TEAM: stripes of the alignment run on different warps; stripe s may start a batch of 32 rows
// once stripe s-1 has published those rows' boundary column (TeamSync).
template <bool TEAM>
__device__ __forceinline__ bool rel_stripe(RelState& S, uint32_t* prof, RelFrame* frame, uint32_t s, int lane, const TeamSync ts) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    bool sync_ok = true;
    __syncwarp();
    rel_profile(prof, frame, s, lane);
    const uint32_t Vs = lds_u32v(S.frame + RFRAME(V));
    const unsigned long long bases_all = lds_u64(S.frame + RFRAME(bases));
    const unsigned long long b_cur = bases_all + 4ull * s * (Vs + 1), b_next = bases_all + 4ull * (s + 1) * (Vs + 1);
    __syncwarp();
    asm volatile("" : "+r"(S.pf_lane) :: "memory");                  // profile loads below may not move above this point
    S.has_prev = s > 0;
    S.has_next = s + 1 < lds_u32v(S.frame + RFRAME(NS));
    S.stripe = s;
#if HGPU_REL_BCO_STORE
    S.bnext = b_next;
#endif
    // --- row 0: Hhat = 0 everywhere, base 0
    uint32_t A[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) A[k] = 0u;
    S.dst = lds_u64(S.frame + RFRAME(H)) + ((uint64_t)s * (G::UNITS * 32) + lane) * 16;
    stg_cs_v4(S.dst, 0u, 0u, 0u, 0u);
    stg_cs_v4_512(S.dst, 0u, 0u, 0u, 0u);
    S.dst += S.row_bytes;
    if (lane == 31) stg_u32(b_next, 0u);
    if (s == 0 && lane == 0) stg_u32(b_cur, 0u);
    if (TEAM) team_publish(ts.vprog, ts.pub_idx, ts.pub_base + 1u, lane);
    S.b = 0u;                                                          // "previous batch" of the first batch: row 0 (base 0) in lane 31
#if HGPU_REL_SMEM_BASES
    sts_u32(REL_BASE_ADDR(S, (uint32_t)lane), 0u); sts_u32(REL_BASE_ADDR(S, (uint32_t)lane + 32u), 0u);   // row 0: base 0
#endif
    const unsigned long long plan = lds_u64(S.frame + RFRAME(plan));
    uint32_t npl = 0;
    if ((uint32_t)lane < Vs) npl = ldg_u32(plan, lane);
#pragma unroll 1
    for (uint32_t r0 = 0; r0 < Vs; r0 += 32) {
        const uint32_t rr = r0 + lane;
        const uint32_t mpl = npl;
        if (rr + 32 < Vs) npl = ldg_u32(plan, rr + 32);
        S.bprev = S.b;
        if (s > 0) {
            if (TEAM) {                                                // rows r0 .. r0+32 of the stripe to the left must be complete
                const uint32_t need_rows = (r0 + 33 < Vs + 1) ? r0 + 33 : Vs + 1;
                if (!team_wait(ts.vprog, ts.wait_idx, ts.wait_base + need_rows, lane)) sync_ok = false;
            }
            S.b = (rr < Vs) ? ldg_u32(b_cur, rr + 1) : 0u;
#if HGPU_REL_SMEM_BASES
            sts_u32(REL_BASE_ADDR(S, rr + 1u), S.b);
#endif
        }
#if HGPU_REL_SMEM_BASES
        sts_u32(REL_PLAN_ADDR(S, lane), mpl);
        __syncwarp();
#endif
        const int nb = (Vs - r0) < 32u ? (int)(Vs - r0) : 32;
#pragma unroll 1
        for (int q = 0; q < nb; ++q) {
#if HGPU_REL_SMEM_BASES
            const uint32_t pl = lds_u32v(REL_PLAN_ADDR(S, q));
            const int bi = (int)lds_u32v(REL_BASE_ADDR(S, r0 + (uint32_t)q + 1u));   // stripe 0: replaced inside the row
#else
            const uint32_t pl = __shfl_sync(FULL, mpl, q);
            const int bi = __shfl_sync(FULL, (int)S.b, q);             // stripe 0: replaced inside the row
#endif
            row_rel(A, S, pl, bi, q, r0 + q + 1);
        }
#if HGPU_REL_SMEM_BASES
        if (s == 0 && lane < nb) stg_u32(b_cur + 4ull * (rr + 1), lds_u32v(REL_BASE_ADDR(S, rr + 1u)));
#else
        if (s == 0 && lane < nb) stg_u32(b_cur + 4ull * (rr + 1), S.b);
#endif
        if (s == 0) __syncwarp();                                      // later generic rows read these through other lanes' loads
#if !HGPU_REL_BCO_STORE
        if (S.has_next && lane < nb) stg_u32(b_next + 4ull * (rr + 1), S.bco);
#endif
        if (TEAM) team_publish(ts.vprog, ts.pub_idx, ts.pub_base + ((r0 + 32 < Vs) ? r0 + 32 : Vs) + 1, lane);
    }
    __syncwarp();
    return sync_ok;
}

// frame of an alignment (lane 0 writes it; callers __syncwarp before the first stripe)
__device__ __forceinline__ void rel_frame_init(RelFrame* frame, const GraphView& gv, const uint32_t* plan, uint8_t* slot,
                                               const uint8_t* seq, uint32_t V, uint32_t L, int sm, int sx) {
    constexpr int NW = DP_NW16;
    const uint32_t NS = Geo<NW, true>::stripes(L);
    frame->meta0 = (unsigned long long)(uintptr_t)gv.meta0;
    frame->pred_off = (unsigned long long)(uintptr_t)gv.pred_off;
    frame->pred_rank = (unsigned long long)(uintptr_t)gv.pred_rank;
    frame->plan = (unsigned long long)(uintptr_t)plan;
    frame->seq = (unsigned long long)(uintptr_t)seq;
    frame->H = (unsigned long long)(uintptr_t)slot;
    frame->bases = (unsigned long long)(uintptr_t)(slot + (uint64_t)(V + 1) * NS * NW * 128);
    frame->V = V; frame->L = L; frame->NS = NS; frame->sm = sm; frame->sx = sx; frame->pad = 0;
}

__device__ __forceinline__ void rel_state_init(RelState& S, uint32_t* prof, RelFrame* frame, uint32_t NS, int gap, int lane) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    S.lane = lane;
    S.g2 = pack2(gap);
    S.pf_lane = (uint32_t)__cvta_generic_to_shared(prof + lane * 4);
    S.frame = (uint32_t)__cvta_generic_to_shared(frame);
    S.row_bytes = NS * (uint32_t)(G::UNITS * 32 * 16);
    S.b = 0; S.bprev = 0; S.bco = 0; S.stripe = 0; S.dst = 0; S.has_prev = false; S.has_next = false;
    asm volatile("" : "+r"(S.pf_lane), "+r"(S.frame), "+r"(S.g2), "+r"(S.row_bytes));
}

// The fill of one alignment by one warp (TEAM false) or by the warps of a team (warp trank takes stripes trank, trank + tsize,
// ...; its progress word vprog[trank] = stripe * (V + 1) + rows done + 1 is cleared by the caller before every alignment).
template <bool TEAM>
__device__ __noinline__ bool dp_fill_rel(const GraphView& gv, const uint32_t* plan, uint8_t* slot, uint8_t* wsm,
                                         const uint8_t* seq, uint32_t V, uint32_t L, int sm, int sx, int gap, int lane,
                                         uint32_t trank, uint32_t tsize, volatile uint32_t* vprog) {
    constexpr int NW = DP_NW16;
    using G = Geo<NW, true>;
    uint32_t* prof = reinterpret_cast<uint32_t*>(wsm);
    RelFrame* frame = reinterpret_cast<RelFrame*>(wsm + G::PROF_BYTES);
    const uint32_t NS = G::stripes(L);
    __syncwarp();
    if (lane == 0) rel_frame_init(frame, gv, plan, slot, seq, V, L, sm, sx);
    RelState S;
    rel_state_init(S, prof, frame, NS, gap, lane);
    bool ok = true;
    for (uint32_t s = TEAM ? trank : 0u; s < NS; s += TEAM ? tsize : 1u) {
        TeamSync ts{vprog, trank, TEAM ? (trank + tsize - 1) % tsize : 0u, s * (V + 1), s > 0 ? (s - 1) * (V + 1) : 0u};
        ok = rel_stripe<TEAM>(S, prof, frame, s, lane, ts) && ok;
    }
    __threadfence_block();
    __syncwarp();
    return ok;
}

}  // namespace hgpu
