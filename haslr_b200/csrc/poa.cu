// Host side of the batched POA path: edge scheduler, memory plan, C ABI (hgpu_poa_*).
//
// Replaces the pthread edge queue of the reference (src/haslr_assemble/src/Assemble.cpp:365-434,562-605):
// instead of T threads pulling one edge each from a mutex-guarded cursor and calling SPOA, all edges of a call
// are queued on the device and pulled by the warps of one persistent kernel (poa_device.cuh).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "poa_device.cuh"

using namespace hgpu;

struct PoaPass {                 // one scheduling pass keeps its consensus pool alive until the results are gathered
    DevBuf<uint8_t> pool;
};

struct PoaState {
    DevBuf<uint8_t> arena, ws, arena_team, ws_team, d_bases, d_out;
    DevBuf<uint64_t> seg_ptr, cons_pos, d_off;
    DevBuf<uint32_t> seg_len, e_seg_off, items, status, cons_len, out_nodes, counters, pool_tab32, pool_first;
    DevBuf<uint8_t> pool_tab8;
    DevBuf<unsigned long long> stats, pool_cursor;
    std::vector<std::unique_ptr<PoaPass>> passes;
    std::vector<std::unique_ptr<PoaPass>> spare;   // consensus pools of earlier calls, reused (a cudaMalloc/cudaFree pair per call costs tens of ms)
    hgpu_poa_stats st{};
    bool timing = false;
    uint64_t cfg_arena_bytes = 0;
    uint32_t cfg_max_warps = 0;
    uint32_t last_n_edges = 0;
    uint64_t last_total = 0;
    WsLayout last_wl{};          // layout / slot size of the most recent launch (debug inspection)
    uint64_t last_slot = 0;
    bool have_result = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t stream2 = nullptr;   // the team kernel (few huge edges) runs beside the warp-per-edge kernel
    static constexpr int NCS = 16;    // size classes whose arenas fit the budget together run side by side on these
    cudaStream_t cstream[NCS] = {};
    cudaEvent_t cev[NCS] = {};
    uint32_t cfg_team = 8;            // 0 disables the team kernel
    double cfg_team_min_cells = 2.0e8;
    uint32_t cfg_teams_per_sm = 2;    // resident teams per SM (HGPU_TEAMS_PER_SM); the rest of the SM runs warp-per-edge blocks
    double cfg_budget_frac = 0.88;    // share of the free device memory the arenas may take (HGPU_BUDGET_FRAC)
    double cfg_team_alpha = 1.5;      // an edge goes to the team kernel when it holds more than alpha x (batch cells / busy warps) (HGPU_TEAM_ALPHA)
    uint32_t cfg_deep_min_reads = 10; // edges with at least this many supporting reads run in k_poa_edges_deep (HGPU_DEEP_MIN_READS)
    int cfg_force = 0;                // HGPU_FORCE_MODE: 1 = every alignment in int32, 2 = every alignment in REL16 (tests)
    int cfg_pool = 1;                 // HGPU_POOL=0: deep edges in the warp-per-edge / block-per-edge kernels instead of k_poa_pool (A/B runs, tests)
    double cfg_pool_chain = 40.0;     // HGPU_POOL_CHAIN: a row of an edge's serial chain (fill + traceback + graph update + sort, one after the other) takes as long
                                      // as this many row-stripe units of a busy block (config 2: 1.0 us against 25 ns); 0 = no edge is protected
    double cfg_pool_penalty = 0.8;    // HGPU_POOL_PENALTY: the block that gets a protected edge is dealt this share of a block's work less
    double cfg_pool_prot = 0.65;      // HGPU_POOL_PROT: an edge is protected when its chain alone is more than this share of the kernel's expected time (on a busy
                                      // block it would end with or after everything else); config 2, (prot, penalty) -> K3: off 341-343 ms, (0.72, 0.5) 334,
                                      // (0.72, 0.8) 331, (0.65, 0.5) 329 ms: profiles/r2L_protected_edges.log
    uint32_t cfg_pool_ctx = 0;        // HGPU_POOL_CTX: contexts (edges in flight) per pool block, 0 = by the stripes per alignment of the class
    int verbose = 0;                  // HGPU_VERBOSE=1: pass / class plan and per-launch device time on stderr; 2: + when k_poa_pool finished which edge
    DevBuf<unsigned long long> edge_clk;
};

// Grow one of the big per-context buffers. The memory budget counts what the state already holds as reusable, so when the
// allocation fails the other arenas (sized by an earlier, differently shaped call) are given back and it is tried again.
static cudaError_t ensure_big(PoaState* S, DevBuf<uint8_t>& buf, size_t bytes) {
    if (bytes <= buf.n) return cudaSuccess;
    cudaError_t e = buf.alloc(bytes);
    if (e == cudaErrorMemoryAllocation) {
        (void)cudaGetLastError();
        DevBuf<uint8_t>* all[4] = {&S->arena, &S->ws, &S->arena_team, &S->ws_team};
        for (DevBuf<uint8_t>* b : all) if (b != &buf) b->release();
        e = buf.alloc(bytes);
    }
    return e;
}

void poa_state_destroy(PoaState* s) {
    if (!s) return;
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    for (int i = 0; i < PoaState::NCS; ++i) { if (s->cstream[i]) cudaStreamDestroy(s->cstream[i]); if (s->cev[i]) cudaEventDestroy(s->cev[i]); }
    delete s;
}

static PoaState* poa_state(hgpu_t* ctx) {
    if (!ctx->poa) {
        ctx->poa = new PoaState();
        // developer knobs (tests force the team kernel onto small inputs with these)
        if (const char* e = getenv("HGPU_TEAM")) ctx->poa->cfg_team = (uint32_t)atoi(e);
        if (const char* e = getenv("HGPU_TEAM_MIN_CELLS")) ctx->poa->cfg_team_min_cells = atof(e);
        if (const char* e = getenv("HGPU_VERBOSE")) ctx->poa->verbose = atoi(e);
        if (const char* e = getenv("HGPU_TEAM_ALPHA")) ctx->poa->cfg_team_alpha = std::max(0.01, atof(e));
        if (const char* e = getenv("HGPU_TEAMS_PER_SM")) ctx->poa->cfg_teams_per_sm = (uint32_t)std::max(1, std::min(4, atoi(e)));
        if (const char* e = getenv("HGPU_BUDGET_FRAC")) ctx->poa->cfg_budget_frac = std::max(0.1, std::min(0.92, atof(e)));
        if (const char* e = getenv("HGPU_FORCE_MODE")) ctx->poa->cfg_force = atoi(e);
        if (const char* e = getenv("HGPU_DEEP_MIN_READS")) ctx->poa->cfg_deep_min_reads = (uint32_t)atoi(e);
        if (const char* e = getenv("HGPU_POOL")) ctx->poa->cfg_pool = atoi(e);
        if (const char* e = getenv("HGPU_POOL_CHAIN")) ctx->poa->cfg_pool_chain = std::max(0.0, atof(e));
        if (const char* e = getenv("HGPU_POOL_PENALTY")) ctx->poa->cfg_pool_penalty = std::max(0.0, atof(e));
        if (const char* e = getenv("HGPU_POOL_PROT")) ctx->poa->cfg_pool_prot = std::max(0.01, atof(e));
        if (const char* e = getenv("HGPU_POOL_CTX")) ctx->poa->cfg_pool_ctx = (uint32_t)std::max(0, std::min((int)POOL_MAX_CTX, atoi(e)));
    }
    return ctx->poa;
}

static int make_scores(hgpu_t* ctx, int8_t match, int8_t mismatch, int8_t gap, DpScores* sc) {
    sc->sm = (int)match - (int)gap;
    sc->sx = (int)mismatch - (int)gap;
    sc->g = gap;
    int lo = std::min(std::min(sc->sm, sc->sx), sc->g);
    sc->lo_step = lo < 0 ? -lo : 0;
    int hi = std::max(sc->sm, sc->sx);
    sc->hi_step = hi > 0 ? hi : 0;
    // the packed int16 fill assumes a non-positive gap and small scores; anything else runs in int32
    if (gap > 0 || std::abs(sc->sm) > 127 || std::abs(sc->sx) > 127) sc->hi_step = 1 << 20;
    (void)ctx;
    return HGPU_OK;
}

namespace {
struct EdgeEst {
    uint32_t edge;
    uint32_t ncap;       // node capacity needed (estimate)
    uint64_t slot;       // score-matrix bytes needed (estimate)
    double cells;        // DP cells (estimate)
    double work;         // time estimate in "row-stripe units": per alignment (V + 1) x (stripes + POOL_GRAPH_ROWS): the fill computes whole
                         // 512-column stripes whatever the gap is, and the serial graph work per node costs about as much as three of them
    double chain;        // rows of the edge's serial chain, sum over its alignments of (V + 1): what one warp after the other has to walk, however
                         // many warps the stripes of an alignment spread over
    uint32_t lmax;       // longest segment
    bool deep;           // many supporting reads: the graph gets several times wider than the gap (k_poa_edges_deep)
};

// Node-count growth model: every later segment adds about `growth` new nodes per base (SURVEY.md §8(d):
// |V| grows ~ L * (ins + sub) per read). growth >= 1 means the worst case (every base a new node).
static constexpr double POOL_GRAPH_ROWS = 3.0;
void estimate_edge(const uint32_t* len, uint32_t R, double growth, const DpScores& sc, int force, EdgeEst* out) {
    double V = len[0], cells = 0, work = 0, chain = 0;
    uint64_t slot = 0;
    uint32_t lmax = len[0];
    for (uint32_t k = 1; k < R; ++k) {
        uint32_t Vi = (uint32_t)std::min<double>(V + 1.0, 4.0e9);
        const int mode = dp_mode(Vi, len[k], sc, force);
        if (mode == DPM_REL16) out->deep = true;                     // only k_poa_edges_deep carries the REL16 code
        slot = std::max(slot, dp_slot_bytes(Vi, len[k], mode));
        cells += (V + 1.0) * (len[k] + 1.0);
        work += (V + 1.0) * ((double)Geo<DP_NW16, true>::stripes(len[k]) + POOL_GRAPH_ROWS);
        chain += V + 1.0;
        // overhang beyond the graph's current span also becomes new nodes
        double over = len[k] > V ? (double)len[k] - V : 0.0;
        V += growth >= 1.0 ? (double)len[k] : std::min<double>(len[k], growth * len[k] + over + 8.0);
        lmax = std::max(lmax, len[k]);
    }
    double ncap = V + lmax + 64.0;
    out->ncap = (uint32_t)std::min<double>(ncap, 4.0e9);
    out->slot = slot + 4096;
    out->cells = cells;
    out->work = work;
    out->chain = chain;
    out->lmax = lmax;
}
}  // namespace

struct PoaRunOpts {
    uint32_t stop_round = 0xFFFFFFFFu;
    int force_i32 = 0;
    uint32_t max_warps = 0;
};

// Runs all edges to consensus. d_bases is a device pointer; seg_off / edge_seg_off are host arrays.
// Leaves per-edge status/cons_len/cons_pos on the device and in `status_h`/`len_h`.
static int poa_run(hgpu_t* ctx, const uint8_t* d_bases, const uint64_t* seg_off, const uint32_t* edge_seg_off, uint32_t n_edges,
                   const DpScores& sc, const PoaRunOpts& opt, std::vector<uint32_t>& status_h, std::vector<uint32_t>& len_h) {
    bool pool_ran = false; (void)pool_ran;
    PoaState* S = poa_state(ctx);
    cudaStream_t st = ctx->stream;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    for (auto& p : S->passes) if (S->spare.size() < 4) S->spare.push_back(std::move(p));
    S->passes.clear();
    S->have_result = false;
    memset(&S->st, 0, sizeof S->st);

    // ---- non-empty segments, CSR by edge (Assemble.cpp:537 skips empty ones)
    std::vector<uint64_t> seg_ptr; std::vector<uint32_t> seg_len; std::vector<uint32_t> e_off(n_edges + 1, 0);
    seg_ptr.reserve(edge_seg_off[n_edges]); seg_len.reserve(edge_seg_off[n_edges]);
    for (uint32_t e = 0; e < n_edges; ++e) {
        e_off[e] = (uint32_t)seg_len.size();
        for (uint32_t s = edge_seg_off[e]; s < edge_seg_off[e + 1]; ++s) {
            uint64_t b = seg_off[s], en = seg_off[s + 1];
            if (en < b || en - b > 0x7FFFFFFFull) HGPU_FAIL(ctx, HGPU_E_INVALID, "segment %u has invalid bounds", s);
            if (en > b) { seg_ptr.push_back(b); seg_len.push_back((uint32_t)(en - b)); }
        }
    }
    e_off[n_edges] = (uint32_t)seg_len.size();
    const size_t n_seg = seg_len.size();

    HGPU_CUDA(ctx, S->seg_ptr.ensure(n_seg + 1)); HGPU_CUDA(ctx, S->seg_len.ensure(n_seg + 1));
    HGPU_CUDA(ctx, S->e_seg_off.ensure(n_edges + 1)); HGPU_CUDA(ctx, S->items.ensure(n_edges + 1));
    HGPU_CUDA(ctx, S->status.ensure(n_edges + 1)); HGPU_CUDA(ctx, S->cons_len.ensure(n_edges + 1));
    HGPU_CUDA(ctx, S->cons_pos.ensure(n_edges + 1)); HGPU_CUDA(ctx, S->out_nodes.ensure(n_edges + 1));
    HGPU_CUDA(ctx, S->stats.ensure(32)); HGPU_CUDA(ctx, S->pool_cursor.ensure(1)); HGPU_CUDA(ctx, S->counters.ensure(256));
    if (n_seg) {
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->seg_ptr.p, seg_ptr.data(), n_seg * 8, cudaMemcpyHostToDevice, st));
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->seg_len.p, seg_len.data(), n_seg * 4, cudaMemcpyHostToDevice, st));
    }
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->e_seg_off.p, e_off.data(), (size_t)(n_edges + 1) * 4, cudaMemcpyHostToDevice, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->stats.p, 0, 8 * sizeof(unsigned long long), st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->status.p, 0, (size_t)(n_edges + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->cons_len.p, 0, (size_t)(n_edges + 1) * 4, st));
    HGPU_CUDA(ctx, cudaMemsetAsync(S->cons_pos.p, 0, (size_t)(n_edges + 1) * 8, st));

    status_h.assign(n_edges, 0); len_h.assign(n_edges, 0);
    if (n_edges == 0) return HGPU_OK;

    // ---- resident warps and memory budget
    int blocks_per_sm = 0;
    const size_t smem = (size_t)DP_WARPS_PER_BLOCK * DP_SMEM_PER_WARP;
    HGPU_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_poa_edges, 32 * DP_WARPS_PER_BLOCK, smem));
    if (blocks_per_sm < 1) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "k_poa_edges cannot be resident");
    uint32_t max_warps = (uint32_t)ctx->sm_count * blocks_per_sm * DP_WARPS_PER_BLOCK;
    if (S->cfg_max_warps) max_warps = std::min(max_warps, S->cfg_max_warps);
    if (opt.max_warps) max_warps = std::min(max_warps, opt.max_warps);
    max_warps = std::max<uint32_t>(DP_WARPS_PER_BLOCK, max_warps / DP_WARPS_PER_BLOCK * DP_WARPS_PER_BLOCK);
    // the deep-edge kernel: a ring of parked rows per warp, fewer resident warps
    int blocks_per_sm_deep = 0;
    const size_t smem_deep = (size_t)DP_WARPS_PER_BLOCK * DP_SMEM_PER_WARP_DEEP;
    HGPU_CUDA(ctx, cudaFuncSetAttribute(k_poa_edges_deep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_deep));
    HGPU_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_deep, k_poa_edges_deep, 32 * DP_WARPS_PER_BLOCK, smem_deep));
    if (blocks_per_sm_deep < 1) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "k_poa_edges_deep cannot be resident");
    uint32_t max_warps_deep = std::min<uint32_t>(max_warps, (uint32_t)ctx->sm_count * blocks_per_sm_deep * DP_WARPS_PER_BLOCK);
    uint64_t budget = S->cfg_arena_bytes;
    if (!budget) {
        size_t fr = 0, tot = 0;
        HGPU_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
        fr += S->arena.n + S->ws.n + S->arena_team.n + S->ws_team.n;   // what we already hold can be reused (all of it: a budget that
                                                                          // shrinks from call to call changes the wave plan of repeated calls)
        budget = (uint64_t)(fr * S->cfg_budget_frac);
        budget = std::min<uint64_t>(budget, 165ull << 30);
    }

    std::vector<uint32_t> pending(n_edges);
    std::iota(pending.begin(), pending.end(), 0u);
    double growth = 0.10;             // new nodes per base per read, first guess (measured 0.05-0.07 at 9 % read error); retried x2.5 when an edge outgrows it
    std::vector<EdgeEst> est;
    std::vector<uint32_t> items_h;
    const uint64_t budget_total = budget;
    for (int attempt = 0; attempt < 5 && !pending.empty(); ++attempt) {
        budget = budget_total;
        // ---- estimates, largest first (also the LPT order for load balance)
        est.resize(pending.size());
        uint64_t pool_cap = 0;
        for (size_t i = 0; i < pending.size(); ++i) {
            uint32_t e = pending[i];
            uint32_t R = e_off[e + 1] - e_off[e];
            est[i].edge = e;
            est[i].deep = R >= S->cfg_deep_min_reads;
            if (R == 0) { est[i].ncap = 64; est[i].slot = 4096; est[i].cells = 0; est[i].work = 0; est[i].chain = 0; est[i].lmax = 0; continue; }
            estimate_edge(seg_len.data() + e_off[e], R, growth, sc, opt.force_i32, &est[i]);
            uint64_t sum = 0; uint32_t lmax = 0;
            for (uint32_t k = 0; k < R; ++k) { sum += seg_len[e_off[e] + k]; lmax = std::max(lmax, seg_len[e_off[e] + k]); }
            pool_cap += growth >= 1.0 ? sum : std::min<uint64_t>(sum, (uint64_t)(2.0 * lmax * (1.0 + growth)) + 256);
        }
        // ---- edges so large that one warp would be the tail of the whole pass go to the team kernel (a block per edge):
        //      a lone warp on a deep graph sustains ~0.4 G cells/s (one dependent instruction stream, IPC 0.15: profiles/r1i_*)
        //      against ~1 T cells/s for the device
        const bool use_pool = S->cfg_pool != 0 && opt.stop_round == 0xFFFFFFFFu;
        std::vector<EdgeEst> team;
        if (use_pool) {
            // the pool kernel spreads the stripes of an alignment over the warps of its block: an edge that would be the tail of the
            // warp-per-edge kernel (more than alpha x the average share of a warp, several stripes wide) goes there whatever its depth
            double total = 0;
            for (const EdgeEst& x : est) total += x.cells;
            const double share = total / (double)std::max<size_t>(1, std::min<size_t>(est.size(), max_warps));
            const double thr = std::max(S->cfg_team_alpha * share, S->cfg_team_min_cells);
            double shallow = 0;
            for (EdgeEst& x : est) {
                if (!x.deep && x.lmax >= 2u * (uint32_t)Geo<DP_NW16, true>::SW - 1 && x.cells > thr) x.deep = true;
                if (!x.deep) shallow += x.cells;
            }
            // a small shallow remainder (config 2: 48 of 6,033 edges) would only wait for the pool blocks to leave the SMs: it joins them
            if (shallow < 0.10 * total) for (EdgeEst& x : est) x.deep = true;
        } else if (S->cfg_team >= 2 && opt.stop_round == 0xFFFFFFFFu) {
            double total = 0;
            for (const EdgeEst& x : est) total += x.cells;
            // (a) fewer edges than resident teams: the device is mostly empty, every wide edge gets a team (3.3x faster per edge);
            // (b) otherwise only the tail: with W warps busy an edge is "tail" when it holds more than cfg_team_alpha times the
            //     average share of a warp, total / W (alpha 1.5 ~ 1/3000 of the batch on BASELINE config 2, the swept optimum)
            const bool few = est.size() <= (size_t)ctx->sm_count * S->cfg_teams_per_sm;
            const double share = total / (double)std::max<size_t>(1, std::min<size_t>(est.size(), max_warps));
            const double thr = few ? S->cfg_team_min_cells / 8.0 : std::max(S->cfg_team_alpha * share, S->cfg_team_min_cells);
            std::vector<EdgeEst> rest;
            for (const EdgeEst& x : est) {
                const bool wide = x.lmax >= 2u * (uint32_t)Geo<DP_NW16, true>::SW - 1;     // at least 3 stripes to spread
                if (wide && x.cells > thr) team.push_back(x); else rest.push_back(x);
            }
            est.swap(rest);
        }
        auto by_size = [](const EdgeEst& a, const EdgeEst& b) {      // deep edges first, then by slot size, largest first
            if (a.deep != b.deep) return a.deep;
            if (a.slot != b.slot) return a.slot > b.slot;
            return a.edge < b.edge;
        };
        std::sort(est.begin(), est.end(), by_size);
        std::sort(team.begin(), team.end(), [](const EdgeEst& a, const EdgeEst& b) { return a.cells != b.cells ? a.cells > b.cells : a.edge < b.edge; });
        items_h.resize(est.size() + team.size());
        for (size_t i = 0; i < est.size(); ++i) items_h[i] = est[i].edge;
        for (size_t i = 0; i < team.size(); ++i) items_h[est.size() + i] = team[i].edge;
        HGPU_CUDA(ctx, cudaMemcpyAsync(S->items.p, items_h.data(), items_h.size() * 4, cudaMemcpyHostToDevice, st));

        std::unique_ptr<PoaPass> pass;
        for (size_t q = 0; q < S->spare.size(); ++q)
            if (S->spare[q]->pool.n >= pool_cap + 256 && S->spare[q]->pool.n <= 2 * (pool_cap + 256) + (1u << 20)) {
                pass = std::move(S->spare[q]); S->spare.erase(S->spare.begin() + q); break;
            }
        if (!pass) { pass.reset(new PoaPass()); HGPU_CUDA(ctx, pass->pool.alloc(pool_cap + 256)); }
        HGPU_CUDA(ctx, cudaMemsetAsync(S->pool_cursor.p, 0, sizeof(unsigned long long), st));
        HGPU_CUDA(ctx, cudaMemsetAsync(S->counters.p, 0, 256 * 4, st));

        if (S->verbose) fprintf(stderr, "[poa] host: %.1f ms to the first launch of attempt %d\n", host_ms(), attempt);
        if (S->timing) HGPU_CUDA(ctx, cudaEventRecord(S->ev0, st));
        bool team_launched = false;
        if (!team.empty()) {
            const int TEAM = S->cfg_team >= 8 ? 8 : 4;
            const uint32_t teams_per_sm = std::min<uint32_t>(S->cfg_teams_per_sm, TEAM == 8 ? 2u : 4u);
            uint64_t tslot = 0; uint32_t nc = 64;
            for (const EdgeEst& x : team) { tslot = std::max(tslot, x.slot); nc = std::max(nc, x.ncap); }
            tslot = (tslot + 127) / 128 * 128;
            const WsLayout twl = ws_layout(nc, nc + nc / 4 + 64);
            const uint64_t tbudget = budget / 2;              // the other half stays with the warp-per-edge kernel
            uint32_t teams = (uint32_t)std::min<uint64_t>({(uint64_t)team.size(), (uint64_t)ctx->sm_count * teams_per_sm, tbudget / (tslot + twl.bytes)});
            if (teams == 0) {
                for (const EdgeEst& x : team) est.push_back(x);          // does not fit even once: let the classes report it
                std::sort(est.begin(), est.end(), by_size);
                for (size_t i = 0; i < est.size(); ++i) items_h[i] = est[i].edge;
                HGPU_CUDA(ctx, cudaMemcpyAsync(S->items.p, items_h.data(), items_h.size() * 4, cudaMemcpyHostToDevice, st));
                team.clear();
            } else {
                budget -= (uint64_t)teams * (tslot + twl.bytes);
                HGPU_CUDA(ctx, S->arena_team.ensure((size_t)teams * tslot));
                HGPU_CUDA(ctx, S->ws_team.ensure((size_t)teams * twl.bytes));
                PoaArgs a{};
                a.bases = d_bases; a.seg_ptr = S->seg_ptr.p; a.seg_len = S->seg_len.p; a.e_seg_off = S->e_seg_off.p;
                a.items = S->items.p + est.size(); a.n_items = (uint32_t)team.size(); a.counter = S->counters.p + 255;
                a.status = S->status.p; a.cons_len = S->cons_len.p; a.cons_pos = S->cons_pos.p; a.out_nodes = S->out_nodes.p;
                a.pool = pass->pool.p; a.pool_cap = pass->pool.n; a.pool_cursor = S->pool_cursor.p;
                a.ws = S->ws_team.p; a.wl = twl; a.arena = S->arena_team.p; a.slot_bytes = tslot;
                a.sc = sc; a.stats = S->stats.p; a.stop_round = opt.stop_round; a.force_i32 = opt.force_i32;
                const size_t tsmem = (size_t)TEAM * DP_SMEM_PER_WARP_DEEP + (TEAM + 4) * 4;
                HGPU_CUDA(ctx, cudaEventRecord(S->ev_fork, st));          // uploads and memsets above are on `st`
                HGPU_CUDA(ctx, cudaStreamWaitEvent(S->stream2, S->ev_fork, 0));
                if (TEAM == 8) {
                    HGPU_CUDA(ctx, cudaFuncSetAttribute(k_poa_edges_team<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                    k_poa_edges_team<8><<<teams, 32 * 8, tsmem, S->stream2>>>(a);
                } else {
                    HGPU_CUDA(ctx, cudaFuncSetAttribute(k_poa_edges_team<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                    k_poa_edges_team<4><<<teams, 32 * 4, tsmem, S->stream2>>>(a);
                }
                HGPU_CUDA(ctx, cudaGetLastError());
                HGPU_CUDA(ctx, cudaEventRecord(S->ev_join, S->stream2));
                ctx->launches++; S->st.dp_launches++;
                team_launched = true;
                if (S->verbose) {
                    double tc = 0; for (const EdgeEst& x : team) tc += x.cells;
                    fprintf(stderr, "[poa] attempt %d growth %.2f: team kernel %zu edges on %u teams, slot %.1f MB, ncap %u, %.3e cells (largest %.3e)\n",
                            attempt, growth, team.size(), teams, tslot / 1048576.0, nc, tc, team[0].cells);
                }
            }
        }

        // ---- size classes: a class ends where the slot estimate has halved, unless memory is no constraint
        struct Cls { size_t a, b; uint64_t slot; WsLayout wl; uint32_t warps; bool deep; bool pool; uint32_t ctx_per_block, pool_blocks; double cells, work; };
        const uint32_t pool_blocks_max = (uint32_t)HGPU_POOL_BLOCKS_PER_SM * (uint32_t)ctx->sm_count;   // k_poa_pool: HGPU_POOL_BLOCKS_PER_SM blocks of POOL_WARPS warps per SM (1 x 16)
        std::vector<Cls> classes;
        constexpr uint64_t POOL_SMALL_SLOT = 8ull << 20;
        uint32_t n_pool_classes = 0;
        size_t n_deep = 0;
        while (n_deep < est.size() && est[n_deep].deep) ++n_deep;
        size_t i = 0;
        while (i < est.size()) {
            const bool deep = i < n_deep;
            const size_t end = deep ? n_deep : est.size();           // a class never mixes the two kernels
            const bool pool = deep && use_pool;
            const uint32_t mw = pool ? pool_blocks_max * (uint32_t)POOL_MAX_CTX : deep ? max_warps_deep : max_warps;   // pool: slots = contexts
            Cls c; c.a = i; c.slot = (est[i].slot + 127) / 128 * 128; c.deep = deep; c.pool = pool; c.ctx_per_block = 0; c.pool_blocks = 0; c.cells = 0; c.work = 0;
            auto plan = [&](size_t a, size_t b, Cls& cc) {
                uint32_t nc = 64;
                for (size_t q = a; q < b; ++q) nc = std::max(nc, est[q].ncap);
                uint32_t ec = nc + nc / 4 + 64;
                cc.wl = ws_layout(nc, ec);
                uint64_t per = cc.slot + cc.wl.bytes;
                uint64_t w = budget / per;
                cc.warps = (uint32_t)std::min<uint64_t>(w, mw);
            };
            plan(i, end, c);
            size_t j = end;
            // pool contexts are sized per class: split where the slot estimate has halved - but not below POOL_SMALL_SLOT (small slots
            // waste little when shared) and into at most POOL_MAX_CLASSES classes (the kernel's parameter block holds that many)
            const bool split = pool ? (c.slot > POOL_SMALL_SLOT && n_pool_classes + 1 < (uint32_t)POOL_MAX_CLASSES) : c.warps < mw;
            if (split) {
                j = i + 1;
                while (j < end && est[j].slot * 2 > c.slot) ++j;
                plan(i, j, c);
            }
            if (pool) ++n_pool_classes;
            c.b = j;
            for (size_t q = c.a; q < c.b; ++q) { c.cells += est[q].cells; c.work += est[q].work; }
            classes.push_back(c);
            i = j;
        }
        if (classes.size() > 256) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "too many size classes (%zu)", classes.size());

        // ---- pool classes: ONE persistent kernel for all of them. Every block holds contexts of several classes (a context = a
        //      slot + workspace of its class's size), dealt so that each class gets contexts in proportion to its estimated time
        //      (EdgeEst::work) and every SM sees the same mix; a context whose class has run dry takes edges of the classes with
        //      smaller slots. So the classes cannot finish at different times any more (they did: 367 ... 611 ms on config 2 as six
        //      side-by-side kernels), and the few huge edges get the warps the small ones leave.
        bool pool_launched = false;
        {
            std::vector<size_t> pc;
            for (size_t ci = 0; ci < classes.size(); ++ci) if (classes[ci].pool && classes[ci].warps) pc.push_back(ci);
            if (!pc.empty()) {
                const size_t K = pc.size();
                double tot_work = 0; uint64_t tot_items = 0;
                for (size_t k = 0; k < K; ++k) { tot_work += classes[pc[k]].work + 1.0; tot_items += classes[pc[k]].b - classes[pc[k]].a; }
                // a context of class k may run edges of later classes: its workspace must hold their graphs too
                for (size_t k = K; k-- > 1;)
                    if (classes[pc[k]].wl.ncap > classes[pc[k - 1]].wl.ncap) classes[pc[k - 1]].wl = ws_layout(classes[pc[k]].wl.ncap, classes[pc[k]].wl.ecap);
                const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(pool_blocks_max, tot_items));
                uint32_t E = S->cfg_pool_ctx ? S->cfg_pool_ctx : (uint32_t)std::min<uint64_t>(POOL_MAX_CTX, (tot_items + blocks - 1) / blocks);
                E = std::max<uint32_t>(1, std::min<uint32_t>(E, (uint32_t)POOL_MAX_CTX));
                const uint32_t T = blocks * E;
                // quotas: proportional to work, at least one, at most the class's edges; then within the memory budget
                std::vector<uint32_t> quota(K);
                uint32_t given = 0;
                for (size_t k = 0; k < K; ++k) {
                    const uint32_t n_items = (uint32_t)(classes[pc[k]].b - classes[pc[k]].a);
                    uint32_t qk = (uint32_t)std::lround(T * (classes[pc[k]].work + 1.0) / tot_work);
                    quota[k] = std::max<uint32_t>(1, std::min<uint32_t>({qk, n_items, classes[pc[k]].warps}));
                    given += quota[k];
                }
                for (size_t k = 0; given > T && k < 64 * K; ++k) {            // rounding / minimums may overshoot: trim the largest quotas
                    size_t w = 0;
                    for (size_t q = 1; q < K; ++q) if (quota[q] > quota[w]) w = q;
                    if (quota[w] <= 1) break;
                    --quota[w]; --given;
                }
                for (bool grew = true; given < T && grew;) {                 // contexts left over: classes that still have more edges than contexts
                    grew = false;
                    for (size_t k = 0; k < K && given < T; ++k) {
                        const uint32_t n_items = (uint32_t)(classes[pc[k]].b - classes[pc[k]].a);
                        if (quota[k] < n_items && quota[k] < classes[pc[k]].warps) { ++quota[k]; ++given; grew = true; }
                    }
                }
                auto mem_of = [&](size_t k) { return (uint64_t)quota[k] * (classes[pc[k]].slot + classes[pc[k]].wl.bytes); };
                uint64_t mem = 0;
                for (size_t k = 0; k < K; ++k) mem += mem_of(k);
                for (int guard = 0; mem > budget && guard < 100000; ++guard) {
                    size_t w = K;
                    for (size_t q = 0; q < K; ++q) if (quota[q] > 1 && (w == K || mem_of(q) > mem_of(w))) w = q;
                    if (w == K) break;
                    mem -= classes[pc[w]].slot + classes[pc[w]].wl.bytes; --quota[w]; --given;
                }
                // deal the contexts: context g (block g / E) goes to the class furthest behind its quota, so every block gets the mix
                std::vector<uint8_t> ctx_class(T, 0xFF);
                std::vector<uint32_t> ctx_slot(T, 0), dealt(K, 0);
                for (uint32_t g = 0; g < given && g < T; ++g) {
                    // spread over blocks first: context g of the deal sits in block g % blocks, position g / blocks
                    const uint32_t where = (g % blocks) * E + g / blocks;
                    size_t best = K; double lag = -1e300;
                    for (size_t k = 0; k < K; ++k) {
                        if (dealt[k] >= quota[k]) continue;
                        const double l = (double)(g + 1) * quota[k] / given - dealt[k];
                        if (l > lag) { lag = l; best = k; }
                    }
                    if (best == K) break;
                    ctx_class[where] = (uint8_t)best; ctx_slot[where] = dealt[best]++;
                }
                // first edges. Every class has its edges sorted by estimated time; the quota[k] heaviest are dealt to its contexts here,
                // heaviest edge of all classes first, each to the block with the least work so far that still has a free context of the
                // class (LPT), so the blocks' heavy work ends together; the rest of the class waits in its queue and goes to whichever
                // context is free first.
                std::vector<uint32_t> ctx_first(T, 0xFFFFFFFFu);
                {
                    for (size_t k = 0; k < K; ++k) {
                        const Cls& c = classes[pc[k]];
                        std::sort(est.begin() + c.a, est.begin() + c.b, [](const EdgeEst& x, const EdgeEst& y) { return x.work != y.work ? x.work > y.work : x.edge < y.edge; });
                        for (size_t q = c.a; q < c.b; ++q) items_h[q] = est[q].edge;
                    }
                    HGPU_CUDA(ctx, cudaMemcpyAsync(S->items.p, items_h.data(), est.size() * 4, cudaMemcpyHostToDevice, st));
                    std::vector<std::vector<std::vector<uint32_t>>> free_pos(K, std::vector<std::vector<uint32_t>>(blocks));
                    for (uint32_t b = 0; b < blocks; ++b)
                        for (uint32_t j = E; j-- > 0;) { const uint32_t w = b * E + j; if (ctx_class[w] != 0xFF) free_pos[ctx_class[w]][b].push_back(w); }
                    struct Cand { double work; uint32_t k; size_t q; };
                    std::vector<Cand> cand;
                    for (size_t k = 0; k < K; ++k)
                        for (uint32_t x = 0; x < dealt[k]; ++x) cand.push_back({est[classes[pc[k]].a + x].work, (uint32_t)k, classes[pc[k]].a + x});
                    std::sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) { return x.work != y.work ? x.work > y.work : x.q < y.q; });
                    // Protected edges. The alignments of one edge are a serial chain (fill, traceback, graph update, sort, next fill); on a
                    // block that is as busy as the others the chain of the heaviest edge of config 2 takes 341 ms (230 ms alone) while the
                    // rest of the kernel is done after 316 ms. An edge whose chain alone is most of the kernel's expected time (cfg_pool_prot)
                    // charges its block with `penalty` x a block's share of the work on top of its own, so the deal gives that block the
                    // lightest first edges of every class: its warps are free for the chain long before the others finish. Few edges
                    // qualify (config 2: 5 of 6,033); protecting every deep edge only moves work to the other blocks (37 blocks charged:
                    // the tail went and everything else ended 26 ms later, profiles/r2K_protected_edges.log). What is left of the tail is
                    // the chain itself: 29 alignments of a graph that grows to ~9,000 nodes, ~11 ms each on an almost empty block.
                    const double block_share = tot_work / blocks;
                    uint32_t n_prot = 0;
                    std::vector<double> block_work(blocks, 0.0);
                    for (const Cand& cd : cand) {
                        uint32_t bb = blocks;
                        for (uint32_t b = 0; b < blocks; ++b)
                            if (!free_pos[cd.k][b].empty() && (bb == blocks || block_work[b] < block_work[bb])) bb = b;
                        if (bb == blocks) break;              // cannot happen: dealt[k] contexts of the class exist
                        ctx_first[free_pos[cd.k][bb].back()] = est[cd.q].edge;
                        free_pos[cd.k][bb].pop_back();
                        block_work[bb] += cd.work;
                        if (S->cfg_pool_chain > 0 && blocks > 1 && n_prot < blocks / 8 && est[cd.q].chain * S->cfg_pool_chain > S->cfg_pool_prot * block_share) {
                            block_work[bb] += S->cfg_pool_penalty * block_share;
                            ++n_prot;
                        }
                    }
                    if (S->verbose) {
                        double lo = 1e300, hi = 0, sum = 0;
                        for (double w : block_work) { lo = std::min(lo, w); hi = std::max(hi, w); sum += w; }
                        fprintf(stderr, "[poa] first edges dealt: estimated work per block min %.3g mean %.3g max %.3g (incl. the charge of %u protected edges)\n", lo, sum / blocks, hi, n_prot);
                    }
                }
                PoolArgs pa{};
                uint64_t a_off = 0, w_off = 0;
                for (size_t k = 0; k < K; ++k) { a_off += (uint64_t)quota[k] * classes[pc[k]].slot; w_off += (uint64_t)quota[k] * classes[pc[k]].wl.bytes; }
                HGPU_CUDA(ctx, ensure_big(S, S->arena_team, a_off)); HGPU_CUDA(ctx, ensure_big(S, S->ws_team, w_off));
                HGPU_CUDA(ctx, S->pool_tab8.ensure(T)); HGPU_CUDA(ctx, S->pool_tab32.ensure(T));
                HGPU_CUDA(ctx, cudaMemcpyAsync(S->pool_tab8.p, ctx_class.data(), T, cudaMemcpyHostToDevice, st));
                HGPU_CUDA(ctx, cudaMemcpyAsync(S->pool_tab32.p, ctx_slot.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
                HGPU_CUDA(ctx, S->pool_first.ensure(T));
                HGPU_CUDA(ctx, cudaMemcpyAsync(S->pool_first.p, ctx_first.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
                S->st.arena_bytes = std::max<uint64_t>(S->st.arena_bytes, a_off);
                budget -= std::min<uint64_t>(budget, a_off + w_off);
                a_off = 0; w_off = 0;
                for (size_t k = 0; k < K; ++k) {
                    const Cls& c = classes[pc[k]];
                    PoolClass& q = pa.cls[k];
                    q.items = S->items.p + c.a + dealt[k]; q.n_items = (uint32_t)(c.b - c.a) - dealt[k]; q.counter = S->counters.p + pc[k];   // the first dealt[k] went out with ctx_first
                    q.ws = S->ws_team.p + w_off; q.wl = c.wl; q.arena = S->arena_team.p + a_off; q.slot_bytes = c.slot;
                    a_off += (uint64_t)quota[k] * c.slot; w_off += (uint64_t)quota[k] * c.wl.bytes;
                    if (S->verbose)
                        fprintf(stderr, "[poa] attempt %d growth %.2f pool class %zu/%zu: %u edges, %u contexts, slot %.1f MB, ws %.1f MB, %.3e cells, work share %.1f %%\n",
                                attempt, growth, k, K, q.n_items, quota[k], c.slot / 1048576.0, c.wl.bytes / 1048576.0, c.cells, 100.0 * (c.work + 1.0) / tot_work);
                }
                pa.n_cls = (uint32_t)K; pa.ctx_class = S->pool_tab8.p; pa.ctx_slot = S->pool_tab32.p; pa.ctx_first = S->pool_first.p;
                PoaArgs& a = pa.a;
                a.bases = d_bases; a.seg_ptr = S->seg_ptr.p; a.seg_len = S->seg_len.p; a.e_seg_off = S->e_seg_off.p;
                a.status = S->status.p; a.cons_len = S->cons_len.p; a.cons_pos = S->cons_pos.p; a.out_nodes = S->out_nodes.p;
                a.pool = pass->pool.p; a.pool_cap = pass->pool.n; a.pool_cursor = S->pool_cursor.p;
                a.sc = sc; a.stats = S->stats.p; a.stop_round = opt.stop_round; a.force_i32 = opt.force_i32;
#if HGPU_PHASE_CLOCKS
                HGPU_CUDA(ctx, cudaMemsetAsync(S->stats.p + 8, 0, 24 * sizeof(unsigned long long), st));
                a.phase_clk = S->stats.p + 8;
                { unsigned long long* tbp = S->stats.p + 8 + 10; HGPU_CUDA(ctx, cudaMemcpyToSymbolAsync(g_tb_counters, &tbp, sizeof tbp, 0, cudaMemcpyHostToDevice, st)); }
#endif
                if (S->verbose >= 2) {
                    HGPU_CUDA(ctx, S->edge_clk.ensure((size_t)n_edges * 2));
                    HGPU_CUDA(ctx, cudaMemsetAsync(S->edge_clk.p, 0, (size_t)n_edges * 16, st));
                    a.edge_clk = S->edge_clk.p;
                }
                const size_t psmem = (size_t)POOL_WARPS * DP_SMEM_PER_WARP_DEEP + sizeof(PoolShared);
                HGPU_CUDA(ctx, cudaFuncSetAttribute(k_poa_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
                HGPU_CUDA(ctx, cudaEventRecord(S->ev_fork, st));          // uploads and memsets above are on `st`
                HGPU_CUDA(ctx, cudaStreamWaitEvent(S->stream2, S->ev_fork, 0));
                k_poa_pool<<<blocks, 32 * POOL_WARPS, psmem, S->stream2>>>(pa, E);
                HGPU_CUDA(ctx, cudaGetLastError());
                HGPU_CUDA(ctx, cudaEventRecord(S->ev_join, S->stream2));
                ctx->launches++; S->st.dp_launches++;
                pool_launched = true; pool_ran = true;
                if (S->verbose) fprintf(stderr, "[poa] k_poa_pool: %u blocks x %u contexts (%u in use), %.1f GB of slots\n", blocks, E, given, (a_off + w_off) / 1073741824.0);
            }
        }

        // ---- launches. A class is one persistent kernel with its own slots and workspaces; consecutive classes whose
        //      memory fits the budget together form a wave and run side by side (each on its own stream), so a class of a
        //      few huge edges does not hold the device alone. Waves follow each other on `st`.
        struct Launch { size_t ci; uint32_t warps, blocks; uint64_t arena_off, ws_off; };
        std::vector<std::vector<Launch>> waves(1);
        uint64_t wave_a = 0, wave_w = 0, arena_need = 0, ws_need = 0;
        for (size_t ci = 0; ci < classes.size(); ++ci) {
            Cls& c = classes[ci];
            const uint32_t n_items = (uint32_t)(c.b - c.a);
            if (c.pool && c.warps) continue;                  // launched above, all pool classes in one kernel
            if (c.warps == 0) {
                // does not fit the device budget even alone: report per edge, keep going
                std::vector<uint32_t> code(1, ST_TOO_LARGE);
                for (size_t q = c.a; q < c.b; ++q)
                    HGPU_CUDA(ctx, cudaMemcpyAsync(S->status.p + est[q].edge, code.data(), 4, cudaMemcpyHostToDevice, st));
                HGPU_CUDA(ctx, cudaStreamSynchronize(st));
                continue;
            }
            uint32_t warps, blocks;
            {
                warps = std::min<uint32_t>(c.warps, n_items ? n_items : 1);
                blocks = (warps + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK;
                warps = blocks * DP_WARPS_PER_BLOCK;
                if ((uint64_t)warps > c.warps) {
                    if (c.warps >= (uint32_t)DP_WARPS_PER_BLOCK) { blocks = c.warps / DP_WARPS_PER_BLOCK; warps = blocks * DP_WARPS_PER_BLOCK; }
                    else {
                        // the budget admits fewer slots than one block has warps: report per edge instead of allocating past the budget
                        std::vector<uint32_t> code(1, ST_TOO_LARGE);
                        for (size_t q = c.a; q < c.b; ++q)
                            HGPU_CUDA(ctx, cudaMemcpyAsync(S->status.p + est[q].edge, code.data(), 4, cudaMemcpyHostToDevice, st));
                        HGPU_CUDA(ctx, cudaStreamSynchronize(st));
                        continue;
                    }
                }
            }
            const uint64_t na = (uint64_t)warps * c.slot, nw = (uint64_t)warps * c.wl.bytes;
            if (!waves.back().empty() && (wave_a + wave_w + na + nw > budget || waves.back().size() >= (size_t)PoaState::NCS)) {
                waves.emplace_back(); wave_a = 0; wave_w = 0;
            }
            waves.back().push_back({ci, warps, blocks, wave_a, wave_w});
            wave_a += na; wave_w += nw;
            arena_need = std::max(arena_need, wave_a); ws_need = std::max(ws_need, wave_w);
        }
        HGPU_CUDA(ctx, ensure_big(S, S->arena, arena_need));
        HGPU_CUDA(ctx, ensure_big(S, S->ws, ws_need));
        S->st.arena_bytes = std::max<uint64_t>(S->st.arena_bytes, arena_need);
        for (size_t wi = 0; wi < waves.size(); ++wi) {
            const std::vector<Launch>& wave = waves[wi];
            if (wave.empty()) continue;
            const bool side_by_side = wave.size() > 1;
            cudaEvent_t vb = nullptr, ve = nullptr;
            std::vector<cudaEvent_t> vclass;                  // verbose: when every class of the wave ended
            if (S->verbose) { cudaEventCreate(&vb); cudaEventCreate(&ve); cudaEventRecord(vb, st); }
            if (side_by_side) HGPU_CUDA(ctx, cudaEventRecord(S->ev_fork, st));
            for (size_t li = 0; li < wave.size(); ++li) {
                const Launch& ln = wave[li];
                Cls& c = classes[ln.ci];
                const uint32_t n_items = (uint32_t)(c.b - c.a);
                cudaStream_t ls = side_by_side ? S->cstream[li] : st;
                if (side_by_side) HGPU_CUDA(ctx, cudaStreamWaitEvent(ls, S->ev_fork, 0));
                if (opt.stop_round != 0xFFFFFFFFu)   // debug inspection looks for the one workspace that holds a graph
                    HGPU_CUDA(ctx, cudaMemsetAsync(S->ws.p + ln.ws_off, 0, (size_t)ln.warps * c.wl.bytes, ls));
                PoaArgs a{};
                a.bases = d_bases; a.seg_ptr = S->seg_ptr.p; a.seg_len = S->seg_len.p; a.e_seg_off = S->e_seg_off.p;
                a.items = S->items.p + c.a; a.n_items = n_items; a.counter = S->counters.p + ln.ci;
                a.status = S->status.p; a.cons_len = S->cons_len.p; a.cons_pos = S->cons_pos.p; a.out_nodes = S->out_nodes.p;
                a.pool = pass->pool.p; a.pool_cap = pass->pool.n; a.pool_cursor = S->pool_cursor.p;
                a.ws = S->ws.p + ln.ws_off; a.wl = c.wl; a.arena = S->arena.p + ln.arena_off; a.slot_bytes = c.slot;
                S->last_wl = c.wl; S->last_slot = c.slot;
                a.sc = sc; a.stats = S->stats.p; a.stop_round = opt.stop_round; a.force_i32 = opt.force_i32;
#if HGPU_PHASE_CLOCKS
                HGPU_CUDA(ctx, cudaMemsetAsync(S->stats.p + 8, 0, 16 * sizeof(unsigned long long), ls));
                a.phase_clk = S->stats.p + 8;
#endif
                if (const char* e = getenv("HGPU_PROBE")) { a.probe = (uint32_t)atoi(e); a.probe_round = 1; }
                if (const char* e = getenv("HGPU_PROBE_ROUND")) a.probe_round = (uint32_t)atoi(e);
                if (c.deep) k_poa_edges_deep<<<ln.blocks, 32 * DP_WARPS_PER_BLOCK, smem_deep, ls>>>(a);
                else k_poa_edges<<<ln.blocks, 32 * DP_WARPS_PER_BLOCK, smem, ls>>>(a);
                HGPU_CUDA(ctx, cudaGetLastError());
                ctx->launches++; S->st.dp_launches++;
                if (side_by_side) {
                    HGPU_CUDA(ctx, cudaEventRecord(S->cev[li], ls));
                    HGPU_CUDA(ctx, cudaStreamWaitEvent(st, S->cev[li], 0));
                }
                if (S->verbose) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ls); vclass.push_back(e); }
                if (S->verbose) {
                    double cc = 0, cmax = 0; for (size_t q = c.a; q < c.b; ++q) { cc += est[q].cells; cmax = std::max(cmax, est[q].cells); }
                    fprintf(stderr, "[poa] attempt %d growth %.2f wave %zu/%zu class %zu/%zu%s: %u edges on %u warps, slot %.1f MB, ws %.1f MB, %.3e cells (largest edge %.3e)\n",
                            attempt, growth, wi, waves.size(), ln.ci, classes.size(), c.pool ? " (pool)" : c.deep ? " (deep)" : "", n_items, ln.warps, c.slot / 1048576.0, c.wl.bytes / 1048576.0, cc, cmax);
                }
            }
            if (S->verbose) {
                cudaEventRecord(ve, st); cudaEventSynchronize(ve);
                float ms = 0; cudaEventElapsedTime(&ms, vb, ve);
                fprintf(stderr, "[poa] wave %zu: %.1f ms; classes ended at", wi, ms);
                for (size_t q = 0; q < vclass.size(); ++q) {
                    float cm = 0; cudaEventElapsedTime(&cm, vb, vclass[q]);
                    fprintf(stderr, " %zu:%.0f", wave[q].ci, cm);
                    cudaEventDestroy(vclass[q]);
                }
                fprintf(stderr, " ms\n");
                cudaEventDestroy(vb); cudaEventDestroy(ve);
            }
        }
        if (team_launched || pool_launched) {
            HGPU_CUDA(ctx, cudaStreamWaitEvent(st, S->ev_join, 0));
            if (S->verbose) {
                const auto t0 = std::chrono::steady_clock::now();
                cudaStreamSynchronize(st);
                fprintf(stderr, "[poa] team / pool kernel outlasted the other classes by %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            }
        }
        if (S->timing) {
            HGPU_CUDA(ctx, cudaEventRecord(S->ev1, st));
            HGPU_CUDA(ctx, cudaEventSynchronize(S->ev1));
            float ms = 0; cudaEventElapsedTime(&ms, S->ev0, S->ev1);
            S->st.ms_dp += ms;
        }
        HGPU_CUDA(ctx, cudaMemcpyAsync(status_h.data(), S->status.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, st));
        HGPU_CUDA(ctx, cudaStreamSynchronize(st));
        S->passes.push_back(std::move(pass));
        // pool cursor of this pass is only meaningful for the pass; retry whatever did not fit
        std::vector<uint32_t> next;
        for (uint32_t e : pending) {
            uint32_t s_ = status_h[e];
            if (s_ == ST_CAPACITY || s_ == ST_TOO_LARGE || s_ == ST_POOL) next.push_back(e);
        }
        if (growth >= 1.0) break;     // worst case already tried
        pending.swap(next);
        growth = attempt >= 2 ? 1.0 : growth * 2.5;
    }
    HGPU_CUDA(ctx, cudaMemcpyAsync(len_h.data(), S->cons_len.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, st));
    unsigned long long sth[8];
    HGPU_CUDA(ctx, cudaMemcpyAsync(sth, S->stats.p, sizeof sth, cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
#if HGPU_PHASE_CLOCKS
    {
        unsigned long long pc[24];
        HGPU_CUDA(ctx, cudaMemcpy(pc, S->stats.p + 8, sizeof pc, cudaMemcpyDeviceToHost));
        static const char* nm_edges[9] = {"queue", "init", "fill", "traceback", "add_alignment", "toposort", "dp_records", "consensus", "publish"};
        static const char* nm_pool[9] = {"claim/idle", "open", "stripes", "traceback", "add_alignment", "records+toposort", "dp_records+plan", "advance", "-"};
        const char* const* nm = pool_ran ? nm_pool : nm_edges;
        double tot = 0; for (int i = 0; i < 9; ++i) tot += (double)pc[i];
        fprintf(stderr, "[phase clocks, warp-cycles of the last class]");
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.1f%%", nm[i], 100.0 * (double)pc[i] / (tot > 0 ? tot : 1));
        fprintf(stderr, "\n");
        if (pc[14]) fprintf(stderr, "[toposort] nodes %llu, batch passes %llu, DFS roots %llu (one per %.1f nodes), visits %llu (%.2f per root)\n",
                            pc[14], pc[15], pc[16], (double)pc[14] / (double)(pc[16] ? pc[16] : 1), pc[17], (double)pc[17] / (double)(pc[16] ? pc[16] : 1));
        if (pc[13]) fprintf(stderr, "[traceback] path steps %llu, walk iterations %llu (%.2f steps each), tiles loaded %llu (%.1f steps each), generic steps %llu\n",
                            pc[13], pc[11], (double)pc[13] / (double)(pc[11] ? pc[11] : 1), pc[10], (double)pc[13] / (double)(pc[10] ? pc[10] : 1), pc[12]);
    }
#endif
    if (S->verbose) fprintf(stderr, "[poa] host: %.1f ms for the whole run (kernels %.1f ms)\n", host_ms(), S->st.ms_dp);
    if (S->verbose >= 2 && S->edge_clk.p) {
        // the pool's time line: how many edges were open at each tenth of the kernel, and the edges that ended last
        std::vector<unsigned long long> clk((size_t)n_edges * 2);
        HGPU_CUDA(ctx, cudaMemcpy(clk.data(), S->edge_clk.p, clk.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull, t1 = 0;
        for (uint32_t e = 0; e < n_edges; ++e) if (clk[2 * e] && clk[2 * e + 1]) { t0 = std::min(t0, clk[2 * e]); t1 = std::max(t1, clk[2 * e + 1]); }
        if (t1 > t0) {
            const double span = (double)(t1 - t0);
            fprintf(stderr, "[poa] pool time line (%.1f ms): edges open at each 5 %%:", span / 1e6);
            for (int q = 0; q < 20; ++q) {
                const unsigned long long t = t0 + (unsigned long long)(span * (q + 0.5) / 20);
                uint32_t open = 0;
                for (uint32_t e = 0; e < n_edges; ++e) if (clk[2 * e] && clk[2 * e] <= t && clk[2 * e + 1] > t) ++open;
                fprintf(stderr, " %u", open);
            }
            fprintf(stderr, "\n");
            std::vector<uint32_t> ord;
            for (uint32_t e = 0; e < n_edges; ++e) if (clk[2 * e] && clk[2 * e + 1]) ord.push_back(e);
            std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return clk[2 * x + 1] > clk[2 * y + 1]; });
            for (size_t i = 0; i < ord.size() && i < 12; ++i) {
                const uint32_t e = ord[i];
                const uint32_t s0 = edge_seg_off[e], R = edge_seg_off[e + 1] - s0;
                uint64_t bases = 0; for (uint32_t k = 0; k < R; ++k) bases += seg_off[s0 + k + 1] - seg_off[s0 + k];
                fprintf(stderr, "[poa]   edge %u: %u reads, %.0f bp mean, started %.1f ms, ended %.1f ms (%.1f ms)\n", e, R, R ? (double)bases / R : 0.0,
                        (clk[2 * e] - t0) / 1e6, (clk[2 * e + 1] - t0) / 1e6, (clk[2 * e + 1] - clk[2 * e]) / 1e6);
            }
        }
    }
    if (sth[7]) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "k_poa_pool: a block ran out of tasks while edges were still open");
    S->st.cells = sth[0]; S->st.cells_padded = sth[1]; S->st.alignments = sth[2]; S->st.alignments_i32 = sth[3] & 0xFFFFFFFFull; S->st.alignments_rel16 = sth[3] >> 32; S->st.bases_in = sth[4];
    return HGPU_OK;
}

// offsets + gather into d_out (device) ; returns total bytes in *total
static int poa_gather(hgpu_t* ctx, uint32_t n_edges, const std::vector<uint32_t>& len_h, uint64_t* out_cons_off,
                      uint8_t* d_out, uint64_t out_cap, uint64_t* total) {
    PoaState* S = poa_state(ctx);
    cudaStream_t st = ctx->stream;
    uint64_t off = 0;
    for (uint32_t e = 0; e < n_edges; ++e) { out_cons_off[e] = off; off += len_h[e]; }
    out_cons_off[n_edges] = off;
    *total = off;
    S->st.bases_out = off;
    if (n_edges == 0 || off == 0) return HGPU_OK;
    HGPU_CUDA(ctx, S->d_off.ensure(n_edges + 1));
    HGPU_CUDA(ctx, cudaMemcpyAsync(S->d_off.p, out_cons_off, (size_t)(n_edges + 1) * 8, cudaMemcpyHostToDevice, st));
    uint32_t blocks = std::min<uint32_t>(n_edges, (uint32_t)ctx->sm_count * 16);
    k_poa_gather<<<blocks, 256, 0, st>>>(S->cons_pos.p, S->cons_len.p, S->d_off.p, n_edges, d_out, out_cap);
    HGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches++; S->st.other_launches++;
    return HGPU_OK;
}

static int poa_prepare(hgpu_t* ctx, const uint64_t* seg_off, const uint32_t* edge_seg_off, uint32_t n_edges, uint32_t band,
                       uint64_t* out_cons_off, uint32_t* out_status) {
    if (!ctx) return HGPU_E_INVALID;
    if (!edge_seg_off || !out_cons_off || !out_status || (!seg_off && n_edges && edge_seg_off[n_edges]))
        HGPU_FAIL(ctx, HGPU_E_INVALID, "null argument");
    if (band != 0) HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "band must be 0: the reference runs SPOA's full (unbanded) DP");
    for (uint32_t e = 0; e < n_edges; ++e)
        if (edge_seg_off[e + 1] < edge_seg_off[e]) HGPU_FAIL(ctx, HGPU_E_INVALID, "edge_seg_off not monotone at %u", e);
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    PoaState* S = poa_state(ctx);
    if (!S->ev0) {
        HGPU_CUDA(ctx, cudaEventCreate(&S->ev0)); HGPU_CUDA(ctx, cudaEventCreate(&S->ev1));
        HGPU_CUDA(ctx, cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming)); HGPU_CUDA(ctx, cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
        HGPU_CUDA(ctx, cudaStreamCreateWithFlags(&S->stream2, cudaStreamNonBlocking));
        for (int i = 0; i < PoaState::NCS; ++i) {
            HGPU_CUDA(ctx, cudaStreamCreateWithFlags(&S->cstream[i], cudaStreamNonBlocking));
            HGPU_CUDA(ctx, cudaEventCreateWithFlags(&S->cev[i], cudaEventDisableTiming));
        }
    }
    return HGPU_OK;
}

extern "C" int hgpu_poa_batch_dev(hgpu_t* ctx, const uint8_t* d_bases, const uint64_t* seg_off, const uint32_t* edge_seg_off,
                                  uint32_t n_edges, int8_t match, int8_t mismatch, int8_t gap, uint32_t band,
                                  uint8_t* d_out_cons, uint64_t out_cons_cap, uint64_t* out_cons_off, uint32_t* out_status) {
    int rc = poa_prepare(ctx, seg_off, edge_seg_off, n_edges, band, out_cons_off, out_status);
    if (rc) return rc;
    PoaState* S = poa_state(ctx);
    DpScores sc; make_scores(ctx, match, mismatch, gap, &sc);
    std::vector<uint32_t> status_h, len_h;
    PoaRunOpts ropt; ropt.force_i32 = S->cfg_force;
    rc = poa_run(ctx, d_bases, seg_off, edge_seg_off, n_edges, sc, ropt, status_h, len_h);
    if (rc) return rc;
    for (uint32_t e = 0; e < n_edges; ++e) out_status[e] = status_h[e];
    uint64_t total = 0;
    S->last_n_edges = n_edges;
    // offsets first: the caller learns the needed size even when its buffer is too small
    uint64_t off = 0;
    for (uint32_t e = 0; e < n_edges; ++e) off += len_h[e];
    S->last_total = off; S->have_result = true;
    if (off > out_cons_cap || (off && !d_out_cons)) {
        uint64_t o = 0;
        for (uint32_t e = 0; e < n_edges; ++e) { out_cons_off[e] = o; o += len_h[e]; }
        out_cons_off[n_edges] = o;
        HGPU_FAIL(ctx, HGPU_E_NOSPACE, "consensus needs %llu bytes, caller gave %llu", (unsigned long long)off, (unsigned long long)out_cons_cap);
    }
    rc = poa_gather(ctx, n_edges, len_h, out_cons_off, d_out_cons, out_cons_cap, &total);
    if (rc) return rc;
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HGPU_OK;
}

extern "C" int hgpu_poa_batch(hgpu_t* ctx, const uint8_t* bases, const uint64_t* seg_off, const uint32_t* edge_seg_off,
                              uint32_t n_edges, int8_t match, int8_t mismatch, int8_t gap, uint32_t band,
                              uint8_t* out_cons, uint64_t out_cons_cap, uint64_t* out_cons_off, uint32_t* out_status) {
    int rc = poa_prepare(ctx, seg_off, edge_seg_off, n_edges, band, out_cons_off, out_status);
    if (rc) return rc;
    PoaState* S = poa_state(ctx);
    const uint64_t n_bases = n_edges ? seg_off[edge_seg_off[n_edges]] : 0;
    if (n_bases && !bases) HGPU_FAIL(ctx, HGPU_E_INVALID, "null bases");
    HGPU_CUDA(ctx, S->d_bases.ensure(n_bases + 16));
    if (n_bases) HGPU_CUDA(ctx, cudaMemcpyAsync(S->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, ctx->stream));
    HGPU_CUDA(ctx, S->d_out.ensure(out_cons_cap + 16));
    rc = hgpu_poa_batch_dev(ctx, S->d_bases.p, seg_off, edge_seg_off, n_edges, match, mismatch, gap, band,
                            S->d_out.p, out_cons_cap, out_cons_off, out_status);
    if (rc) return rc;
    const uint64_t total = out_cons_off[n_edges];
    if (total) {
        if (!out_cons) HGPU_FAIL(ctx, HGPU_E_INVALID, "null out_cons");
        HGPU_CUDA(ctx, cudaMemcpyAsync(out_cons, S->d_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HGPU_OK;
}

extern "C" int hgpu_poa_fetch(hgpu_t* ctx, uint8_t* out_cons, uint64_t out_cons_cap) {
    if (!ctx) return HGPU_E_INVALID;
    PoaState* S = poa_state(ctx);
    if (!S->have_result) HGPU_FAIL(ctx, HGPU_E_INVALID, "no POA result to fetch");
    if (S->last_total > out_cons_cap) HGPU_FAIL(ctx, HGPU_E_NOSPACE, "consensus needs %llu bytes", (unsigned long long)S->last_total);
    if (S->last_total == 0) return HGPU_OK;
    if (!out_cons) HGPU_FAIL(ctx, HGPU_E_INVALID, "null out_cons");
    HGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n_edges = S->last_n_edges;
    std::vector<uint32_t> len_h(n_edges);
    HGPU_CUDA(ctx, cudaMemcpyAsync(len_h.data(), S->cons_len.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> off(n_edges + 1);
    HGPU_CUDA(ctx, S->d_out.ensure(S->last_total + 16));
    uint64_t total = 0;
    int rc = poa_gather(ctx, n_edges, len_h, off.data(), S->d_out.p, S->last_total, &total);
    if (rc) return rc;
    HGPU_CUDA(ctx, cudaMemcpyAsync(out_cons, S->d_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    HGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HGPU_OK;
}

extern "C" int hgpu_poa_get_stats(const hgpu_t* ctx, hgpu_poa_stats* out) {
    if (!ctx || !out) return HGPU_E_INVALID;
    if (!ctx->poa) { memset(out, 0, sizeof *out); return HGPU_OK; }
    *out = ctx->poa->st;
    return HGPU_OK;
}

extern "C" int hgpu_poa_set_timing(hgpu_t* ctx, int enabled) {
    if (!ctx) return HGPU_E_INVALID;
    poa_state(ctx)->timing = enabled != 0;
    return HGPU_OK;
}

extern "C" int hgpu_poa_configure(hgpu_t* ctx, uint64_t arena_bytes, uint32_t max_warps) {
    if (!ctx) return HGPU_E_INVALID;
    PoaState* S = poa_state(ctx);
    S->cfg_arena_bytes = arena_bytes;
    S->cfg_max_warps = max_warps;
    return HGPU_OK;
}

extern "C" int hgpu_poa_debug(hgpu_t* ctx, const uint8_t* bases, const uint64_t* seg_off, uint32_t n_segs, uint32_t n_prior,
                              int8_t match, int8_t mismatch, int8_t gap, int force_i32, int reserved,
                              int32_t* H, uint64_t H_cap, int32_t* aln_node, int32_t* aln_pos, uint32_t aln_cap,
                              uint32_t* rank2node, uint8_t* node_code, uint32_t* pred_off, uint32_t* pred_node, uint32_t* pred_weight,
                              uint32_t node_cap, uint32_t edge_cap, hgpu_poa_dbg_sizes* sizes) {
    (void)reserved;
    if (!ctx || !seg_off || !sizes) return HGPU_E_INVALID;
    if (n_prior == 0) HGPU_FAIL(ctx, HGPU_E_INVALID, "n_prior must be >= 1");
    uint32_t edge_seg_off[2] = {0, n_segs};
    uint64_t dummy_off[2]; uint32_t dummy_status[1];
    int rc = poa_prepare(ctx, seg_off, edge_seg_off, 1, 0, dummy_off, dummy_status);
    if (rc) return rc;
    PoaState* S = poa_state(ctx);
    cudaStream_t st = ctx->stream;
    const uint64_t n_bases = seg_off[n_segs];
    HGPU_CUDA(ctx, S->d_bases.ensure(n_bases + 16));
    if (n_bases) HGPU_CUDA(ctx, cudaMemcpyAsync(S->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, st));
    DpScores sc; make_scores(ctx, match, mismatch, gap, &sc);
    PoaRunOpts opt; opt.stop_round = n_prior; opt.force_i32 = force_i32; opt.max_warps = DP_WARPS_PER_BLOCK;
    std::vector<uint32_t> status_h, len_h;
    rc = poa_run(ctx, S->d_bases.p, seg_off, edge_seg_off, 1, sc, opt, status_h, len_h);
    if (rc) return rc;
    if (status_h[0] != ST_OK) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "debug edge ended with status %u", status_h[0]);
    std::vector<uint32_t> lens;
    for (uint32_t s = 0; s < n_segs; ++s) if (seg_off[s + 1] > seg_off[s]) lens.push_back((uint32_t)(seg_off[s + 1] - seg_off[s]));
    if (lens.empty()) { memset(sizes, 0, sizeof *sizes); return HGPU_OK; }
    // any of the block's warps may have pulled the edge: find the workspace that holds a graph
    const WsLayout wl = S->last_wl;
    const uint64_t slot = S->last_slot;
    std::vector<uint8_t> wsh;
    wsh.resize(wl.bytes * DP_WARPS_PER_BLOCK);
    HGPU_CUDA(ctx, cudaMemcpyAsync(wsh.data(), S->ws.p, wsh.size(), cudaMemcpyDeviceToHost, st));
    HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    int w = -1;
    for (int q = 0; q < DP_WARPS_PER_BLOCK; ++q) {
        const uint32_t* hdr = reinterpret_cast<const uint32_t*>(wsh.data() + (size_t)q * wl.bytes + wl.o_hdr);
        if (hdr[HDR_N_NODES] != 0 && hdr[HDR_N_NODES] <= wl.ncap) { w = q; break; }
    }
    if (w < 0) HGPU_FAIL(ctx, HGPU_E_INTERNAL, "debug: no workspace holds a graph");
    uint8_t* base = wsh.data() + (size_t)w * wl.bytes;
    GraphView g = bind_graph(base, wl);
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(base + wl.o_hdr);
    const uint32_t N = *g.n_nodes, NE = *g.n_edges;
    if (N > node_cap || NE > edge_cap) HGPU_FAIL(ctx, HGPU_E_NOSPACE, "debug: graph has %u nodes / %u edges", N, NE);
    sizes->n_nodes = N; sizes->n_edges = NE; sizes->aln_len = 0; sizes->L = 0;
    uint32_t pe = 0;
    for (uint32_t r = 0; r < N; ++r) {
        uint32_t v = g.rank2node[r];
        if (rank2node) rank2node[r] = v;
        if (pred_off) pred_off[r] = pe;
        for (uint32_t x = g.in_head[v]; x != NIL; x = g.e_next_in[x]) {
            if (pred_node) pred_node[pe] = g.e_begin[x];
            if (pred_weight) pred_weight[pe] = g.e_w[x];
            ++pe;
        }
    }
    if (pred_off) pred_off[N] = pe;
    if (node_code) for (uint32_t i = 0; i < N; ++i) node_code[i] = (uint8_t)"ACGT"[g.code[i]];
    if (lens.size() <= n_prior) return HGPU_OK;      // graph only
    const uint32_t V = hdr[HDR_LAST_V], L = hdr[HDR_LAST_L];
    const int mode = (int)hdr[HDR_LAST_P16];
    const int bias = (int)hdr[HDR_LAST_BIAS];
    sizes->L = L;
    const uint32_t n = *g.aln_len;
    sizes->aln_len = n;
    if (aln_node && aln_pos) {
        if (n > aln_cap) HGPU_FAIL(ctx, HGPU_E_NOSPACE, "debug: alignment has %u pairs", n);
        for (uint32_t t = 0; t < n; ++t) {   // stored last pair first
            int32_t rk = g.aln_rank[n - 1 - t];
            aln_node[t] = rk < 0 ? -1 : (int32_t)g.rank2node[rk];
            aln_pos[t] = g.aln_pos[n - 1 - t];
        }
    }
    if (H) {
        const uint64_t cells = (uint64_t)(V + 1) * (L + 1);
        if (cells > H_cap) HGPU_FAIL(ctx, HGPU_E_NOSPACE, "debug: H needs %llu cells", (unsigned long long)cells);
        DevBuf<int32_t> dH;
        HGPU_CUDA(ctx, dH.alloc(cells));
        uint8_t* slot_p = S->arena.p + (uint64_t)w * slot;
        if (mode == DPM_ABS16) k_poa_dump_H<DP_NW16, true><<<ctx->sm_count * 4, 256, 0, st>>>(slot_p, V, L, bias, sc.g, dH.p);
        else if (mode == DPM_REL16) k_poa_dump_H<DP_NW16, true, true><<<ctx->sm_count * 4, 256, 0, st>>>(slot_p, V, L, 0, sc.g, dH.p);
        else k_poa_dump_H<DP_NW32, false><<<ctx->sm_count * 4, 256, 0, st>>>(slot_p, V, L, bias, sc.g, dH.p);
        HGPU_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        HGPU_CUDA(ctx, cudaMemcpyAsync(H, dH.p, cells * 4, cudaMemcpyDeviceToHost, st));
        HGPU_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return HGPU_OK;
}
