// The backbone + POA path of haslr_assemble from inputs in host memory to consensus strings in host memory:
// reference src/haslr_assemble/src/main.cpp:116-207 (load_alignment's text side, fix_overlapping_alignments,
// build_compact_longreads, bbg_build_graph, the cleaning stages, asm_calc_edge_coordinates_MT, asm_cal_cns_seq_MT).
// The hit table never comes back to the host: hgpu_paf_tokenize leaves it on the device, hgpu_compact_lr_dev /
// hgpu_backbone_edges_dev / hgpu_edge_coords_dev read it there; what travels is the text up, and compact reads, the edge
// table, coordinates and consensus down. Used by bin/haslr_assemble (files written) and by libhaslr_path.so (bench, tests).
#include <sys/time.h>
#include <cstring>
#include <zlib.h>

#include "haslr.hpp"

namespace haslr {

static double wall() { struct timeval t; gettimeofday(&t, nullptr); return t.tv_sec + t.tv_usec / 1e6; }

int run_path(const PathInputs& in, Options& opt, const std::vector<hgpu_t*>& ctxs, const std::string& d, bool logs, PathResult& r) {
    const bool files = !d.empty();
    hgpu_t* ctx = ctxs[0];
    const ContigStore& contigs = in.contigs;
    const SeqStore& reads = in.reads;
    const uint32_t n_reads = (uint32_t)reads.size();
    const double t_start = wall();
    double t0 = t_start;
    auto lap = [&](double& slot) { const double t = wall(); slot += t - t0; t0 = t; };

    // (0) PAF text -> hit table, on the device and staying there
    uint64_t rows = 0, ops = 0;
    if (hgpu_paf_tokenize(ctx, in.paf_data(), in.paf_size(), &rows, &ops) != HGPU_OK ||
        hgpu_hits_group(ctx, n_reads, nullptr) != HGPU_OK) {
        fprintf(stderr, "[ERROR] PAF: %s\n", hgpu_last_error(ctx));
        return HGPU_E_INVALID;
    }
    r.n_rows = rows;
    lap(r.t.tokenize);

    // (i) filters + per-read sort + overlap fix + chaining
    CompactReads& cl = r.cl;
    {
        hgpu_k1_params p{opt.min_aln_sim, opt.uniq_freq, opt.max_uniq_dev, opt.min_aln_block, opt.min_aln_mapq};
        // the element records themselves are only needed for compact_uniq.txt and the coordinate log: without an output directory
        // they are neither fetched nor allocated (zero-filling 40 bytes per PAF row cost more than the kernels of this stage)
        const bool want_elems = files;
        if (want_elems) { cl.elems.resize(rows + 1); cl.tid.resize(rows + 1); cl.rev.resize(rows + 1); }
        cl.off.resize((size_t)n_reads + 1);
        uint64_t n = 0;
        int rc = hgpu_compact_lr_dev(ctx, n_reads, contigs.mean_kmer.data(), (uint32_t)contigs.size(), &p, want_elems ? cl.elems.data() : nullptr,
                                     want_elems ? cl.tid.data() : nullptr, want_elems ? cl.rev.data() : nullptr, cl.off.data(), &n);
        if (rc != HGPU_OK) { fprintf(stderr, "[ERROR] hgpu_compact_lr_dev: %s\n", hgpu_last_error(ctx)); return rc; }
        cl.elems.resize(want_elems ? n : 0); cl.tid.resize(want_elems ? n : 0); cl.rev.resize(want_elems ? n : 0);
    }
    if (files) write_compact(cl, d + "/compact_uniq.txt");
    lap(r.t.k1);

    // (ii) edge table on the GPU, graph container on the host
    Graph& g = r.g;
    {
        uint64_t pairs = 0;
        for (size_t q = 0; q + 1 < cl.off.size(); ++q) if (cl.off[q + 1] - cl.off[q] > 1) pairs += cl.off[q + 1] - cl.off[q] - 1;
        const size_t cap = 2 * pairs + 1;
        std::vector<uint64_t> key(cap); std::vector<uint32_t> soff(cap + 1); std::vector<hgpu_edge_supp> supp(cap); std::vector<uint8_t> keep(cap);
        uint64_t n = 0;
        int rc = hgpu_backbone_edges_dev(ctx, opt.min_edge_sup, cap, key.data(), soff.data(), supp.data(), keep.data(), &n);
        if (rc != HGPU_OK) { fprintf(stderr, "[ERROR] hgpu_backbone_edges_dev: %s\n", hgpu_last_error(ctx)); return rc; }
        key.resize(n); soff.resize(n + 1); keep.resize(n);
        graph_from_edge_table(g, contigs.size(), key, soff, supp, nullptr);
    }
    if (files) { write_stats(g, contigs, d + "/backbone.01.init.stat"); write_gfa(g, contigs, d + "/backbone.01.init.gfa"); }
    lap(r.t.k2);

    // graph cleaning (host; Cleaning.cpp). Log files only when an output directory was given.
    const std::string nolog;
    auto lg = [&](const char* name) { return files ? d + name : nolog; };
    const int weak = remove_weak_edges(g, opt.min_edge_sup);
    if (files) { write_stats(g, contigs, d + "/backbone.02.weakEdge.stat"); write_gfa(g, contigs, d + "/backbone.02.weakEdge.gfa"); }
    int tips = clean_tips(g, 1, lg("/backbone.03.tip.log"));
    tips += clean_tips(g, 2, lg("/backbone.03.tip.log"));
    tips += clean_tips(g, 3, lg("/backbone.03.tip.log"));
    if (files) { write_stats(g, contigs, d + "/backbone.03.tip.stat"); write_gfa(g, contigs, d + "/backbone.03.tip.gfa"); }
    const int sb = clean_simple_bubbles(g, 4, lg("/backbone.04.simplebubble.log"));
    if (files) { write_stats(g, contigs, d + "/backbone.04.simplebubble.stat"); write_gfa(g, contigs, d + "/backbone.04.simplebubble.gfa"); }
    const int pb = clean_super_bubbles(g, lg("/backbone.05.superbubble.log"));
    if (files) { write_stats(g, contigs, d + "/backbone.05.superbubble.stat"); write_gfa(g, contigs, d + "/backbone.05.superbubble.gfa"); }
    const int mb = clean_small_bubbles(g, lg("/backbone.06.smallbubble.log"));
    if (files) { write_stats(g, contigs, d + "/backbone.06.smallbubble.stat"); write_gfa(g, contigs, d + "/backbone.06.smallbubble.gfa"); }
    if (files) report_branching(g, d + "/backbone.branching.log");
    fprintf(stderr, "       cleaning: removed %d weak edges, %d tips, %d simple / %d super / %d small bubbles\n", weak, tips, sb, pb, mb);
    lap(r.t.clean);

    // (iv) edge coordinates
    std::vector<EdgeRef>& edges = r.edges;
    enumerate_edges(g, 11, edges);
    if (calc_edge_coordinates(g, edges, contigs, reads, cl, ctx, (files && logs) ? d + "/log_coordinate.txt" : std::string()) != 0) return HGPU_E_INTERNAL;
    lap(r.t.coords);

    // (iii) all edges in one batched POA call per GPU
    enumerate_edges(g, 12, edges);
    r.poa_bases = 0;
    if (!edges.empty() && call_consensus(g, edges, reads, ctxs, d + "/log_consensus.txt", files && logs, opt.num_threads, &r.poa_bases) != 0) return HGPU_E_INTERNAL;
    lap(r.t.poa);
    r.t.total = wall() - t_start;
    uLong crc = crc32(0L, Z_NULL, 0);
    r.cons_bytes = 0;
    for (const EdgeRef& er : edges) {
        const std::string& c = g[er.node1].edges[er.rev1][(er.node2 << 1) | er.rev2].cns_seq;
        crc = crc32(crc, (const Bytef*)c.data(), (uInt)c.size());
        r.cons_bytes += c.size();
    }
    r.cons_crc = (uint32_t)crc;
    return HGPU_OK;
}

}  // namespace haslr

// ---------------------------------------------------------------------------------------------------------------------
// C entry points of libhaslr_path.so (bench.py / tests drive the whole path through these; declared in include/haslr_path.h)
// ---------------------------------------------------------------------------------------------------------------------
#include "../../include/haslr_path.h"

struct haslr_path {
    haslr::PathInputs in;
    haslr::Options opt;
    hgpu_t* staged_on = nullptr;        // the context whose staging buffer holds the PAF text
};

extern "C" int haslr_path_open(const char* contigs_fa, const char* reads_fa, const char* paf, haslr_path_t** out) {
    if (!contigs_fa || !reads_fa || !paf || !out) return HGPU_E_INVALID;
    haslr_path* p = new haslr_path();
    haslr::load_fasta(contigs_fa, p->in.contigs, &p->in.contigs);
    haslr::load_fasta(reads_fa, p->in.reads, nullptr);
    haslr::read_text_file(paf, p->in.paf_text);
    p->opt.uniq_freq = haslr::calc_uniq_freq(p->in.contigs);
    *out = p;
    return HGPU_OK;
}

extern "C" void haslr_path_close(haslr_path_t* p) { delete p; }

extern "C" int haslr_path_sizes(const haslr_path_t* p, uint64_t* n_contigs, uint64_t* n_reads, uint64_t* read_bases, uint64_t* paf_bytes) {
    if (!p) return HGPU_E_INVALID;
    if (n_contigs) *n_contigs = p->in.contigs.size();
    if (n_reads) *n_reads = p->in.reads.size();
    if (read_bases) *read_bases = p->in.reads.seq.size();
    if (paf_bytes) *paf_bytes = p->in.paf_size();
    return HGPU_OK;
}

extern "C" int haslr_path_run(haslr_path_t* p, hgpu_t* const* ctxs, uint32_t n_ctx, uint32_t threads, const char* out_dir, haslr_path_result* res) {
    if (!p || !ctxs || n_ctx == 0 || !res) return HGPU_E_INVALID;
    std::vector<hgpu_t*> cv(ctxs, ctxs + n_ctx);
    // first run on this context: the PAF text moves into the context's page-locked staging buffer, from where every pass uploads it by
    // DMA (171 MB of config 2: 3 ms against 15 ms from pageable memory); the pageable copy stays (another context may follow)
    if (p->staged_on != cv[0] && !p->in.paf_text.empty()) {
        void* pinned = nullptr;
        if (p->in.paf_text.size() <= (8ull << 30) && hgpu_host_staging(cv[0], 1, p->in.paf_text.size(), &pinned) == HGPU_OK) {
            memcpy(pinned, p->in.paf_text.data(), p->in.paf_text.size());
            p->in.paf_ptr = (const char*)pinned; p->in.paf_len = p->in.paf_text.size();
            p->staged_on = cv[0];
        } else { p->in.paf_ptr = nullptr; p->in.paf_len = 0; p->staged_on = nullptr; }
    }
    haslr::Options opt = p->opt;
    opt.num_threads = threads ? threads : 1;
    opt.gpus = (int)n_ctx;
    haslr::PathResult r;
    const int rc = haslr::run_path(p->in, opt, cv, out_dir ? out_dir : "", out_dir != nullptr, r);
    if (rc != HGPU_OK) return rc;
    res->n_rows = r.n_rows; res->n_edges = r.edges.size(); res->poa_bases = r.poa_bases; res->cons_bytes = r.cons_bytes; res->cons_crc = r.cons_crc;
    res->s_tokenize = r.t.tokenize; res->s_k1 = r.t.k1; res->s_k2 = r.t.k2; res->s_clean = r.t.clean; res->s_coords = r.t.coords; res->s_poa = r.t.poa;
    res->s_total = r.t.total;
    return HGPU_OK;
}
