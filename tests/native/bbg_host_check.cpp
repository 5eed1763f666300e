// TEST-ONLY host harness for the binary's backbone-graph code (haslr_b200/host/bbg.cpp): builds the graph container from
// an edge table (the layout hgpu_backbone_edges returns), then runs the cleaning stages and writers in main.cpp's
// order into out_dir, under the reference's file names. Never shipped.
#include <string>
#include <vector>
#include "../../haslr_b200/host/haslr.hpp"

using namespace haslr;

extern "C" int bbghost_run(uint32_t n_contigs, const uint32_t* contig_len, uint64_t n_entries, const uint64_t* key, const uint32_t* supp_off,
                           const hgpu_edge_supp* supp, uint32_t min_edge_sup, const char* out_dir) {
    ContigStore contigs;
    contigs.off.push_back(0);
    for (uint32_t i = 0; i < n_contigs; ++i) { contigs.seq.append(contig_len[i], 'A'); contigs.off.push_back(contigs.seq.size()); }
    contigs.kmer_count.assign(n_contigs, 0); contigs.mean_kmer.assign(n_contigs, 0.0);
    std::vector<uint64_t> k(key, key + n_entries);
    std::vector<uint32_t> so(supp_off, supp_off + n_entries + 1);
    std::vector<hgpu_edge_supp> sp(supp, supp + so[n_entries]);
    Graph g;
    graph_from_edge_table(g, n_contigs, k, so, sp, nullptr);
    const std::string d(out_dir);
    write_stats(g, contigs, d + "/backbone.01.init.stat"); write_gfa(g, contigs, d + "/backbone.01.init.gfa");
    remove_weak_edges(g, min_edge_sup);
    write_stats(g, contigs, d + "/backbone.02.weakEdge.stat"); write_gfa(g, contigs, d + "/backbone.02.weakEdge.gfa");
    clean_tips(g, 1, d + "/backbone.03.tip.log"); clean_tips(g, 2, d + "/backbone.03.tip.log"); clean_tips(g, 3, d + "/backbone.03.tip.log");
    write_stats(g, contigs, d + "/backbone.03.tip.stat"); write_gfa(g, contigs, d + "/backbone.03.tip.gfa");
    clean_simple_bubbles(g, 4, d + "/backbone.04.simplebubble.log");
    write_stats(g, contigs, d + "/backbone.04.simplebubble.stat"); write_gfa(g, contigs, d + "/backbone.04.simplebubble.gfa");
    clean_super_bubbles(g, d + "/backbone.05.superbubble.log");
    write_stats(g, contigs, d + "/backbone.05.superbubble.stat"); write_gfa(g, contigs, d + "/backbone.05.superbubble.gfa");
    clean_small_bubbles(g, d + "/backbone.06.smallbubble.log");
    write_stats(g, contigs, d + "/backbone.06.smallbubble.stat"); write_gfa(g, contigs, d + "/backbone.06.smallbubble.gfa");
    report_branching(g, d + "/backbone.branching.log");
    return 0;
}
