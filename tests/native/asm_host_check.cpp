// TEST-ONLY host harness for the binary's stitching code (haslr_b200/host/assemble.cpp: edge enumeration, simple-path
// extraction, assemble_path / write_assembly) on top of the graph code of bbg.cpp. Two steps, because the consensus of
// every edge comes from outside (the GPU in the binary, the oracle / golden fixture in the test):
//   asmhost_prepare : edge table -> graph -> cleaning -> the edges in asm_get_next_edge order, with their supports
//   asmhost_finish  : per-edge anchor positions, consensus strings and support counts -> asm.final.fa / .ann
// Never shipped.
#include <string>
#include <vector>
#include "../../haslr_b200/host/haslr.hpp"

using namespace haslr;

static Graph G;
static ContigStore CONTIGS;
static std::vector<EdgeRef> EDGES;

extern "C" int asmhost_prepare(uint32_t n_contigs, const char* seq, const uint64_t* off, uint64_t n_entries, const uint64_t* key,
                               const uint32_t* supp_off, const hgpu_edge_supp* supp, uint32_t min_edge_sup, const char* scratch_dir) {
    CONTIGS = ContigStore();
    CONTIGS.seq.assign(seq, seq + off[n_contigs]);
    CONTIGS.off.assign(off, off + n_contigs + 1);
    CONTIGS.kmer_count.assign(n_contigs, 0); CONTIGS.mean_kmer.assign(n_contigs, 0.0);
    std::vector<uint64_t> k(key, key + n_entries);
    std::vector<uint32_t> so(supp_off, supp_off + n_entries + 1);
    std::vector<hgpu_edge_supp> sp(supp, supp + so[n_entries]);
    G.clear();
    graph_from_edge_table(G, n_contigs, k, so, sp, nullptr);
    const std::string d(scratch_dir);
    remove_weak_edges(G, min_edge_sup);
    clean_tips(G, 1, d + "/tip.log"); clean_tips(G, 2, d + "/tip.log"); clean_tips(G, 3, d + "/tip.log");
    clean_simple_bubbles(G, 4, d + "/simple.log");
    clean_super_bubbles(G, d + "/super.log");
    clean_small_bubbles(G, d + "/small.log");
    enumerate_edges(G, 11, EDGES);
    return (int)EDGES.size();
}

// edges as (node1, rev1, node2, rev2) and their edge_supp lists flattened; returns the number of supports
extern "C" int asmhost_edges(uint32_t* out4, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint32_t supp_cap) {
    uint32_t n = 0;
    out_supp_off[0] = 0;
    for (size_t e = 0; e < EDGES.size(); ++e) {
        const EdgeRef& er = EDGES[e];
        out4[4 * e] = er.node1; out4[4 * e + 1] = er.rev1; out4[4 * e + 2] = er.node2; out4[4 * e + 3] = er.rev2;
        for (const EdgeSupp& s : G[er.node1].edges[er.rev1][(er.node2 << 1) | er.rev2].edge_supp) {
            if (n >= supp_cap) return -1;
            out_supp[n++] = hgpu_edge_supp{s.lr_id, s.cmp_head_id, s.cmp_tail_id};
        }
        out_supp_off[e + 1] = n;
    }
    return (int)n;
}

extern "C" int asmhost_finish(const uint32_t* head_end, const uint32_t* tail_beg, const uint32_t* n_cns, const char* cons, const uint64_t* cons_off,
                              const char* out_dir) {
    enumerate_edges(G, 12, EDGES);           // as main.cpp does before the consensus stage (same order)
    for (size_t e = 0; e < EDGES.size(); ++e) {
        const EdgeRef& er = EDGES[e];
        Edge& e1 = G[er.node1].edges[er.rev1][(er.node2 << 1) | er.rev2];
        Edge& e2 = G[er.node2].edges[1 - er.rev2][(er.node1 << 1) | (1 - er.rev1)];
        e1.head_end = e2.tail_beg = head_end[e];
        e1.tail_beg = e2.head_end = tail_beg[e];
        e1.cns_supp.assign(n_cns[e], CnsSupp{0, 0, 0, 0}); e2.cns_supp.assign(n_cns[e], CnsSupp{0, 0, 0, 0});   // stitching looks at the count only
        e1.cns_seq.assign(cons + cons_off[e], cons + cons_off[e + 1]);
        e2.cns_seq = revcomp(e1.cns_seq);
    }
    write_assembly(G, CONTIGS, out_dir);
    return 0;
}

// the segments call_consensus would feed to the POA for cns_supp entries (lr_id, lr_strand, spos, epos), concatenated
extern "C" long long asmhost_segments(const char* read_seq, const uint64_t* read_off, uint32_t n_reads, const uint32_t* cns4, uint32_t n_cns,
                                      char* out, uint64_t out_cap, uint64_t* out_off) {
    SeqStore reads;
    reads.seq.assign(read_seq, read_seq + read_off[n_reads]);
    reads.off.assign(read_off, read_off + n_reads + 1);
    uint64_t at = 0;
    out_off[0] = 0;
    for (uint32_t i = 0; i < n_cns; ++i) {
        const CnsSupp s{cns4[4 * i], cns4[4 * i + 1], cns4[4 * i + 2], cns4[4 * i + 3]};
        const uint32_t cnt = segment_length(reads, s);
        if (at + cnt > out_cap) return -1;
        write_segment(reads, s, cnt, out + at);
        at += cnt;
        out_off[i + 1] = at;
    }
    return (long long)at;
}
