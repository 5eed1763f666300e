// haslr_assemble (B200 edition) — host side above the C ABI. Shared declarations.
//
// Keeps the command line and the output files of the reference's haslr_assemble (reference
// src/haslr_assemble/src/main.cpp:28-228, Commandline.cpp:68-242) so the unmodified haslr.py driver can call it
// (bin/haslr.py:54-77). The three hot stages run on the GPU through include/haslr_b200.h; graph cleaning, edge
// coordinates and stitching are host code restated from the reference's behaviour (citations at each function).
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/haslr_b200.h"

namespace haslr {

struct Options {
    std::string contig_path, long_path, mapping_path, out_dir;
    bool long_fofn = false, mapping_fofn = false;
    uint32_t min_aln_block = 500;      // Commandline.cpp:46-66
    double min_aln_sim = 0.85;
    uint32_t min_aln_mapq = 55;
    double max_uniq_dev = 0.15;
    uint32_t min_edge_sup = 3;
    uint32_t num_threads = 1;
    double uniq_freq = 0;              // Contig.cpp:162-174
    int gpus = 1;                      // extension: --gpus N shards the POA edges over N devices
    std::string prog_version = "0.8a1-b200";
};

// Sequences are kept one byte per base, already folded the way the reference's 2-bit codec folds them
// (Compressed_sequence.cpp:10-19,57: A/C/G/T any case, everything else reads back as 'A').
struct SeqStore {
    std::vector<uint64_t> off;         // n+1
    std::string seq;
    size_t size() const { return off.empty() ? 0 : off.size() - 1; }
    uint32_t len(size_t i) const { return (uint32_t)(off[i + 1] - off[i]); }
    const char* data(size_t i) const { return seq.data() + off[i]; }
};

struct ContigStore : SeqStore {
    std::vector<uint32_t> kmer_count;  // KC:i:
    std::vector<double> mean_kmer;     // km:f:
};

// Compact long reads as the host keeps them: the elements, and per element the contig and strand of the hit behind it
// (all the host ever needs of the PAF hit table, which itself stays on the device: hgpu_paf_tokenize -> hgpu_compact_lr_dev).
struct CompactReads {
    std::vector<hgpu_cl_elem> elems;
    std::vector<uint32_t> tid;         // per element: t_id of its hit
    std::vector<uint8_t> rev;          // per element: is_rev of its hit
    std::vector<uint32_t> off;         // n_reads+1
};

struct EdgeSupp { uint32_t lr_id, lr_strand, cmp_head_id, cmp_tail_id; };          // Backbone_graph.hpp:23-29
struct CnsSupp { uint32_t lr_id, lr_strand, spos, epos; };                           // Backbone_graph.hpp:31-37
struct Edge {                                                                        // Backbone_graph.hpp:39-48
    uint32_t head_end = 0, tail_beg = 0;
    uint32_t flag = 0;
    std::string cns_seq;
    std::vector<EdgeSupp> edge_supp;
    std::vector<CnsSupp> cns_supp;
};
struct Node { std::map<uint32_t, Edge> edges[2]; };                                  // Backbone_graph.hpp:50-54
typedef std::vector<Node> Graph;

// io.cpp
void load_fasta(const std::string& path, SeqStore& out, ContigStore* contig_meta);
void load_fofn(const std::string& path, std::vector<std::string>& files);
void read_text_file(const std::string& path, std::vector<char>& text);      // appended; a missing final line feed is added
double calc_uniq_freq(const ContigStore& c);
std::string revcomp(const std::string& s);
FILE* open_write(const std::string& path);
FILE* open_append(const std::string& path);

// bbg.cpp
void graph_from_edge_table(Graph& g, size_t n_contigs, const std::vector<uint64_t>& key, const std::vector<uint32_t>& supp_off,
                           const std::vector<hgpu_edge_supp>& supp, const std::vector<uint8_t>* keep);
int remove_weak_edges(Graph& g, uint32_t min_edge_sup);
void write_stats(const Graph& g, const ContigStore& contigs, const std::string& path);
void write_gfa(const Graph& g, const ContigStore& contigs, const std::string& path);
void write_compact(const CompactReads& cl, const std::string& path);
int clean_tips(Graph& g, int max_depth, const std::string& logpath);
int clean_simple_bubbles(Graph& g, int max_depth, const std::string& logpath);
int clean_super_bubbles(Graph& g, const std::string& logpath);
int clean_small_bubbles(Graph& g, const std::string& logpath);
void report_branching(const Graph& g, const std::string& logpath);

// assemble.cpp
struct EdgeRef { uint32_t node1, rev1, node2, rev2; };
void enumerate_edges(Graph& g, uint32_t flag, std::vector<EdgeRef>& out);
int calc_edge_coordinates(Graph& g, const std::vector<EdgeRef>& edges, const ContigStore& contigs, const SeqStore& reads,
                          const CompactReads& cl, hgpu_t* ctx, const std::string& logpath);
uint32_t segment_length(const SeqStore& reads, const CnsSupp& s);                     // Assemble.cpp:529-532 (uint32 arithmetic, quirk Q7)
void write_segment(const SeqStore& reads, const CnsSupp& s, uint32_t cnt, char* out);   // the read, or its reverse complement, from spos
int call_consensus(Graph& g, const std::vector<EdgeRef>& edges, const SeqStore& reads, const std::vector<hgpu_t*>& ctxs,
                   const std::string& logpath, bool write_log, unsigned threads, uint64_t* bases_in);
void write_assembly(Graph& g, const ContigStore& contigs, const std::string& out_dir);

// pipeline.cpp: the path "PAF text + sequences in host memory -> consensus of every backbone edge in host memory" (SURVEY 8d),
// i.e. main.cpp:116-207 of the reference: K0 tokenise, K1 compact reads, K2 edge table, host graph cleaning, K4 coordinates,
// segment gather, K3 batched POA. out_dir empty: no files are written (bench / library use).
struct PathInputs {
    ContigStore contigs;
    SeqStore reads;
    std::vector<char> paf_text;
    // where the text is read from: paf_text, unless the caller moved it into page-locked memory (haslr_path_run does, once)
    const char* paf_ptr = nullptr; size_t paf_len = 0;
    const char* paf_data() const { return paf_ptr ? paf_ptr : paf_text.data(); }
    size_t paf_size() const { return paf_ptr ? paf_len : paf_text.size(); }
};
struct PathTimes { double tokenize = 0, k1 = 0, k2 = 0, clean = 0, coords = 0, poa = 0, total = 0; };      // wall seconds
struct PathResult {
    Graph g;
    CompactReads cl;
    std::vector<EdgeRef> edges;        // the POA'd edges in asm_get_next_edge order
    uint64_t n_rows = 0, poa_bases = 0, cons_bytes = 0;
    uint32_t cons_crc = 0;             // CRC-32 of the consensus strings concatenated in edge order
    PathTimes t;
};
int run_path(const PathInputs& in, Options& opt, const std::vector<hgpu_t*>& ctxs, const std::string& out_dir, bool logs, PathResult& r);

}  // namespace haslr
