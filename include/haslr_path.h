/* haslr_b200 — the whole backbone + POA path behind one call (libhaslr_path.so: C++ host code above the C ABI of
 * haslr_b200.h). Replaces the stage sequence of the reference's main() between "alignments loaded" and "consensus called":
 * src/haslr_assemble/src/main.cpp:116-207 (fix_overlapping_alignments, build_compact_longreads, bbg_build_graph,
 * bbg_remove_weak_edges, the Cleaning.cpp stages, asm_calc_edge_coordinates_MT, asm_cal_cns_seq_MT) plus the text side of
 * load_alignment (Longread.cpp:234-302). bin/haslr_assemble runs the same code and writes the reference's files;
 * bench.py times this call (inputs in host memory -> consensus in host memory). */
#ifndef HASLR_PATH_H
#define HASLR_PATH_H
#include <stdint.h>

#include "haslr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct haslr_path haslr_path_t;

/* Load contigs (FASTA with KC:i: / km:f: tags), long reads (FASTA/FASTQ, gzip ok) and the PAF text into host memory. */
int  haslr_path_open(const char* contigs_fa, const char* reads_fa, const char* paf, haslr_path_t** out);
void haslr_path_close(haslr_path_t* p);
int  haslr_path_sizes(const haslr_path_t* p, uint64_t* n_contigs, uint64_t* n_reads, uint64_t* read_bases, uint64_t* paf_bytes);

typedef struct {
    uint64_t n_rows;        /* PAF rows tokenised */
    uint64_t n_edges;       /* backbone edges that went through POA (after cleaning) */
    uint64_t poa_bases;     /* long-read bases fed to POA: the numerator of the Mbases/s metric */
    uint64_t cons_bytes;    /* consensus bases produced */
    uint32_t cons_crc;      /* CRC-32 of the consensus strings concatenated in edge order */
    uint32_t pad;
    double   s_tokenize, s_k1, s_k2, s_clean, s_coords, s_poa, s_total;   /* wall seconds per stage (copies and host work included) */
} haslr_path_result;

/* One pass of the path with the reference's default options (--aln-block 500 --aln-sim 0.85 --edge-sup 3). ctxs: one context per GPU
 * (the POA edges are dealt to them by estimated cost); threads: host threads for the segment gather; out_dir: NULL = keep
 * everything in memory, else the reference's output files are written there (the directory must exist). */
int  haslr_path_run(haslr_path_t* p, hgpu_t* const* ctxs, uint32_t n_ctx, uint32_t threads, const char* out_dir, haslr_path_result* res);

#ifdef __cplusplus
}
#endif
#endif /* HASLR_PATH_H */
