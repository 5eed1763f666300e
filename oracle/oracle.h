/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C entry points of the CPU restatement of the haslr_assemble hot path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product library (haslr_b200/csrc) never includes or links anything in oracle/.
 *
 * K1/K2 (oracle_k12.cpp) and K4 (oracle_coords.cpp) restate Longread.cpp / Backbone_graph.cpp / the coordinate part
 * of Assemble.cpp and are pinned against outputs of the reference binary built by oracle/Makefile (tests/golden/). K3 (oracle_poa.cpp) wraps the restated SPOA
 * 1.1.3 (spoa_restated/spoa.hpp): PARITY UNPINNED — see that header.
 */
#ifndef HASLR_ORACLE_H
#define HASLR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- K3: per-edge POA consensus (Assemble.cpp:479-560 + SPOA) ---------------------------------------- */
/* Same data contract as hgpu_poa_batch: bases are ASCII, segment s is bases[seg_off[s] .. seg_off[s+1]),
 * edge e owns segments edge_seg_off[e] .. edge_seg_off[e+1). Empty segments are skipped (Assemble.cpp:537).
 * simd != 0 selects the SSE4.1 int16 row fill when it fits (identical results). threads = pthread count, one
 * edge per grab (Assemble.cpp:386-434). out_cells: total DP cells (|V|+1)(L+1) summed over alignments.
 * out_nodes[e] = final node count of edge e's graph (may be NULL). Returns 0, or -1 if out_cap is too small. */
int oracle_poa_batch(const uint8_t* bases, const uint64_t* seg_off, const uint32_t* edge_seg_off, uint32_t n_edges,
                     int match, int mismatch, int gap, int simd, int threads,
                     uint8_t* out_cons, uint64_t out_cap, uint64_t* out_cons_off,
                     uint64_t* out_cells, uint32_t* out_nodes);

/* Debug: build the graph from the first n_prior non-empty segments of one edge, then align segment n_prior
 * against it. Outputs (any may be NULL): H as int32 row-major (n_nodes+1)*(L+1) in SPOA's H space;
 * alignment pairs; graph in rank order. Returns n_nodes, or -1 on capacity error. */
typedef struct {
    uint32_t n_nodes, n_edges, aln_len, L;
} oracle_poa_dbg_sizes;
int oracle_poa_debug(const uint8_t* bases, const uint64_t* seg_off, uint32_t n_segs, uint32_t n_prior,
                     int match, int mismatch, int gap,
                     int32_t* H, uint64_t H_cap,
                     int32_t* aln_node, int32_t* aln_pos, uint32_t aln_cap,
                     uint32_t* rank2node, uint8_t* node_code, uint32_t* pred_off, uint32_t* pred_node, uint32_t* pred_weight,
                     uint32_t node_cap, uint32_t edge_cap, oracle_poa_dbg_sizes* sizes);

/* ---- K1: PAF hits -> compact long reads (Longread.cpp:182-302,374-624) -------------------------------- */
typedef struct {
    uint32_t n_hits;
    const uint32_t* q_start; const uint32_t* q_end;
    const uint32_t* t_id; const uint32_t* t_len; const uint32_t* t_start; const uint32_t* t_end;
    const uint32_t* n_match; const uint32_t* n_block;
    const uint8_t* is_rev; const uint8_t* mapq;
    const uint32_t* cg_off;   /* n_hits+1 offsets into cg_ops */
    const uint32_t* cg_ops;   /* run-length CIGAR: (len << 2) | op, op 0 = M, 1 = I, 2 = D (anything else) */
} oracle_hits_t;

typedef struct {
    double min_aln_sim;       /* 0.85 */
    double uniq_freq;         /* Contig.cpp:162-174 */
    double max_uniq_dev;      /* 0.15 */
    uint32_t min_aln_block;   /* 500 */
    uint32_t min_aln_mapq;    /* 55 */
} oracle_k1_params;

/* One element of a compact long read, post overlap-fix. cg window = kept part of the hit's CIGAR:
 * ops [cg_lo .. cg_hi] of the hit, the first with length cg_lo_len, the last with length cg_hi_len. */
typedef struct {
    uint32_t hit;             /* index of the source PAF hit */
    uint32_t q_start, q_end, t_start, t_end, n_match, n_block;
    uint32_t cg_lo, cg_lo_len, cg_hi, cg_hi_len;
} oracle_cl_elem;

/* hits of read r are read_off[r] .. read_off[r+1) in PAF order. out_elems capacity n_hits.
 * out_read_off has n_reads+1 entries. Returns total elements. */
int64_t oracle_compact_lr(const oracle_hits_t* hits, const uint32_t* read_off, uint32_t n_reads,
                          const double* mean_kmer, const oracle_k1_params* prm,
                          oracle_cl_elem* out_elems, uint32_t* out_read_off);

/* ---- K2: compact long reads -> backbone edge table (Backbone_graph.cpp:10-25,148-171,348-375) -------- */
typedef struct { uint32_t lr_id_strand; /* lr_id | strand << 31 */ uint32_t cmp_head; uint32_t cmp_tail; } oracle_edge_supp;
/* Directed edge entries (every undirected edge appears as edge and twin) sorted by key64 =
 * ((node1<<1|rev1) << 32) | (node2<<1|rev2), i.e. the reference's std::map iteration order.
 * Capacity: 2 * n_pairs entries and 2 * n_pairs supports. keep[e] = support >= min_edge_sup.
 * Returns the number of directed entries. */
int64_t oracle_backbone_edges(const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off, uint32_t n_reads,
                              uint32_t min_edge_sup,
                              uint64_t* out_key, uint32_t* out_supp_off, oracle_edge_supp* out_supp, uint8_t* out_keep);

/* ---- K0: PAF text -> hit table (Longread.cpp:250-289, text side only; all rows kept) -------------------------- */
/* Returns the number of rows, -1 if a capacity is too small, -2 for a line with fewer than 12 columns. Empty lines are
 * skipped. cg_off has rows + 1 entries; *n_ops_out = cg_off[rows]. */
int64_t oracle_parse_paf(const char* text, uint64_t n_bytes, uint64_t row_cap, uint64_t op_cap,
                         uint32_t* q_id, uint32_t* q_len, uint32_t* q_start, uint32_t* q_end, uint8_t* is_rev,
                         uint32_t* t_id, uint32_t* t_len, uint32_t* t_start, uint32_t* t_end, uint32_t* n_match,
                         uint32_t* n_block, uint8_t* mapq, uint32_t* cg_off, uint32_t* cg_ops, uint64_t* n_ops_out);

/* ---- K4: edge coordinates (Assemble.cpp:24-155,157-363) --------------------------------------------------- */
/* edge e: rev1 = edge_rev[e] & 1, rev2 = edge_rev[e] >> 1 & 1; its supports are supp[supp_off[e] .. supp_off[e+1]) in
 * edge_supp order (lr_id_strand's strand bit is ignored); elems / cl_read_off = the compact long reads;
 * hit_is_rev / cg_off / cg_ops = the PAF hit table the elements point into. Outputs are aligned with the inputs:
 * one oracle_edge_coord per edge, one oracle_supp_coord per support (in_best = 1 for members of best_lrs1 n best_lrs2;
 * lr_start / lr_end as asm_find_lr_pos returns them, -1 = refused; a support yields a cns_supp entry iff in_best and
 * both are != -1: {lr_id, lr_strand, lr_start + 1, lr_end - 1}). */
typedef struct { uint32_t int1_lo, int1_hi, int2_lo, int2_hi, c1, c2, n_best, n_cns; } oracle_edge_coord;
typedef struct { int64_t lr_start, lr_end; uint32_t lr_strand, in_best; } oracle_supp_coord;
int oracle_edge_coords(uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const oracle_edge_supp* supp,
                       const oracle_cl_elem* elems, const uint32_t* cl_read_off, const uint32_t* read_len,
                       const uint8_t* hit_is_rev, const uint32_t* cg_off, const uint32_t* cg_ops,
                       oracle_edge_coord* out_edge, oracle_supp_coord* out_supp);

#ifdef __cplusplus
}
#endif
#endif
