/* haslr_b200 — C ABI of the B200-native haslr_assemble hot path.
 *
 * The reference (vpc-ccg/haslr) has no plugin/FFI layer; its hot path is reached through C++ stage functions
 * called from src/haslr_assemble/src/main.cpp. Each entry point below names the reference interface it
 * replaces (file:line relative to /root/reference/src/haslr_assemble/src). INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns HGPU_OK (0) or a
 * negative HGPU_E_* code and never calls exit(); hgpu_last_error() gives the detail string. One context per
 * GPU; a context is not thread-safe. Unless a function says "_dev", all pointers are HOST pointers and the
 * call copies host->device and device->host itself. All kernels of a call are ordered on the context's stream
 * (hgpu_set_stream; default: the legacy default stream).
 */
#ifndef HASLR_B200_H
#define HASLR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HGPU_OK             0
#define HGPU_E_INVALID     -1   /* bad argument */
#define HGPU_E_CUDA        -2   /* CUDA runtime error (see hgpu_last_error) */
#define HGPU_E_NOMEM       -3   /* device or host allocation failed */
#define HGPU_E_NOSPACE     -4   /* caller-provided output capacity too small */
#define HGPU_E_UNSUPPORTED -5   /* valid request this build cannot serve (e.g. band != 0) */
#define HGPU_E_INTERNAL    -6   /* a kernel reported an inconsistent state */

typedef struct hgpu_ctx hgpu_t;

int         hgpu_create(int device /* -1 = current */, hgpu_t** out);
void        hgpu_destroy(hgpu_t* ctx);
int         hgpu_set_stream(hgpu_t* ctx, void* cuda_stream /* cudaStream_t, NULL = default */);
const char* hgpu_strerror(int code);
const char* hgpu_last_error(const hgpu_t* ctx);
int         hgpu_abi_version(void);   /* 3 since the device-resident stages (hgpu_hits_group, *_dev) and hgpu_stage_stats were added; 4: hgpu_host_staging */
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t    hgpu_launch_count(const hgpu_t* ctx);
/* Page-locked host staging buffer `which` (0 or 1) of the context, at least `bytes` long: owned by the context, grown on demand,
 * kept until hgpu_destroy; growing it invalidates the pointer returned before. A host buffer handed to a non-_dev call is
 * copied by DMA at link speed when it lies in such a buffer (from pageable memory the driver stages the copy through a
 * bounce buffer at a fifth of that). The reference has no counterpart: its segments are std::string::substr copies made per
 * edge inside the thread that aligns them (Assemble.cpp:529-532); here the host gathers all segments of a batch once,
 * straight into this buffer (host/assemble.cpp call_consensus). */
int         hgpu_host_staging(hgpu_t* ctx, uint32_t which, uint64_t bytes, void** out);

/* Per-stage figures of the most recent call of each stage: CUDA-event time of the stage's kernels (copies excluded; only
 * when timing is on, hgpu_set_timing), kernel launches, the units the stage processed (what the roofline figures of
 * DESIGN.md are counted in) and the bytes this context has copied between host and device since it was created. */
typedef struct {
    float    ms_k0, ms_k1, ms_k2, ms_k4;
    uint32_t launches_k0, launches_k1, launches_k2, launches_k4;
    uint64_t k0_text_bytes, k0_rows, k0_ops;      /* PAF bytes scanned, rows and CIGAR runs emitted */
    uint64_t k1_hits, k1_reads, k1_elems;         /* hit rows read, reads, compact elements written */
    uint64_t k2_pairs, k2_entries;                /* adjacent pairs, directed edge entries */
    uint64_t k4_edges, k4_supports, k4_runs;      /* edges, supports, CIGAR runs walked */
    uint64_t h2d_bytes, d2h_bytes;
} hgpu_stage_stats;
int hgpu_get_stage_stats(const hgpu_t* ctx, hgpu_stage_stats* out);
/* CUDA-event timing of the kernels of every stage (one event pair per call; also switches hgpu_poa_set_timing). */
int hgpu_set_timing(hgpu_t* ctx, int enabled);

/* ---- (0) PAF text -> hit table -----------------------------------------------------------------------------
 * Replaces the text side of load_alignment (Longread.cpp:250-289): getline, str_split on tabs, str2type<uint32_t> on
 * columns 1-4 and 6-12 (Common.hpp:126-133), the strand character of column 5, the first optional column that starts
 * with "cg:Z:". Every row is kept (the load filters F1-F4 are part of hgpu_compact_lr); empty lines are skipped; a line
 * with fewer than 12 columns fails the call with HGPU_E_INVALID (the reference would index past its fields vector).
 * text: minimap2 PAF, n_bytes of it (whole lines; a final line without a line feed is accepted). The CIGAR payload
 * becomes run-length operations (len << 2) | op, op 0 = M, 1 = I, 2 = anything else — the layout hgpu_hits_t takes.
 * hgpu_paf_tokenize uploads and scans the text and reports the table's size; hgpu_paf_fetch then fills caller arrays of
 * *out_n_rows entries (cg_off: *out_n_rows + 1, cg_ops: *out_n_ops). */
int hgpu_paf_tokenize(hgpu_t* ctx, const char* text, uint64_t n_bytes, uint64_t* out_n_rows, uint64_t* out_n_ops);
int hgpu_paf_fetch(hgpu_t* ctx, uint32_t* q_id, uint32_t* q_len, uint32_t* q_start, uint32_t* q_end, uint8_t* is_rev,
                   uint32_t* t_id, uint32_t* t_len, uint32_t* t_start, uint32_t* t_end, uint32_t* n_match, uint32_t* n_block,
                   uint8_t* mapq, uint32_t* cg_off, uint32_t* cg_ops);

/* The table hgpu_paf_tokenize made STAYS on the device (so does the one hgpu_compact_lr uploads): the *_dev entry points of
 * (i), (ii) and (iv) read it there and nothing but results travels back. hgpu_hits_group slices it by read - rows must be
 * grouped by ascending read id (Longread.cpp:57-84 assumes it silently; here a violation is HGPU_E_INVALID, as is a read id
 * >= n_reads) - and optionally returns the n_reads + 1 offsets. */
int hgpu_hits_group(hgpu_t* ctx, uint32_t n_reads, uint32_t* out_read_off /* may be NULL */);

/* ---- (i) PAF hits -> compact long reads ------------------------------------------------------------------
 * Replaces, per long read: load_alignment's filters F1-F4 and per-read sort (Longread.cpp:234-302),
 * process_lr_alignment_group (Longread.cpp:182-232), fix_overlapping_alignments (Longread.cpp:430-512) and
 * build_compact_longreads / find_best_scheduling (Longread.cpp:514-624; Longread.hpp:89).
 * Hits of read r are rows read_off[r] .. read_off[r+1) of the column arrays, in PAF order; reads in ascending id.
 */
typedef struct {
    uint32_t n_hits;
    const uint32_t* q_start; const uint32_t* q_end;          /* PAF col 3,4 */
    const uint32_t* t_id; const uint32_t* t_len;             /* PAF col 6 (integer name), 7 */
    const uint32_t* t_start; const uint32_t* t_end;          /* PAF col 8,9 */
    const uint32_t* n_match; const uint32_t* n_block;        /* PAF col 10,11 */
    const uint8_t* is_rev; const uint8_t* mapq;              /* PAF col 5 == '-', col 12 */
    const uint32_t* cg_off;   /* n_hits+1 offsets into cg_ops */
    const uint32_t* cg_ops;   /* run-length cg:Z: ops: (len << 2) | op, op 0 = M, 1 = I, 2 = D/other */
} hgpu_hits_t;

typedef struct {
    double   min_aln_sim;     /* --aln-sim, default 0.85  (Commandline.cpp:46-66) */
    double   uniq_freq;       /* calc_uniq_freq, Contig.cpp:162-174 */
    double   max_uniq_dev;    /* --uniq-dev, default 0.15 */
    uint32_t min_aln_block;   /* --aln-block, default 500 */
    uint32_t min_aln_mapq;    /* 55 */
} hgpu_k1_params;

/* One element of a compact long read, coordinates after the overlap fix; the kept part of the hit's CIGAR is
 * ops [cg_lo .. cg_hi], the first with length cg_lo_len and the last with cg_hi_len. */
typedef struct {
    uint32_t hit;
    uint32_t q_start, q_end, t_start, t_end, n_match, n_block;
    uint32_t cg_lo, cg_lo_len, cg_hi, cg_hi_len;
} hgpu_cl_elem;

/* out_elems: capacity n_hits. out_read_off: n_reads+1. *out_n = number of elements written. */
int hgpu_compact_lr(hgpu_t* ctx, const hgpu_hits_t* hits, const uint32_t* read_off, uint32_t n_reads,
                    const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm,
                    hgpu_cl_elem* out_elems, uint32_t* out_read_off, uint64_t* out_n);

/* Same on the resident, grouped hit table; the elements also stay on the device for (ii) / (iv). out_elems, out_tid (contig
 * of each element) and out_rev (its strand) may each be NULL when the host has no use for them. */
int hgpu_compact_lr_dev(hgpu_t* ctx, uint32_t n_reads, const double* mean_kmer, uint32_t n_contigs, const hgpu_k1_params* prm,
                        hgpu_cl_elem* out_elems, uint32_t* out_tid, uint8_t* out_rev, uint32_t* out_read_off, uint64_t* out_n);

/* ---- (ii) compact long reads -> backbone edge table -----------------------------------------------------
 * Replaces bbg_build_graph + bbg_add_edge (Backbone_graph.cpp:148-171,10-25; Backbone_graph.hpp:60) and the
 * rule of bbg_remove_weak_edges (Backbone_graph.cpp:348-375; Backbone_graph.hpp:63).
 * Output: directed entries (each undirected edge appears as edge and twin) sorted by
 * key64 = ((node1<<1|rev1) << 32) | (node2<<1|rev2)  — the iteration order of the reference's
 * graph[node].edges[rev] std::maps; supports of an entry in ascending (read id, element index) order;
 * out_keep[e] = 1 iff the entry survives bbg_remove_weak_edges. Capacities: 2*n_pairs entries (+1 for
 * out_supp_off) and 2*n_pairs supports, n_pairs = sum over reads of max(len-1, 0). */
typedef struct { uint32_t lr_id_strand; /* lr_id | lr_strand << 31 */ uint32_t cmp_head; uint32_t cmp_tail; } hgpu_edge_supp;

int hgpu_backbone_edges(hgpu_t* ctx, const uint32_t* cl_tid, const uint8_t* cl_rev, const uint32_t* cl_read_off,
                        uint32_t n_reads, uint32_t min_edge_sup,
                        uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint8_t* out_keep,
                        uint64_t* out_n_entries);

/* Same on the compact reads hgpu_compact_lr[_dev] left on the device. entry_cap = capacity of the output arrays in entries
 * (out_supp_off: entry_cap + 1); HGPU_E_NOSPACE if 2 * pairs exceeds it. */
int hgpu_backbone_edges_dev(hgpu_t* ctx, uint32_t min_edge_sup, uint64_t entry_cap,
                            uint64_t* out_key, uint32_t* out_supp_off, hgpu_edge_supp* out_supp, uint8_t* out_keep,
                            uint64_t* out_n_entries);

/* ---- (iv) edge coordinates: the stretch of each supporting long read that spans an edge's gap -----------------
 * (numbered after (iii) because it was added after it; in the pipeline it runs between (ii) + graph cleaning and (iii).)
 * Replaces asm_calc_single_edge_coordinates with asm_best_supported_interval_contig1/2 and asm_find_lr_pos
 * (Assemble.cpp:24-155,157-363) and the pthread edge queue of asm_calc_edge_coordinates_MT around them (Assemble.cpp:436-477).
 * Edge e: rev1 = edge_rev[e] & 1, rev2 = (edge_rev[e] >> 1) & 1 (strands of node1 / node2 as asm_get_next_edge yields
 * them); its supports are supp[supp_off[e] .. supp_off[e+1]) in edge_supp order (the strand bit of lr_id_strand is
 * ignored). elems / cl_read_off are hgpu_compact_lr's outputs, read_len the long-read lengths, hit_is_rev / cg_off /
 * cg_ops the hit table the elements point into (hgpu_hits_t). Outputs are aligned with the inputs: one
 * hgpu_edge_coord per edge and one hgpu_supp_coord per support. A support yields a cns_supp entry
 * {lr_id, lr_strand, spos = lr_start + 1, epos = lr_end - 1} (Assemble.cpp:322-325) iff in_best and both lr_start and
 * lr_end are != -1; an edge with n_cns == 0 keeps no consensus support (Assemble.cpp:244-251,354-361). */
typedef struct {
    uint32_t int1_lo, int1_hi;   /* best supported interval on the head contig (best_int1) */
    uint32_t int2_lo, int2_hi;   /* ... on the tail contig (best_int2) */
    uint32_t c1, c2;             /* contig1_pos, contig2_pos */
    uint32_t n_best;             /* |best_lrs1 n best_lrs2| */
    uint32_t n_cns;              /* of which both CIGAR walks succeeded = cns_supp.size() */
} hgpu_edge_coord;
typedef struct {
    int64_t  lr_start, lr_end;   /* asm_find_lr_pos results (exclusive bounds), -1 = refused; meaningful iff in_best */
    uint32_t lr_strand;          /* strand of the read relative to the edge */
    uint32_t in_best;            /* 1 iff the support is in both best sets */
} hgpu_supp_coord;
int hgpu_edge_coords(hgpu_t* ctx, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                     const hgpu_cl_elem* elems, const uint32_t* cl_read_off, uint32_t n_reads, const uint32_t* read_len,
                     const uint8_t* hit_is_rev, const uint32_t* cg_off, const uint32_t* cg_ops, uint32_t n_hits,
                     hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp);

/* Same with the compact reads and the hit table resident on the device (only edges, supports and read lengths go up). */
int hgpu_edge_coords_dev(hgpu_t* ctx, uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                         const uint32_t* read_len, uint32_t n_reads, hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp);

/* ---- (iii) batched POA consensus -------------------------------------------------------------------------
 * Replaces the body of asm_calc_single_cns_seq (Assemble.cpp:499-554), i.e. per backbone edge the five SPOA
 * calls createAlignmentEngine(kNW, match, mismatch, gap), createGraph(), align_sequence_with_graph(),
 * add_alignment(), generate_consensus(), and the pthread edge queue around it (Assemble.cpp:365-434,562-605).
 * bases: ASCII, non-ACGT reads as 'A' (Compressed_sequence.cpp:57). Segment s = bases[seg_off[s] .. seg_off[s+1]);
 * edge e owns segments edge_seg_off[e] .. edge_seg_off[e+1) in the order they are fed to the graph; empty
 * segments are skipped (Assemble.cpp:537). band must be 0 (full DP, as SPOA).
 * out_cons (capacity out_cons_cap bytes) receives the concatenated consensus strings, out_cons_off n_edges+1
 * offsets, out_status n_edges per-edge codes (0 = ok). If out_cons_cap is too small the call returns
 * HGPU_E_NOSPACE, out_cons_off[n_edges] holds the needed size, and hgpu_poa_fetch can still collect the result.
 */
int hgpu_poa_batch(hgpu_t* ctx, const uint8_t* bases, const uint64_t* seg_off, const uint32_t* edge_seg_off,
                   uint32_t n_edges, int8_t match, int8_t mismatch, int8_t gap, uint32_t band,
                   uint8_t* out_cons, uint64_t out_cons_cap, uint64_t* out_cons_off, uint32_t* out_status);

/* Same, with `bases` and `out_cons` already DEVICE pointers (seg_off/edge_seg_off/out_cons_off/out_status stay
 * host arrays: they are metadata the host scheduler needs). */
int hgpu_poa_batch_dev(hgpu_t* ctx, const uint8_t* d_bases, const uint64_t* seg_off, const uint32_t* edge_seg_off,
                       uint32_t n_edges, int8_t match, int8_t mismatch, int8_t gap, uint32_t band,
                       uint8_t* d_out_cons, uint64_t out_cons_cap, uint64_t* out_cons_off, uint32_t* out_status);

int hgpu_poa_fetch(hgpu_t* ctx, uint8_t* out_cons, uint64_t out_cons_cap);

typedef struct {
    uint64_t cells;            /* sum over alignments of (|V|+1)*(L+1): the algorithmic DP cells */
    uint64_t cells_padded;     /* cells actually computed (stripe padding included) */
    uint64_t alignments;       /* graph-NW alignments run */
    uint64_t alignments_i32;   /* of which in the int32 kernel (scores the packed int16 arithmetic cannot hold) */
    uint64_t bases_in;         /* segment bases consumed */
    uint64_t bases_out;        /* consensus bases produced */
    uint64_t dp_launches, update_launches, other_launches;
    float    ms_dp, ms_update, ms_other;   /* CUDA-event time per kernel family, only if timing enabled */
    uint64_t arena_bytes;      /* score-matrix arena size used */
    uint64_t alignments_rel16; /* alignments whose range exceeds plain int16 and ran in row-relative int16 cells */
} hgpu_poa_stats;
int hgpu_poa_get_stats(const hgpu_t* ctx, hgpu_poa_stats* out);
/* CUDA-event timing of the POA kernels: one event pair around the launches of a scheduling pass, read after the
 * synchronisation the pass ends with anyway (ms_dp). */
int hgpu_poa_set_timing(hgpu_t* ctx, int enabled);
/* Tuning knobs: arena_bytes = score-matrix arena (0 = auto), max_warps = resident warps of the warp-per-edge kernels (0 = auto). */
int hgpu_poa_configure(hgpu_t* ctx, uint64_t arena_bytes, uint32_t max_warps);

/* Debug/inspection (used by the parity tests): graph after the first n_prior non-empty segments of ONE edge, in
 * rank order, plus the score matrix and alignment of the next segment. H is written in the reference's H space
 * (row-major (V+1)*(L+1) int32). force_i32: 0 = the encoding the batch path would pick, 1 = int32 cells, 2 = int16 cells relative
 * to the row base (REL16); `reserved` is ignored. Any output may be NULL. */
typedef struct { uint32_t n_nodes, n_edges, aln_len, L; } hgpu_poa_dbg_sizes;
int hgpu_poa_debug(hgpu_t* ctx, const uint8_t* bases, const uint64_t* seg_off, uint32_t n_segs, uint32_t n_prior,
                   int8_t match, int8_t mismatch, int8_t gap, int force_i32, int reserved,
                   int32_t* H, uint64_t H_cap, int32_t* aln_node, int32_t* aln_pos, uint32_t aln_cap,
                   uint32_t* rank2node, uint8_t* node_code, uint32_t* pred_off, uint32_t* pred_node, uint32_t* pred_weight,
                   uint32_t node_cap, uint32_t edge_cap, hgpu_poa_dbg_sizes* sizes);

#ifdef __cplusplus
}
#endif
#endif /* HASLR_B200_H */
