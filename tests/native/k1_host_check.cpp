// TEST-ONLY host harness for haslr_b200/csrc/k1_core.cuh (the per-read core is __host__ __device__):
// runs the product's load filters + k1_process_read read by read on the CPU, and exposes the libstdc++-ordered
// sort so it can be compared with std::sort on tie-heavy input. Never shipped.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../haslr_b200/csrc/k1_core.cuh"

using namespace hgpu;

extern "C" int64_t k1host_compact_lr(const uint32_t* q_start, const uint32_t* q_end, const uint32_t* t_id, const uint32_t* t_len,
                                     const uint32_t* t_start, const uint32_t* t_end, const uint32_t* n_match, const uint32_t* n_block,
                                     const uint8_t* is_rev, const uint8_t* mapq, const uint32_t* cg_off, const uint32_t* cg_ops,
                                     const uint32_t* read_off, uint32_t n_reads, const double* mean_kmer,
                                     double min_aln_sim, double uniq_freq, double max_uniq_dev, uint32_t min_aln_block, uint32_t min_aln_mapq,
                                     ClElem* out, uint32_t* out_read_off) {
    HitCols h{q_start, q_end, t_id, t_len, t_start, t_end, n_match, n_block, is_rev, mapq, cg_off, cg_ops};
    K1Params p{min_aln_sim, uniq_freq, max_uniq_dev, min_aln_block, min_aln_mapq};
    int64_t n_out = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        out_read_off[r] = (uint32_t)n_out;
        const uint32_t b = read_off[r], e = read_off[r + 1], cap = e - b + 1;
        std::vector<uint32_t> idx(cap), dp(cap), cand(cap);
        std::vector<int32_t> prevc(cap);
        std::vector<uint8_t> take(cap);
        std::vector<K1Hit> hit(cap);
        uint32_t cnt = 0;
        for (uint32_t i = b; i < e; ++i) if (k1_load_filter(h, i, mean_kmer, p)) idx[cnt++] = i;
        n_out += k1_process_read(h, mean_kmer, p, idx.data(), cnt, hit.data(), dp.data(), prevc.data(), cand.data(), take.data(), out + n_out);
    }
    out_read_off[n_reads] = (uint32_t)n_out;
    return n_out;
}

// sorts idx[0..n) by (q_end, q_start) with the product's restated algorithm and with std::sort; returns #mismatches
extern "C" int k1host_sort_check(const uint32_t* q_end, const uint32_t* q_start, uint32_t n, uint32_t* out_idx) {
    std::vector<uint32_t> a(n), b(n);
    for (uint32_t i = 0; i < n; ++i) a[i] = b[i] = i;
    KeyLess lt{q_end, q_start};
    libstdcxx_sort(a.data(), (int)n, lt);
    std::sort(b.begin(), b.end(), lt);
    int bad = 0;
    for (uint32_t i = 0; i < n; ++i) { bad += a[i] != b[i]; if (out_idx) out_idx[i] = a[i]; }
    return bad;
}
