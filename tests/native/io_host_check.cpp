// TEST-ONLY host harness for the binary's FASTA/FASTQ loader (haslr_b200/host/io.cpp): loads a file and hands back the
// folded sequences, record offsets and contig header tags. Never shipped. (Links libhaslr_b200.so because io.cpp's PAF
// loader calls the GPU tokeniser; nothing here does.)
#include <cstring>
#include "../../haslr_b200/host/haslr.hpp"

extern "C" long long iohost_load_fasta(const char* path, int contig_meta, char* seq_out, unsigned long long seq_cap,
                                       unsigned long long* off_out, unsigned long long off_cap, unsigned* kc_out, double* km_out) {
    haslr::ContigStore c;
    haslr::load_fasta(path, c, contig_meta ? &c : nullptr);
    if (c.seq.size() > seq_cap || c.off.size() > off_cap) return -1;
    memcpy(seq_out, c.seq.data(), c.seq.size());
    for (size_t i = 0; i < c.off.size(); ++i) off_out[i] = c.off[i];
    if (contig_meta) for (size_t i = 0; i < c.size(); ++i) { kc_out[i] = c.kmer_count[i]; km_out[i] = c.mean_kmer[i]; }
    return (long long)c.size();
}
extern "C" double iohost_uniq_freq(const char* path) {
    haslr::ContigStore c;
    haslr::load_fasta(path, c, &c);
    return haslr::calc_uniq_freq(c);
}
