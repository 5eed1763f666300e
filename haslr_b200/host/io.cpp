// Input side: FASTA/FASTQ (plain or gzip), PAF, plus the small shared helpers.
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "haslr.hpp"

namespace haslr {

// an empty path means "no file wanted" (library use: run_path without an output directory): the bytes go to /dev/null
FILE* open_write(const std::string& path) {
    FILE* fp = fopen(path.empty() ? "/dev/null" : path.c_str(), "w");
    if (!fp) { fprintf(stderr, "[ERROR] could not open file for writing: %s\n", path.c_str()); exit(EXIT_FAILURE); }
    return fp;
}
FILE* open_append(const std::string& path) {
    FILE* fp = fopen(path.empty() ? "/dev/null" : path.c_str(), "a");
    if (!fp) { fprintf(stderr, "[ERROR] could not open file for appending: %s\n", path.c_str()); exit(EXIT_FAILURE); }
    return fp;
}

// what a base reads back as after the reference's 2-bit round trip (Compressed_sequence.cpp:10-19,46-62)
static inline char fold_base(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 'A';
    }
}

namespace { struct FoldTable { char t[256]; FoldTable() { for (int i = 0; i < 256; ++i) t[i] = fold_base((unsigned char)i); } } FOLD; }

std::string revcomp(const std::string& s) {      // Common.cpp:186-193 (inputs here are pure ACGT)
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return r;
}

namespace {
struct GzLines {
    gzFile fp; std::vector<char> buf; size_t beg = 0, end = 0; bool eof = false;
    explicit GzLines(const std::string& path) : buf(1 << 20) {
        fp = gzopen(path.c_str(), "r");
        if (!fp) { fprintf(stderr, "[ERROR] could not open file: %s\n", path.c_str()); exit(EXIT_FAILURE); }
    }
    ~GzLines() { gzclose(fp); }
    // next line without its terminator; false at end of file
    bool next(std::string& line) {
        line.clear();
        while (true) {
            if (beg == end) {
                if (eof) return !line.empty();
                int n = gzread(fp, buf.data(), (unsigned)buf.size());
                if (n <= 0) { eof = true; return !line.empty(); }
                beg = 0; end = (size_t)n;
            }
            char* nl = (char*)memchr(buf.data() + beg, '\n', end - beg);
            if (nl) {
                line.append(buf.data() + beg, nl - (buf.data() + beg));
                beg = (size_t)(nl - buf.data()) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(buf.data() + beg, end - beg);
            beg = end;
        }
    }
};
}  // namespace

// FASTA or FASTQ, multi-line, optionally gzipped (what kseq.h accepts). Record i gets id i (file order), as in
// load_contig_compressed / load_longread_compressed (Contig.cpp:43-107, Longread.cpp:109-162).
void load_fasta(const std::string& path, SeqStore& out, ContigStore* meta) {
    GzLines in(path);
    std::string line;
    if (out.off.empty()) out.off.push_back(0);
    bool have = in.next(line);
    while (have) {
        if (line.empty()) { have = in.next(line); continue; }
        if (line[0] != '>' && line[0] != '@') { fprintf(stderr, "[ERROR] %s: not a FASTA/FASTQ header: %.40s\n", path.c_str(), line.c_str()); exit(EXIT_FAILURE); }
        const bool fastq = line[0] == '@';
        if (meta) {
            // KC:i: and km:f: from the header comment (Contig.cpp:63-66; the reference dereferences a NULL strstr when absent)
            size_t sp = line.find_first_of(" \t");
            const char* comment = sp == std::string::npos ? "" : line.c_str() + sp + 1;
            const char* p1 = strstr(comment, "KC:i:");
            const char* p2 = strstr(comment, "km:f:");
            if (!p1 || !p2) { fprintf(stderr, "[ERROR] %s: contig header without KC:i:/km:f: tags: %.60s\n", path.c_str(), line.c_str()); exit(EXIT_FAILURE); }
            meta->kmer_count.push_back((uint32_t)strtoul(p1 + 5, NULL, 10));
            meta->mean_kmer.push_back(strtod(p2 + 5, NULL));
        }
        size_t start = out.seq.size();
        have = in.next(line);
        while (have && !(line.size() && (line[0] == '>' || (fastq && line[0] == '+') || (!fastq && line[0] == '@')))) {
            {   // bulk append, then fold in place (sequence lines carry no blanks in practice; strip them if they do)
                const size_t at = out.seq.size();
                out.seq.append(line);
                char* p = &out.seq[at];
                size_t w = 0;
                for (size_t k = 0; k < line.size(); ++k) { const unsigned char c = (unsigned char)p[k]; if (c != ' ' && c != '\t') p[w++] = FOLD.t[c]; }
                out.seq.resize(at + w);
            }
            have = in.next(line);
        }
        if (fastq && have && line[0] == '+') {     // skip the quality block: as many characters as the sequence
            size_t need = out.seq.size() - start, got = 0;
            while (got < need && (have = in.next(line))) got += line.size();
            have = in.next(line);
        }
        out.off.push_back(out.seq.size());
    }
}

void load_fofn(const std::string& path, std::vector<std::string>& files) {
    std::ifstream fin(path.c_str());
    if (!fin.is_open()) { fprintf(stderr, "[ERROR] could not open file: %s\n", path.c_str()); exit(EXIT_FAILURE); }
    std::string line;
    while (getline(fin, line)) if (!line.empty()) files.push_back(line);
}

// A text file read whole (the PAF: it is tokenised on the GPU, hgpu_paf_tokenize). Further files of a fofn are appended.
void read_text_file(const std::string& path, std::vector<char>& text) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { fprintf(stderr, "[ERROR] (read_text_file) could not open file: %s\n", path.c_str()); exit(EXIT_FAILURE); }
    struct stat sb;
    const size_t guess = (fstat(fileno(fp), &sb) == 0 && sb.st_size > 0) ? (size_t)sb.st_size : (size_t)(1 << 24);
    size_t n = text.size();
    text.resize(n + guess + 1);
    while (true) {
        if (n == text.size()) text.resize(text.size() + text.size() / 2 + 1);
        const size_t got = fread(text.data() + n, 1, text.size() - n, fp);
        if (got == 0) break;
        n += got;
    }
    fclose(fp);
    text.resize(n);
    if (n && text[n - 1] != '\n') text.push_back('\n');
}

// mean km of the 20 longest contigs, pairs ordered descending by (len, km) — Contig.cpp:162-174
double calc_uniq_freq(const ContigStore& c) {
    std::vector<std::pair<uint32_t, double>> v(c.size());
    for (size_t i = 0; i < c.size(); ++i) v[i] = {c.len(i), c.mean_kmer[i]};
    std::sort(v.begin(), v.end(), std::greater<std::pair<uint32_t, double>>());
    double freq = 0;
    size_t i = 0;
    for (; i < 20 && i < v.size(); ++i) freq += v[i].second;
    return freq / (double)i;
}

}  // namespace haslr
