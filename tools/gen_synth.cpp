// Synthetic input generator for the haslr_assemble hot path.
// Emits the three files haslr_assemble reads: contigs.fa (SRCs with KC:i:/km:f: headers),
// reads.fa (integer-named long reads in id order) and map.paf (minimap2-style PAF with cg:Z:,
// rows grouped by read in ascending read id). Everything is derived from one seed with an
// in-file PRNG so that the output is identical on every box.
//
// usage: gen_synth OUTDIR GENOME_BP N_READS [MEAN_READ_LEN=8000] [SEED=1]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return next() % n; }
    double expo(double mean) { return -mean * std::log(1.0 - uni()); }
    double normal(double mu, double sd) {
        double u1 = 1.0 - uni(), u2 = uni();
        return mu + sd * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct Src { uint32_t beg, end; double km; };

static char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; } }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s OUTDIR GENOME_BP N_READS [MEAN_READ_LEN=8000] [SEED=1]\n", argv[0]); return 2; }
    std::string out = argv[1];
    const uint64_t G = strtoull(argv[2], nullptr, 10);
    const uint32_t n_reads = (uint32_t)strtoul(argv[3], nullptr, 10);
    const double mean_len = argc > 4 ? atof(argv[4]) : 8000.0;
    const uint64_t seed = argc > 5 ? strtoull(argv[5], nullptr, 10) : 1;
    const double p_ins = 0.04, p_del = 0.03, p_sub = 0.02;

    Rng rg(seed);
    std::string genome(G, 'A');
    for (uint64_t i = 0; i < G; ++i) genome[i] = "ACGT"[rg.next() >> 62];

    // SRCs tiling the genome
    std::vector<Src> srcs;
    {
        Rng r(seed * 7919 + 1);
        uint64_t pos = (uint64_t)r.expo(50.0);
        while (true) {
            uint64_t len = 250 + (uint64_t)r.expo(200.0);
            if (pos + len > G) break;
            double km = r.normal(30.0, 1.5);
            if (r.uni() < 0.03) km *= (double)(2 + r.below(4));
            srcs.push_back({(uint32_t)pos, (uint32_t)(pos + len), km});
            if (r.uni() < 0.10) pos = pos + len - (1 + r.below(29));
            else pos = pos + len + (uint64_t)r.expo(50.0);
        }
    }
    {
        FILE* f = fopen((out + "/contigs.fa").c_str(), "w");
        if (!f) { perror("contigs.fa"); return 1; }
        for (size_t i = 0; i < srcs.size(); ++i) {
            uint32_t len = srcs[i].end - srcs[i].beg;
            uint32_t kc = (uint32_t)std::lround(srcs[i].km * (len > 48 ? len - 48 : 1));
            fprintf(f, ">%zu LN:i:%u KC:i:%u km:f:%.1f\n", i, len, kc, srcs[i].km);
            fwrite(genome.data() + srcs[i].beg, 1, len, f);
            fputc('\n', f);
        }
        fclose(f);
    }

    FILE* fr = fopen((out + "/reads.fa").c_str(), "w");
    FILE* fp = fopen((out + "/map.paf").c_str(), "w");
    if (!fr || !fp) { perror("open"); return 1; }
    std::vector<char> fwd;        // erroneous read, genome-forward orientation
    std::vector<uint32_t> g2r;    // for each genome offset in the sampled interval: read position it maps to
    std::vector<uint8_t> gop;     // per genome offset: 0 = match, 1 = substitution, 2 = deleted
    std::vector<uint32_t> ins_after;  // inserted read bases emitted after genome offset i (before i+1)
    std::string rd, cg;
    uint64_t n_paf = 0;
    for (uint32_t rid = 0; rid < n_reads; ++rid) {
        Rng r(seed * 1000003ull + 17 + rid);
        uint64_t glen = (uint64_t)std::max(1000.0, r.normal(mean_len, mean_len * 0.15));
        if (glen > G) glen = G;
        uint64_t a = r.below(G - glen + 1);
        bool rev = r.uni() < 0.5;
        fwd.clear(); g2r.assign(glen, 0); gop.assign(glen, 0); ins_after.assign(glen, 0);
        for (uint64_t i = 0; i < glen; ++i) {
            double u = r.uni();
            g2r[i] = (uint32_t)fwd.size();
            if (u < p_del) { gop[i] = 2; }
            else if (u < p_del + p_sub) { gop[i] = 1; char c = genome[a + i]; char d; do { d = "ACGT"[r.next() >> 62]; } while (d == c); fwd.push_back(d); }
            else { gop[i] = 0; fwd.push_back(genome[a + i]); }
            while (r.uni() < p_ins) { fwd.push_back("ACGT"[r.next() >> 62]); ins_after[i]++; }
        }
        const uint32_t rlen = (uint32_t)fwd.size();
        rd.assign(fwd.begin(), fwd.end());
        if (rev) { std::reverse(rd.begin(), rd.end()); for (auto& c : rd) c = comp(c); }
        fprintf(fr, ">%u\n", rid);
        fwrite(rd.data(), 1, rd.size(), fr);
        fputc('\n', fr);

        // hits: every SRC overlapping [a, a+glen) by >= 100 bp
        size_t lo = std::lower_bound(srcs.begin(), srcs.end(), (uint32_t)a, [](const Src& s, uint32_t v) { return s.end <= v; }) - srcs.begin();
        for (size_t si = lo; si < srcs.size() && srcs[si].beg < a + glen; ++si) {
            uint64_t ts = std::max<uint64_t>(a, srcs[si].beg), te = std::min<uint64_t>(a + glen, srcs[si].end);
            if (te < ts + 100) continue;
            // minimap2 rarely reaches contig ends exactly: random end clipping, occasionally heavy
            if (r.uni() < 0.5) ts += r.below(12);
            if (r.uni() < 0.5) te -= r.below(12);
            if (r.uni() < 0.03) ts += r.below((te - ts) / 3 + 1);
            uint64_t gi0 = ts - a, gi1 = te - a;  // genome offsets [gi0, gi1)
            while (gi0 < gi1 && gop[gi0] == 2) ++gi0;           // CIGAR must start with M
            while (gi1 > gi0 && gop[gi1 - 1] == 2) --gi1;       // ... and end with M
            if (gi1 < gi0 + 50) continue;
            uint32_t nm = 0, nb = 0;
            cg.clear();
            char last = 0; uint32_t run = 0;
            auto push = [&](char op, uint32_t n) {
                if (n == 0) return;
                if (op == last) { run += n; } else { if (run) cg += std::to_string(run) + last; last = op; run = n; }
                nb += n;
            };
            for (uint64_t i = gi0; i < gi1; ++i) {
                if (gop[i] == 2) push('D', 1); else { push('M', 1); nm += gop[i] == 0; }
                if (i + 1 < gi1) push('I', ins_after[i]);
            }
            if (run) cg += std::to_string(run) + last;
            uint32_t qs = g2r[gi0], qe = g2r[gi1 - 1] + 1;  // forward-orientation read interval
            uint32_t q_start = rev ? rlen - qe : qs, q_end = rev ? rlen - qs : qe;
            uint32_t mapq = r.uni() < 0.02 ? (uint32_t)r.below(55) : 60;
            uint32_t tlen = srcs[si].end - srcs[si].beg;
            uint32_t t_start = (uint32_t)(a + gi0 - srcs[si].beg), t_end = (uint32_t)(a + gi1 - srcs[si].beg);
            fprintf(fp, "%u\t%u\t%u\t%u\t%c\t%zu\t%u\t%u\t%u\t%u\t%u\t%u\tNM:i:%u\tms:i:0\tcg:Z:%s\n", rid, rlen, q_start, q_end,
                    rev ? '-' : '+', si, tlen, t_start, t_end, nm, nb, mapq, nb - nm, cg.c_str());
            ++n_paf;
            // occasionally a second, palindromic-looking hit of the same SRC further along the read
            // (exercises the truncation rule of Longread.cpp:187-202)
            if (r.uni() < 0.004 && q_end + 600 < rlen) {
                uint32_t sh = 300 + (uint32_t)r.below(200);
                fprintf(fp, "%u\t%u\t%u\t%u\t%c\t%zu\t%u\t%u\t%u\t%u\t%u\t%u\tNM:i:%u\tms:i:0\tcg:Z:%s\n", rid, rlen, q_start + sh, q_end + sh,
                        rev ? '+' : '-', si, tlen, t_start, t_end, nm, nb, mapq, nb - nm, cg.c_str());
                ++n_paf;
            }
        }
    }
    fclose(fr); fclose(fp);
    fprintf(stderr, "gen_synth: %zu SRCs, %u reads, %llu PAF lines\n", srcs.size(), n_reads, (unsigned long long)n_paf);
    return 0;
}
