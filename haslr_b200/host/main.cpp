// haslr_assemble — drop-in for the reference binary of the same name (reference src/haslr_assemble/src/main.cpp).
// Same options (Commandline.cpp:68-242), same files in the output directory; stages (i)-(iii) run on the GPU.
#include <getopt.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "haslr.hpp"

using namespace haslr;

static double cpu_time() {
    struct rusage t; getrusage(RUSAGE_SELF, &t);
    return t.ru_utime.tv_sec + t.ru_utime.tv_usec / 1e6 + t.ru_stime.tv_sec + t.ru_stime.tv_usec / 1e6;
}
static double real_time() { struct timeval t; gettimeofday(&t, nullptr); return t.tv_sec + t.tv_usec / 1e6; }

static void help_short() { fprintf(stderr, "usage: haslr_assemble -c contig.fasta -l longread.fasta -m lr2contig.paf -d outdir [options]\n"); }
static void help(const Options& o) {
    help_short();
    fprintf(stderr, "\nRequired options:\n"
                    "    -c STR            Path to contigs file (also --contig)\n"
                    "    -l STR            Path to long read dataset (also --long)\n"
                    "    -m STR            Path to mappings of long reads onto contigs (also --mapping)\n"
                    "    -d STR            Path to the output directory (also --dir)\n\nAdvanced options:\n");
    fprintf(stderr, "    --aln-block       Minimum length of alignment block [%d]\n", o.min_aln_block);
    fprintf(stderr, "    --aln-sim         Minimum alignment similarity [%.2lf]\n", o.min_aln_sim);
    fprintf(stderr, "    --uniq-dev        Maximum deviation from mean frequency of uniq contigs [%.2lf]\n", o.max_uniq_dev);
    fprintf(stderr, "    --edge-sup        Minimum number of long read supporting each edge [%d]\n", o.min_edge_sup);
    fprintf(stderr, "\nOther options:\n"
                    "    -t INT            Number of CPU cores to use (also --threads)\n"
                    "    --gpus INT        Number of GPUs the consensus edges are sharded over [1]\n"
                    "    --no-logs         Do not write log_coordinate.txt / log_consensus.txt\n"
                    "    --long-fofn       The file passed by -l is fofn\n"
                    "    --mapping-fofn    The file passed by -m is fofn\n");
    fprintf(stderr, "    --version         Prints version (%s)\n    -h                Prints this help message (also --help)\n\n", o.prog_version.c_str());
}

static bool parse(int argc, char** argv, Options& o, bool& logs) {
    if (argc == 1) { help_short(); return false; }
    static struct option lo[] = {
        {"contig", required_argument, 0, 'c'}, {"long", required_argument, 0, 'l'}, {"mapping", required_argument, 0, 'm'},
        {"dir", required_argument, 0, 'd'}, {"help", no_argument, 0, 'h'}, {"threads", required_argument, 0, 't'},
        {"version", no_argument, 0, 0}, {"long-fofn", no_argument, 0, 0}, {"mapping-fofn", no_argument, 0, 0},
        {"aln-block", required_argument, 0, 0}, {"aln-sim", required_argument, 0, 0}, {"uniq-dev", required_argument, 0, 0},
        {"edge-sup", required_argument, 0, 0}, {"gpus", required_argument, 0, 0}, {"no-logs", no_argument, 0, 0}, {0, 0, 0, 0}};
    int c, idx;
    while ((c = getopt_long(argc, argv, "c:l:m:d:t:h", lo, &idx)) != -1) {
        switch (c) {
            case 'c': o.contig_path = optarg; break;
            case 'l': o.long_path = optarg; break;
            case 'm': o.mapping_path = optarg; break;
            case 'd': o.out_dir = optarg; break;
            case 't': {
                int v = atoi(optarg), hw = (int)std::thread::hardware_concurrency();
                o.num_threads = v < 1 ? 1 : (v > hw ? hw : v);
                break;
            }
            case 'h': help(o); exit(EXIT_SUCCESS);
            case 0:
                if (idx == 6) { fprintf(stdout, "%s\n", o.prog_version.c_str()); exit(EXIT_SUCCESS); }
                else if (idx == 7) o.long_fofn = true;
                else if (idx == 8) o.mapping_fofn = true;
                else if (idx == 9) { int v = atoi(optarg); o.min_aln_block = v < 0 ? 500 : v; }
                else if (idx == 10) { o.min_aln_sim = atof(optarg); if (o.min_aln_sim < 0 || o.min_aln_sim > 1) o.min_aln_sim = 0.85; }
                else if (idx == 11) o.max_uniq_dev = atof(optarg);
                else if (idx == 12) { int v = atoi(optarg); o.min_edge_sup = v < 0 ? 3 : v; }
                else if (idx == 13) { int v = atoi(optarg); o.gpus = v < 1 ? 1 : v; }
                else if (idx == 14) logs = false;
                else { help_short(); return false; }
                break;
            default: help_short(); return false;
        }
    }
    const char* need[4][2] = {{"-c", o.contig_path.c_str()}, {"-l", o.long_path.c_str()}, {"-m", o.mapping_path.c_str()}, {"-d", o.out_dir.c_str()}};
    for (auto& n : need) if (!*n[1]) { fprintf(stderr, "[ERROR] (CommandLine:parseCommandLine) option %s is required!\n", n[0]); help_short(); return false; }
    errno = 0;
    if (mkdir(o.out_dir.c_str(), S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == -1 && errno != EEXIST) return false;
    fprintf(stderr, "\n");
    return true;
}

int main(int argc, char** argv) {
    Options opt;
    bool logs = true;
    if (!parse(argc, argv, opt, logs)) return EXIT_FAILURE;
    const double c0 = cpu_time(), r0 = real_time();
    auto lap = [&]() { fprintf(stderr, "       elapsed time %.2lf CPU seconds (%.2lf real seconds)\n\n", cpu_time() - c0, real_time() - r0); };
    const std::string& d = opt.out_dir;

    // the CUDA contexts (seconds of driver start-up on a cold process) come up while the input files are read
    std::vector<hgpu_t*> ctxs((size_t)opt.gpus, nullptr);
    std::vector<int> ctx_rc((size_t)opt.gpus, HGPU_OK);
    std::thread ctx_thread([&]() {
        for (int i = 0; i < opt.gpus; ++i) ctx_rc[i] = hgpu_create(opt.gpus == 1 ? -1 : i, &ctxs[i]);
    });
    auto join_contexts = [&]() -> bool {
        if (ctx_thread.joinable()) ctx_thread.join();
        for (int i = 0; i < opt.gpus; ++i)
            if (ctx_rc[i] != HGPU_OK) {
                fprintf(stderr, "[ERROR] hgpu_create(device %d): %s — this build needs a B200-class GPU\n", i, hgpu_strerror(ctx_rc[i]));
                return false;
            }
        return true;
    };
    fprintf(stderr, "[NOTE] number of threads: %d, GPUs: %d\n\n", opt.num_threads, opt.gpus);

    fprintf(stderr, "[NOTE] loading contig sequences...\n");
    PathInputs in;
    ContigStore& contigs = in.contigs;
    load_fasta(opt.contig_path, contigs, &contigs);
    fprintf(stderr, "       loaded %zu contigs\n", contigs.size());
    lap();
    fprintf(stderr, "[NOTE] calculating kmer frequency of unique contigs\n");
    opt.uniq_freq = calc_uniq_freq(contigs);
    fprintf(stderr, "       mean: %.2lf\n", opt.uniq_freq);
    lap();

    fprintf(stderr, "[NOTE] loading long read sequences...\n");
    SeqStore& reads = in.reads;
    {
        std::vector<std::string> files;
        if (opt.long_fofn) load_fofn(opt.long_path, files); else files.push_back(opt.long_path);
        for (const auto& f : files) load_fasta(f, reads, nullptr);
    }
    fprintf(stderr, "       loaded %zu long reads\n", reads.size());
    lap();
    fprintf(stderr, "[NOTE] loading alignment between contigs and long reads...\n");
    {
        std::vector<std::string> files;
        if (opt.mapping_fofn) load_fofn(opt.mapping_path, files); else files.push_back(opt.mapping_path);
        for (const auto& f : files) read_text_file(f, in.paf_text);
    }
    fprintf(stderr, "       read %zu bytes of PAF text\n", in.paf_text.size());
    lap();

    // the clock of the backbone + POA path starts with the inputs in host memory and ends with every consensus
    // string in host memory. Tokenising, compact reads, edge table and coordinates run on the GPU with the hit table resident.
    fprintf(stderr, "[NOTE] PAF -> compact long reads -> backbone graph -> cleaning -> coordinates -> consensus (GPU)...\n");
    if (!join_contexts()) return EXIT_FAILURE;       // no CPU path: without the device the run ends here
    PathResult res;
    if (run_path(in, opt, ctxs, d, logs, res) != HGPU_OK) return EXIT_FAILURE;
    Graph& g = res.g;
    fprintf(stderr, "       %llu alignment rows, %zu edges; tokenise %.2f s, compact reads %.2f s, edge table %.2f s, cleaning %.2f s, coordinates %.2f s, consensus %.2f s\n",
            (unsigned long long)res.n_rows, res.edges.size(), res.t.tokenize, res.t.k1, res.t.k2, res.t.clean, res.t.coords, res.t.poa);
    lap();
    fprintf(stderr, "[NOTE] backbone + POA path (inputs in host memory -> consensus in host memory): %.1f long-read Mbases in %.2f s = %.1f Mbases/s\n\n",
            res.poa_bases / 1e6, res.t.total, res.t.total > 0 ? res.poa_bases / 1e6 / res.t.total : 0.0);

    fprintf(stderr, "[NOTE] generating the assembly from the cleaned backbone graph...\n");
    write_assembly(g, contigs, d);
    lap();
    for (hgpu_t* h : ctxs) hgpu_destroy(h);
    fprintf(stderr, "[NOTE] elapsed time %.2lf CPU seconds (%.2lf real seconds)\n\n*** BYE ***\n\n", cpu_time() - c0, real_time() - r0);
    return EXIT_SUCCESS;
}
