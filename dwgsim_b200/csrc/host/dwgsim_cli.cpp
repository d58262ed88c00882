// dwgsim_cli.cpp -- the host shell: a drop-in `dwgsim [options] <in.ref.fa> <out.prefix>` binary.
//
// What stays on the host is what must stay sequential and bit-identical to the reference (north star):
// option parsing (reference src/dwgsim_opt.c), the FASTA census and per-contig pair budget / skip rules
// (src/dwgsim.c:465-625), drand48-driven mutation generation (mut_diref, src/mut.c:591-758, with
// mut_left_justify :481-589) and the .mutations.txt/.vcf writers (mut_print, src/mut.c:781-893).  The read-pair
// loop (src/dwgsim.c:636-1099) is three calls into libdwgsim_b200.so (include/dwgsim_gpu.h), exactly the
// binding INTEGRATION.md describes.  FASTQ goes to <prefix>.bwa.read1/2.fastq.gz and <prefix>.bfast.fastq.gz like
// the reference (src/dwgsim.c:1149-1160): by default as gzip members written on the GPU (dwgsim_gpu_set_compression),
// with --host-gzip by a block-parallel zlib writer (concatenated members, level 6), or uncompressed with --uncompressed.
//
// New long options only (the short-option surface is the reference's): --uncompressed, --host-gzip, --threads N,
// --device D, --batch PAIRS.
#include <getopt.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <chrono>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/dwgsim_gpu.h"

#ifndef PACKAGE_VERSION
#define PACKAGE_VERSION "0.1.17-b200"
#endif

namespace {

// ---- glibc-compatible drand48 (seeded as src/dwgsim_opt.c:387-394 does) ---------------------------------------
struct Drand48 {
    uint64_t x = 0;
    void seed(long sv) { x = ((uint64_t)((sv >> 16) & 0xffff) << 32) | ((uint64_t)(sv & 0xffff) << 16) | 0x330e; }
    inline double next() { x = (x * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull; return std::ldexp((double)x, -48); }
} g_rng;

enum { ILLUMINA = 0, SOLID = 1, IONTORRENT = 2 };
enum : uint64_t { T_NOCHANGE = 0x00, T_INSERT = 0x10, T_SUBST = 0x20, T_DELETE = 0x30, TYPE_MASK = 0x30, BASE_TYPE_MASK = 0x3F };
constexpr int INS_SHIFT = 6, INS_LEN_SHIFT = 59, INS_SHORT_MAX = 26;
constexpr uint64_t INS_LEN_MASK = 0x1F, INS_PAYLOAD_MASK = (1ull << 52) - 1, INS_LONG_MAX = 0xFFFFFFFFull;

uint8_t g_nt4[256];
void nt4_init()
{
    memset(g_nt4, 4, sizeof g_nt4);
    g_nt4['A'] = g_nt4['a'] = 0; g_nt4['C'] = g_nt4['c'] = 1; g_nt4['G'] = g_nt4['g'] = 2; g_nt4['T'] = g_nt4['t'] = 3;
    g_nt4['-'] = 5;
}
const char kBase[8] = {'A', 'C', 'G', 'T', 'N', 0, 0, 0};

struct Options {                              // dwgsim_opt_t, src/dwgsim_opt.h:21-60
    double e_start[2] = {0.02, 0.02}, e_end[2] = {0.02, 0.02}, e_by[2] = {0, 0};
    int is_inner = 0, dist = 500;
    double std_dev = 50;
    long long N = -1;
    double C = 100;
    int length[2] = {70, 70};
    double mut_rate = 0.001, mut_freq = 0.5, indel_frac = 0.1, indel_extend = 0.3;
    int indel_min = 1;
    double rand_read = 0.05;
    int max_n = 0, data_type = ILLUMINA, strandedness = 0, read_one_strand = 0;
    std::string flow_order;                   // as given, then codes
    std::vector<int8_t> flow_codes;
    int use_base_error = 0, is_hap = 0, seed = -1;
    std::string fixed_quality, read_prefix, fn_muts_input, fn_regions_bed;
    bool has_prefix = false, has_fixed_quality = false;
    double quality_std = 2.0;
    int muts_input_type = -1, reads_output_type = 0, output_type = 0, amplicons = 0;
    // this build
    bool uncompressed = false, host_gzip = false;
    int threads = 0, device = 0, gpus = 0;       // gpus: 0 = $DWGSIM_GPUS or 1; N devices starting at `device`, -1 = all
    long long batch = 0;
};

int usage(const Options &o)
{
    fprintf(stderr, "\nProgram: dwgsim (short read simulator; read-pair loop on NVIDIA B200 via libdwgsim_b200)\n");
    fprintf(stderr, "Version: %s\n\n", PACKAGE_VERSION);
    fprintf(stderr, "Usage:   dwgsim [options] <in.ref.fa> <out.prefix>\n\n");
    fprintf(stderr, "Options (same letters, meaning and defaults as nh13/DWGSIM):\n");
    fprintf(stderr, "         -e FLOAT      per base/color/flow error rate of the first read, or START-END [%.3f]\n", o.e_start[0]);
    fprintf(stderr, "         -E FLOAT      per base/color/flow error rate of the second read, or START-END [%.3f]\n", o.e_start[1]);
    fprintf(stderr, "         -i            use the inner distance instead of the outer distance for pairs\n");
    fprintf(stderr, "         -d INT        distance between the two ends for pairs [%d]\n", o.dist);
    fprintf(stderr, "         -s INT        standard deviation of the distance for pairs [%.3f]\n", o.std_dev);
    fprintf(stderr, "         -N INT        number of read pairs (-1 to disable) [%lld]\n", o.N);
    fprintf(stderr, "         -C FLOAT      mean coverage across available positions (-1 to disable) [%.2f]\n", o.C);
    fprintf(stderr, "         -1 INT        length of the first read [%d]\n", o.length[0]);
    fprintf(stderr, "         -2 INT        length of the second read [%d]\n", o.length[1]);
    fprintf(stderr, "         -r FLOAT      rate of mutations [%.4f]\n", o.mut_rate);
    fprintf(stderr, "         -F FLOAT      frequency of given mutation (first haplotype) [%.4f]\n", o.mut_freq);
    fprintf(stderr, "         -R FLOAT      fraction of mutations that are indels [%.2f]\n", o.indel_frac);
    fprintf(stderr, "         -X FLOAT      probability an indel is extended [%.2f]\n", o.indel_extend);
    fprintf(stderr, "         -I INT        the minimum length indel [%d]\n", o.indel_min);
    fprintf(stderr, "         -y FLOAT      probability of a random DNA read [%.2f]\n", o.rand_read);
    fprintf(stderr, "         -n INT        maximum number of Ns allowed in a given read [%d]\n", o.max_n);
    fprintf(stderr, "         -c INT        generate reads for 0: Illumina, 1: SOLiD, 2: Ion Torrent [%d]\n", o.data_type);
    fprintf(stderr, "         -S INT        pair orientation 0: by platform, 1: same strand, 2: opposite strand [%d]\n", o.strandedness);
    fprintf(stderr, "         -A INT        read one 0: random strand, 1: forward, 2: reverse [%d]\n", o.read_one_strand);
    fprintf(stderr, "         -f STRING     the flow order for Ion Torrent data\n");
    fprintf(stderr, "         -B            use a per-base error rate for Ion Torrent data\n");
    fprintf(stderr, "         -H            haploid mode\n");
    fprintf(stderr, "         -z INT        random seed (-1 uses the current time) [%d]\n", o.seed);
    fprintf(stderr, "         -M INT        output 0: reads and mutations, 1: reads only, 2: mutations only [%d]\n", o.output_type);
    fprintf(stderr, "         -m FILE       the mutations txt file to re-create [%s]\n", o.muts_input_type != 0 ? "not using" : o.fn_muts_input.c_str());
    fprintf(stderr, "         -b FILE       the bed-like file set of candidate mutations [%s]\n", o.muts_input_type != 1 ? "not using" : o.fn_muts_input.c_str());
    fprintf(stderr, "         -v FILE       the vcf file set of candidate mutations (use pl tag for strand) [%s]\n", o.muts_input_type != 2 ? "not using" : o.fn_muts_input.c_str());
    fprintf(stderr, "         -x FILE       the bed of regions to cover [%s]\n", o.fn_regions_bed.empty() ? "not using" : o.fn_regions_bed.c_str());
    fprintf(stderr, "         -P STRING     a read prefix to prepend to each read name\n");
    fprintf(stderr, "         -q STRING     a fixed base quality to apply (single character)\n");
    fprintf(stderr, "         -Q FLOAT      standard deviation of the base quality scores [%.2f]\n", o.quality_std);
    fprintf(stderr, "         -o INT        FASTQ files 0: bfast and bwa, 1: bwa only, 2: bfast only [%d]\n", o.reads_output_type);
    fprintf(stderr, "         -a            assume each contig is an amplicon\n");
    fprintf(stderr, "         -h            print this message\n");
    fprintf(stderr, "         --uncompressed  write .fastq instead of .fastq.gz\n");
    fprintf(stderr, "         --host-gzip     compress with zlib on the host (level 6, like the reference) instead of on the GPU\n");
    fprintf(stderr, "         --threads INT   host gzip worker threads [all cores]\n");
    fprintf(stderr, "         --device INT    CUDA device [0]\n");
    fprintf(stderr, "         --gpus INT      devices to spread the read pairs over, starting at --device; -1: all [$DWGSIM_GPUS or 1]\n");
    fprintf(stderr, "         --batch INT     read pairs per device batch\n\n");
    return 1;
}

void parse_rate(const char *str, double *start, double *end)          // src/dwgsim_opt.c:162-179
{
    size_t i, n = strlen(str);
    *start = atof(str);
    for (i = 0; i < n; i++) if (str[i] == ',' || str[i] == '-') break;
    if (n > 0 && i < n - 1) *end = atof(str + i + 1); else *end = *start;
}
bool is_int(const char *s, bool neg_ok)                               // src/dwgsim_opt.c:181-192
{
    size_t len = strlen(s);
    if (!len) return false;
    if (s[0] != '+' && !(neg_ok && s[0] == '-') && !isdigit((unsigned char)s[0])) return false;
    for (size_t i = 1; i < len; i++) if (!isdigit((unsigned char)s[i])) return false;
    return true;
}
int to_int(const char *s, char flag, bool neg_ok)
{
    if (!is_int(s, neg_ok)) { fprintf(stderr, "Error: command line option -%c is not a number [%s]\n", flag, s); exit(1); }
    return atoi(s);
}

// ---- Ion Torrent flow model on the host: only for the -B calibration of the error rate, which draws from the
// same drand48 stream as mut_diref and therefore has to run here (src/dwgsim_opt.c:415-457; model src/dwgsim.c:246-417)
int flow_errors_host(const Options &o, std::vector<uint8_t> &seq, std::vector<uint8_t> &mask, int len, double e, int *n_err_out)
{
    const std::vector<int8_t> &fo = o.flow_codes;
    const int fl = (int)fo.size();
    auto grow = [&](int need) { if ((int)seq.size() <= need) seq.resize((size_t)need * 2 + 16); };
    for (int i = 0; i < len; i++) if (seq[i] >= 4) seq[i] = 0;
    int i, flow_i;
    for (i = 0; i < fl; i++) { int c = len > 0 ? seq[0] : 0; if (c == fo[i]) break; mask[i] = 0; }
    if (i == fl) return -1;
    flow_i = i;
    int prev_c = 4;
    for (i = 0; i < len; i++) {
        const int c = seq[i];
        while (c != fo[flow_i]) { mask[flow_i] = 0; flow_i = (flow_i + 1) % fl; }
        if (prev_c != c) {
            mask[flow_i] = 0;
            int n_err = 0;
            while (g_rng.next() < e) n_err++;
            if (n_err > 0) {
                if (g_rng.next() < 0.5) {
                    grow(len + n_err + 1);
                    for (int j = len - 1; i <= j; j--) seq[j + n_err] = seq[j];
                    for (int j = i; j < i + n_err; j++) seq[j] = (uint8_t)c;
                    len += n_err;
                } else {
                    int hp_l = 0, next_c = 4;
                    for (int j = i; j < len; j++, hp_l++) { next_c = seq[j]; if (c != next_c) break; }
                    if (hp_l < n_err) n_err = hp_l;
                    for (int j = i; j < len - n_err; j++) seq[j] = seq[j + n_err];
                    len -= n_err;
                    mask[flow_i] = 1;
                    if (n_err == hp_l && (i == 0 || prev_c == next_c)) {
                        int j = 0;
                        while (next_c != fo[(flow_i + j) % fl]) j++;
                        if (j <= 0) return len;
                        const int k = (int)(g_rng.next() * j);
                        grow(len + 2);
                        for (int jj = len - 1; i <= jj; jj--) seq[jj + 1] = seq[jj];
                        seq[i] = (uint8_t)fo[(flow_i + k) % fl];
                        len++;
                    }
                }
                *n_err_out += n_err;
            }
            prev_c = c;
        }
    }
    for (i = 0; i < len; i++) {
        const int c = seq[i];
        while (c != fo[flow_i]) {
            int n_err = 0;
            while (g_rng.next() < e) n_err++;
            if (mask[flow_i] == 0 && n_err > 0) {
                grow(len + n_err + 1);
                for (int j = len - 1; i <= j; j--) seq[j + n_err] = seq[j];
                for (int j = i; j < i + n_err; j++) seq[j] = (uint8_t)fo[flow_i];
                len += n_err;
                *n_err_out += n_err;
            }
            flow_i = (flow_i + 1) % fl;
        }
    }
    return len;
}

// returns 1 ok, 0 -> usage
int parse_options(Options &o, int argc, char **argv, int *first_arg)
{
    static const struct option longopts[] = {
        {"uncompressed", no_argument, nullptr, 1000}, {"threads", required_argument, nullptr, 1001},
        {"device", required_argument, nullptr, 1002}, {"gpus", required_argument, nullptr, 1006}, {"batch", required_argument, nullptr, 1003},
        {"host-gzip", no_argument, nullptr, 1004}, {nullptr, 0, nullptr, 0}};
    int c, muts = 0;
    while ((c = getopt_long(argc, argv, "id:s:N:C:1:2:e:E:r:F:R:X:I:c:S:A:n:y:BHf:z:M:m:b:v:x:P:q:Q:o:ah", longopts, nullptr)) >= 0) {
        switch (c) {
            case 'i': o.is_inner = 1; break;
            case 'd': o.dist = to_int(optarg, 'd', false); break;
            case 's': o.std_dev = atof(optarg); break;
            case 'N': o.N = to_int(optarg, 'N', true); o.C = -1; break;
            case 'C': o.C = atof(optarg); o.N = -1; break;
            case '1': o.length[0] = to_int(optarg, '1', false); break;
            case '2': o.length[1] = to_int(optarg, '2', false); break;
            case 'e': parse_rate(optarg, &o.e_start[0], &o.e_end[0]); break;
            case 'E': parse_rate(optarg, &o.e_start[1], &o.e_end[1]); break;
            case 'r': o.mut_rate = atof(optarg); break;
            case 'F': o.mut_freq = atof(optarg); break;
            case 'R': o.indel_frac = atof(optarg); break;
            case 'X': o.indel_extend = atof(optarg); break;
            case 'I': o.indel_min = to_int(optarg, 'I', false); break;
            case 'c': o.data_type = to_int(optarg, 'c', false); break;
            case 'S': o.strandedness = to_int(optarg, 'S', false); break;
            case 'A': o.read_one_strand = to_int(optarg, 'A', false); break;
            case 'n': o.max_n = to_int(optarg, 'n', false); break;
            case 'y': o.rand_read = atof(optarg); break;
            case 'f': o.flow_order = optarg; break;
            case 'B': o.use_base_error = 1; break;
            case 'H': o.is_hap = 1; break;
            case 'h': return 0;
            case 'z': o.seed = to_int(optarg, 'z', true); break;
            case 'M': o.output_type = to_int(optarg, 'M', false); break;
            case 'm': o.fn_muts_input = optarg; o.muts_input_type = 0; muts |= 1; break;
            case 'b': o.fn_muts_input = optarg; o.muts_input_type = 1; muts |= 2; break;
            case 'v': o.fn_muts_input = optarg; o.muts_input_type = 2; muts |= 4; break;
            case 'x': o.fn_regions_bed = optarg; break;
            case 'P': o.read_prefix = optarg; o.has_prefix = true; break;
            case 'q': o.fixed_quality = optarg; o.has_fixed_quality = true; break;
            case 'Q': o.quality_std = atof(optarg); break;
            case 'o': o.reads_output_type = atoi(optarg); break;
            case 'a': o.amplicons = 1; break;
            case 1000: o.uncompressed = true; break;
            case 1001: o.threads = atoi(optarg); break;
            case 1002: o.device = atoi(optarg); break;
            case 1006: o.gpus = atoi(optarg); break;
            case 1003: o.batch = atoll(optarg); break;
            case 1004: o.host_gzip = true; break;
            default: fprintf(stderr, "Unrecognized option: -%c\n", c); return 0;
        }
    }
    if (argc - optind < 2) return 0;
    *first_arg = optind;
#define CHECK(v, lo, hi, name) do { if ((v) < (lo) || (hi) < (v)) { fprintf(stderr, "Error: command line option %s was out of range\n", name); return 0; } } while (0)
    CHECK(o.dist, 0, INT32_MAX, "-d"); CHECK(o.std_dev, 0, INT32_MAX, "-s");
    if (o.N < 0 && o.C < 0) { fprintf(stderr, "Must use one of -N or -C"); return 0; }
    if (0 < o.N && 0 < o.C) { fprintf(stderr, "Cannot use both -N or -C"); return 0; }
    if (0 < o.N) { CHECK(o.N, 1, INT32_MAX, "-N"); CHECK(o.C, INT32_MIN, -1, "-C"); }
    else { CHECK(o.N, INT32_MIN, -1, "-N"); CHECK(o.C, 0, INT32_MAX, "-C"); }
    CHECK(o.length[0], 1, INT32_MAX, "-1"); CHECK(o.length[1], 0, INT32_MAX, "-2");
    for (int i = 0; i < 2; i++) {
        if (o.e_start[i] < 0.0 || 1.0 < o.e_start[i]) { fprintf(stderr, "End %s: the start error is out of range (-e)\n", i ? "two" : "one"); return 0; }
        if (o.e_end[i] < 0.0 || 1.0 < o.e_end[i]) { fprintf(stderr, "End %s: the end error is out of range (-e)\n", i ? "two" : "one"); return 0; }
        if (o.data_type == IONTORRENT && o.e_end[i] != o.e_start[i]) {
            fprintf(stderr, "End %s: a uniform error rate must be given for Ion Torrent data\n", i ? "two" : "one"); return 0;
        }
    }
    CHECK(o.mut_rate, 0, 1.0, "-r"); CHECK(o.indel_frac, 0, 1.0, "-R"); CHECK(o.indel_extend, 0, 1.0, "-X");
    CHECK(o.indel_min, 1, INT32_MAX, "-I"); CHECK(o.data_type, 0, 2, "-c"); CHECK(o.strandedness, 0, 2, "-S");
    CHECK(o.read_one_strand, 0, 2, "-A"); CHECK(o.max_n, 0, INT32_MAX, "-n"); CHECK(o.rand_read, 0, 1.0, "-y");
    if (o.data_type == IONTORRENT && o.flow_order.empty()) { fprintf(stderr, "Error: command line option -f is required\n"); return 0; }
    if (o.has_fixed_quality && o.fixed_quality.size() != 1) { fprintf(stderr, "Error: command line option -q requires one character\n"); return 0; }
    CHECK(o.quality_std, 0, INT32_MAX, "-Q");
    if (o.has_prefix) fprintf(stderr, "Warning: remember to use the -P option with dwgsim_eval\n");
    CHECK(o.reads_output_type, 0, 2, "-o");
    if (muts != 0 && muts != 1 && muts != 2 && muts != 4) { fprintf(stderr, "Error: -m/-b/-v cannot be used together\n"); return 0; }
#undef CHECK
    g_rng.seed(o.seed == -1 ? (long)time(nullptr) : (long)o.seed);
    if (o.data_type == IONTORRENT)
        for (char ch : o.flow_order) o.flow_codes.push_back((int8_t)g_nt4[(uint8_t)ch]);
    if (o.data_type == IONTORRENT && o.use_base_error == 1) {           // src/dwgsim_opt.c:415-457
        double sf = 0.0;
        for (int i = 0; i < 2; i++) {
            if (o.length[i] <= 0) continue;
            fprintf(stderr, "[dwgsim_core] Updating error rate for end %d\n", i + 1);
            if (0 < i && o.length[i] == o.length[1 - i]) {
                o.e_start[i] = o.e_start[1 - i]; o.e_end[i] = o.e_end[1 - i]; o.e_by[i] = o.e_by[1 - i];
                fprintf(stderr, "[dwgsim_core] Using scaling factor from previous end\n[dwgsim_core] Updated with scaling factor %.5lf\n", sf);
                continue;
            }
            std::vector<uint8_t> seq((size_t)o.length[i] * 2 + 64), mask(std::max<size_t>(o.flow_codes.size(), (size_t)o.length[i]) + 64, 0);
            long long n_err = 0, counts = 0;
            int j;
            for (j = 0; j < 1000000; j++) {
                if (j % 10000 == 0) fprintf(stderr, "\r[dwgsim_core] %d", j);
                for (int k = 0; k < o.length[i]; k++) seq[k] = (uint8_t)((int)(g_rng.next() * 4.0) & 3);
                int cur = 0;
                const int s = flow_errors_host(o, seq, mask, o.length[i], o.e_start[i], &cur);
                n_err += cur; counts += s;
            }
            sf = o.e_start[i] / ((int32_t)n_err / (1.0 * (int32_t)counts));
            o.e_start[i] = o.e_end[i] *= sf;
            o.e_by[i] = (o.e_end[i] - o.e_start[i]) / o.length[i];
            fprintf(stderr, "\r[dwgsim_core] %d\n[dwgsim_core] Updated with scaling factor %.5lf!\n", j, sf);
        }
    } else {
        o.e_by[0] = (o.e_end[0] - o.e_start[0]) / o.length[0];
        o.e_by[1] = (o.e_end[1] - o.e_start[1]) / o.length[1];
    }
    if (o.output_type < 0 || o.output_type > 2) { fprintf(stderr, "Error: command line option -M was out of range\n"); return 0; }
    if (o.amplicons == 1 && !o.fn_regions_bed.empty()) { fprintf(stderr, "Error: cannot use a regions BED file (-x) when simulating amplicons (-a)\n"); return 0; }
    return 1;
}

// ---- FASTA (seq_read_fasta, src/mut.c:49-87) over an in-memory file ------------------------------------------------
struct Fasta {                                 // seq_read_fasta, src/mut.c:49-87, over a read-only mapping of the file
    const char *buf = nullptr;
    size_t n = 0, pos = 0;
    bool mapped = false;
    std::vector<char> owned;
    uint8_t keep[256];
    ~Fasta() { if (mapped) munmap((void *)buf, n); }
    bool open(const char *path)
    {
        for (int c = 0; c < 256; c++) keep[c] = (isalpha(c) || c == '-' || c == '.') ? 1 : 0;     // src/mut.c:75
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                buf = (const char *)m; n = (size_t)st.st_size; mapped = true;
                ::close(fd);
                return true;
            }
        }
        FILE *fp = fdopen(fd, "rb");                               // pipes and empty files: read it all
        if (!fp) { ::close(fd); return false; }
        char tmp[1 << 16];
        size_t got;
        while ((got = fread(tmp, 1, sizeof tmp, fp)) > 0) owned.insert(owned.end(), tmp, tmp + got);
        fclose(fp);
        buf = owned.data(); n = owned.size();
        return true;
    }
    // next contig: its symbols (kept as in the file) into seq unless count_only; returns the length or -1
    int64_t next(std::vector<uint8_t> &seq, std::string &name, bool count_only = false)
    {
        while (pos < n && buf[pos] != '>') pos++;
        if (pos >= n) return -1;
        pos++;
        name.clear();
        int c = 0;
        while (pos < n) { c = (unsigned char)buf[pos++]; if (c == ' ' || c == '\t' || c == '\n') break; if (c != '\r') name.push_back((char)c); }
        if (c != '\n') while (pos < n && buf[pos++] != '\n') {}
        // the record ends at the next '>' (anywhere, like the reference's fgetc loop)
        const char *beg = buf + pos;
        const char *end = (const char *)memchr(beg, '>', n - pos);
        if (!end) end = buf + n;
        const size_t span = (size_t)(end - beg);
        pos += span;
        int64_t len = 0;
        if (count_only) {
            for (size_t i = 0; i < span; i++) len += keep[(unsigned char)beg[i]];
            return len;
        }
        seq.resize(span + 1);
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t T = span >= ((size_t)8 << 20) ? std::max<size_t>(1, std::min<size_t>(8, hw / 2)) : 1;
        if (T > 1) {
            // long records: every thread counts what its part of the text keeps, then filters it to its place
            std::vector<size_t> kept(T + 1, 0);
            auto part = [&](size_t t, size_t *a, size_t *b) { *a = span * t / T; *b = span * (t + 1) / T; };
            {
                std::vector<std::thread> th;
                for (size_t t = 0; t < T; t++) th.emplace_back([&, t]() { size_t a, b, c = 0; part(t, &a, &b); for (size_t i = a; i < b; i++) c += keep[(unsigned char)beg[i]]; kept[t + 1] = c; });
                for (auto &x : th) x.join();
            }
            for (size_t t = 0; t < T; t++) kept[t + 1] += kept[t];
            {
                std::vector<std::thread> th;
                uint8_t *base = seq.data();
                for (size_t t = 0; t < T; t++) th.emplace_back([&, t]() {
                    size_t a, b; part(t, &a, &b);
                    uint8_t *out = base + kept[t], *stop = base + kept[t + 1];
                    for (size_t i = a; i < b && out < stop; i++) { const unsigned char ch = (unsigned char)beg[i]; *out = ch; out += keep[ch]; }
                });
                for (auto &x : th) x.join();
            }
            len = (int64_t)kept[T];
            seq.resize((size_t)len);
            return len;
        }
        uint8_t *out = seq.data();
        for (size_t i = 0; i < span; i++) { const unsigned char ch = (unsigned char)beg[i]; *out = ch; out += keep[ch]; }   // branch-free filter
        len = (int64_t)(out - seq.data());
        seq.resize((size_t)len);
        return len;
    }
};

// ---- haplotype arrays: the reference's mutseq_t (src/mut.h:42-47), 64-bit mut_t per base ------------------------------
struct Hap {
    std::vector<uint64_t> s;
    std::vector<uint8_t *> ins;
    ~Hap() { for (auto p : ins) free(p); }
    // every s[i], i < l, is written by mut_diref before it is read: no zero fill (a recycled vector keeps its pages);
    // a fresh allocation asks for transparent huge pages first (8 B/base: first-touch faults were half of mut_diref's time)
    void reset(size_t l)
    {
        for (auto p : ins) free(p);
        ins.clear();
        if (s.capacity() < l + 2) {
            std::vector<uint64_t>().swap(s);
            s.reserve(l + 2);
            huge_pages(s.data(), s.capacity() * sizeof(uint64_t));
        }
        s.resize(l + 2);
        s[l] = s[l + 1] = 0;
    }
    static void huge_pages(void *p, size_t bytes)
    {
#ifdef MADV_HUGEPAGE
        if (bytes < (8u << 20)) return;
        const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, e = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
        if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
#else
        (void)p; (void)bytes;
#endif
    }
};
int long_ins_bytes(uint64_t n) { return 1 + (n <= 0xFF ? 1 : (n <= 0xFFFF ? 2 : 4)) + (int)((n + 3) >> 2); }
uint8_t *long_ins_payload(uint8_t *rec, uint32_t *n)
{
    if (rec[0] == 1) { *n = rec[1]; return rec + 2; }
    if (rec[0] == 2) { uint16_t v; memcpy(&v, rec + 1, 2); *n = v; return rec + 3; }
    uint32_t v; memcpy(&v, rec + 1, 4); *n = v; return rec + 5;
}
uint8_t *long_ins_set_len(uint8_t *rec, uint32_t n)
{
    if (n <= 0xFF) { rec[0] = 1; rec[1] = (uint8_t)n; return rec + 2; }
    if (n <= 0xFFFF) { uint16_t v = (uint16_t)n; rec[0] = 2; memcpy(rec + 1, &v, 2); return rec + 3; }
    rec[0] = 4; memcpy(rec + 1, &n, 4); return rec + 5;
}
bool get_ins(const Hap &h, int64_t i, uint64_t *n, uint64_t *ins)
{
    const uint64_t m = h.s[i];
    *n = (m >> INS_LEN_SHIFT) & INS_LEN_MASK;
    *ins = (m >> INS_SHIFT) & INS_PAYLOAD_MASK;
    return *n != 0;
}

// mut_add_ins with random bases, src/mut.c:282-377
// mut_add_ins, src/mut.c:282-377.  hap < 0: draw the ploidy; bases == nullptr: random bases (num == 0: draw the length too);
// bases given: their length counts and N / unknown letters become random bases
void add_insertion(const Options &o, Hap &h1, Hap &h2, int64_t i, uint64_t c, int hap = -1, const char *bases = nullptr, uint64_t num = 0)
{
    uint64_t ins = 0;
    if (!bases) {
        if (num == 0) do { num++; } while (num < INS_LONG_MAX && ((long long)num < o.indel_min || g_rng.next() < o.indel_extend));
    } else num = strlen(bases);
    if (INS_LONG_MAX < num) num = INS_LONG_MAX;
    if (hap < 0) {
        if (o.is_hap || g_rng.next() < 0.333333) hap = 3;
        else if (g_rng.next() < 0.5) hap = 1;
        else hap = 2;
    }
    if (num <= INS_SHORT_MAX) {
        if (!bases) for (uint64_t j = 0; j < num; j++) ins = (ins << 2) | (uint64_t)(g_rng.next() * 4.0);
        else for (int64_t j = (int64_t)num - 1; 0 <= j; --j) {
            int base = g_nt4[(unsigned char)bases[j]];
            if (base >= 4) base = (int)(g_rng.next() * 4.0);
            ins = (ins << 2) | (uint64_t)base;
        }
        const uint64_t v = (num << INS_LEN_SHIFT) | (ins << INS_SHIFT) | T_INSERT | c;
        if (hap & 1) h1.s[i] = v;
        if (hap & 2) h2.s[i] = v;
        return;
    }
    Hap *hs[2] = {&h1, &h2};
    uint8_t *pl[2] = {nullptr, nullptr};
    for (int x = 0; x < 2; x++) if (hap & (1 << x)) {
        uint8_t *rec = (uint8_t *)calloc((size_t)long_ins_bytes(num), 1);
        hs[x]->ins.push_back(rec);
        pl[x] = long_ins_set_len(rec, (uint32_t)num);
    }
    int byte_i = 0, bit_i = 0;
    for (uint64_t left = num; left > 0; left--) {
        int base;
        if (!bases) base = (int)(g_rng.next() * 4.0);
        else { base = g_nt4[(unsigned char)bases[left - 1]]; if (base >= 4) base = (int)(g_rng.next() * 4.0); }
        const uint8_t b = (uint8_t)(base << (bit_i << 1));
        if (pl[0]) pl[0][byte_i] |= b;
        if (pl[1]) pl[1][byte_i] |= b;
        if (++bit_i == 4) { bit_i = 0; byte_i++; }
    }
    for (int x = 0; x < 2; x++) if (hap & (1 << x))
        hs[x]->s[i] = ((uint64_t)(hs[x]->ins.size() - 1) << INS_SHIFT) | T_INSERT | c;
}

// mut_left_justify_ins, src/mut.c:427-478
void left_justify_ins(Hap &h, int64_t i, std::vector<int64_t> *touched = nullptr)
{
    uint64_t n, ins;
    int64_t j = i;
    if (get_ins(h, i, &n, &ins)) {
        while (0 < j && (h.s[j - 1] & TYPE_MASK) == T_NOCHANGE && ((ins >> ((n - 1) << 1)) & 3) == (h.s[j - 1] & 3)) {
            ins &= ~((uint64_t)3 << ((n - 1) << 1));
            ins = (ins << 2) | (h.s[j - 1] & 3);
            h.s[j] &= 3;
            j--;
        }
        h.s[j] = (n << INS_LEN_SHIFT) | (ins << INS_SHIFT) | T_INSERT | (h.s[j] & 3);
        if (touched && j != i) touched->push_back(j);
        return;
    }
    uint32_t num;
    uint8_t *p = long_ins_payload(h.ins[ins], &num);
    const int nb = (int)((num + 3) >> 2);
    while (0 < j && (h.s[j - 1] & TYPE_MASK) == T_NOCHANGE && (uint64_t)(p[0] & 3) == (h.s[j - 1] & 3)) {
        for (int b = 0; b < nb; b++) { p[b] >>= 2; if (b + 1 < nb) p[b] |= (uint8_t)((p[b + 1] & 3) << 6); }
        p[nb - 1] |= (uint8_t)((h.s[j - 1] & 3) << (((num + 3) & 3) << 1));
        h.s[j] &= 3;
        j--;
    }
    h.s[j] = (ins << INS_SHIFT) | T_INSERT | (h.s[j] & 3);
    if (touched && j != i) touched->push_back(j);
}

void shift_del(Hap &h, int64_t i, int del, std::vector<int64_t> *touched = nullptr)   // src/mut.c:540-553
{
    for (int64_t j = i - 1; j >= 0; j--) {
        if ((h.s[j] & TYPE_MASK) == T_NOCHANGE && (h.s[j] & 3) == (h.s[j + del] & 3)) {
            const uint64_t t = h.s[j];
            h.s[j] = h.s[j + del];
            h.s[j + del] = (t | TYPE_MASK) ^ TYPE_MASK;
            if (touched) touched->push_back(j);
        } else break;
    }
}

// mut_left_justify, src/mut.c:481-589
// mut_left_justify, src/mut.c:482-589.  `events` (optional): the ascending positions that are not NOCHANGE in either
// haplotype when mut_diref's loop ends.  The reference visits every base; all it does at a NOCHANGE position with an
// A/C/G/T base is to clear prev_del, so visiting the events (and looking into a gap only while prev_del is set) is the
// same pass.  `touched` collects the positions the pass writes mutations to (they lie left of the visited event).
void left_justify(const std::vector<uint8_t> &seq, Hap &h1, Hap &h2, const std::vector<int64_t> *events = nullptr,
                  std::vector<int64_t> *touched = nullptr)
{
    const int64_t l = (int64_t)seq.size();
    int prev_del[2] = {0, 0};
    auto step = [&](int64_t i) {
        const uint64_t c0 = g_nt4[seq[i]], c1 = h1.s[i], c2 = h2.s[i];
        if (c0 >= 4) return;
        const uint64_t t1 = c1 & TYPE_MASK, t2 = c2 & TYPE_MASK;
        if (t1 == T_NOCHANGE && t2 == T_NOCHANGE) { prev_del[0] = prev_del[1] = 0; return; }
        int64_t j;
        int del;
        if ((c1 & BASE_TYPE_MASK) == (c2 & BASE_TYPE_MASK)) {
            if (t1 == T_SUBST) prev_del[0] = prev_del[1] = 0;
            else if (t1 == T_DELETE) {
                if (prev_del[0] == 1 || prev_del[1] == 1) return;
                prev_del[0] = prev_del[1] = 1;
                for (j = i + 1, del = 1; j < l && (h1.s[j] & TYPE_MASK) == T_DELETE; j++) del++;
                if (l <= i + del || i == 0) return;
                for (j = i - 1; j >= 0; j--) {
                    if ((h1.s[j] & TYPE_MASK) != T_INSERT && (h2.s[j] & TYPE_MASK) != T_INSERT && (h1.s[j] & TYPE_MASK) != T_DELETE &&
                        (h2.s[j] & TYPE_MASK) != T_DELETE && (h1.s[j] & 3) == (h1.s[j + del] & 3) && (h2.s[j] & 3) == (h2.s[j + del] & 3)) {
                        uint64_t t = h1.s[j]; h1.s[j] = h1.s[j + del]; h1.s[j + del] = (t | TYPE_MASK) ^ TYPE_MASK;
                        t = h2.s[j]; h2.s[j] = h2.s[j + del]; h2.s[j + del] = (t | TYPE_MASK) ^ TYPE_MASK;
                        if (touched) touched->push_back(j);
                    } else break;
                }
            } else { prev_del[0] = prev_del[1] = 0; left_justify_ins(h1, i, touched); left_justify_ins(h2, i, touched); }
        } else {
            if (t1 == T_SUBST || t2 == T_SUBST) prev_del[0] = prev_del[1] = 0;
            else if (t1 == T_DELETE) {
                if (prev_del[0] == 1) return;
                prev_del[0] = 1;
                for (j = i + 1, del = 1; j < l && (h1.s[j] & TYPE_MASK) == T_DELETE; j++) del++;
                if (l <= i + del || i == 0) return;
                shift_del(h1, i, del, touched);
            } else if (t2 == T_DELETE) {
                if (prev_del[1] == 1) return;
                prev_del[1] = 1;
                for (j = i + 1, del = 1; j < l && (h2.s[j] & TYPE_MASK) == T_DELETE; j++) del++;
                if (l <= i + del || i == 0) return;
                shift_del(h2, i, del, touched);
            } else if (t1 == T_INSERT) { prev_del[0] = prev_del[1] = 0; left_justify_ins(h1, i, touched); }
            else if (t2 == T_INSERT) { prev_del[0] = prev_del[1] = 0; left_justify_ins(h2, i, touched); }
        }
    };
    if (!events) { for (int64_t i = 0; i < l; ++i) step(i); return; }
    int64_t last = -1;
    for (const int64_t i : *events) {
        if ((prev_del[0] | prev_del[1]) && i != last + 1)            // the NOCHANGE bases in between: an A/C/G/T one clears prev_del
            for (int64_t j = last + 1; j < i; ++j) if (g_nt4[seq[j]] < 4) { prev_del[0] = prev_del[1] = 0; break; }
        step(i);
        last = i;
    }
}

// mut_diref, random branch, src/mut.c:591-643 + :752-757.  The common case (no mutation at this base) is one LCG
// step and one integer compare: drand48() < r  <=>  X < ceil(r * 2^48) for the 48-bit state X.
// LCG jump-ahead: x_{n+k} = kJumpA[k] x_n + kJumpC[k] (mod 2^48) for the drand48 recurrence
struct LcgJump {
    uint64_t a[9], c[9];
    LcgJump()
    {
        a[0] = 1; c[0] = 0;
        for (int k = 1; k <= 8; k++) { a[k] = (a[k - 1] * 0x5DEECE66Dull) & 0xFFFFFFFFFFFFull; c[k] = (c[k - 1] * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull; }
    }
};
const LcgJump kJump;

// mut_diref, random branch (src/mut.c:606-643), then mut_left_justify.  `events` receives the ascending positions that
// end up mutated (every base of a deletion), `touched` the positions mut_left_justify moved mutations to.
void diref(const Options &o, const std::vector<uint8_t> &seq, Hap &h1, Hap &h2, std::vector<int64_t> *events = nullptr,
           std::vector<int64_t> *touched = nullptr)
{
    const int64_t l = (int64_t)seq.size();
    h1.reset((size_t)l); h2.reset((size_t)l);
    Hap *ret[2] = {&h1, &h2};
    const uint64_t thr = (uint64_t)std::ceil(std::ldexp(o.mut_rate, 48));
    int deleting = 0, del_len = 0;
    std::vector<int64_t> local;
    std::vector<int64_t> &ev = events ? *events : local;
    ev.clear();
    if (touched) touched->clear();
    const uint8_t *sq = seq.data();
    uint64_t *s1 = h1.s.data(), *s2 = h2.s.data();
    for (int64_t i = 0; i < l; ++i) {
        if (!deleting && i + 8 <= l) {
            // eight A/C/G/T bases none of whose draws (one each: `drand48() < mut_rate`) comes out below the rate: the
            // common case, decided with eight independent jump-ahead steps instead of a chain of eight
            uint64_t c8[8];
            uint64_t bad = 0;
            for (int k = 0; k < 8; k++) { c8[k] = g_nt4[sq[i + k]]; bad |= c8[k]; }
            if (bad < 4) {
                const uint64_t x = g_rng.x;
                uint64_t hit = 0;
                for (int k = 1; k <= 8; k++) hit |= (uint64_t)(((kJump.a[k] * x + kJump.c[k]) & 0xFFFFFFFFFFFFull) < thr);
                if (!hit) {
                    for (int k = 0; k < 8; k++) { s1[i + k] = c8[k]; s2[i + k] = c8[k]; }
                    g_rng.x = (kJump.a[8] * x + kJump.c[8]) & 0xFFFFFFFFFFFFull;
                    i += 7;
                    continue;
                }
            }
        }
        uint64_t c = s1[i] = s2[i] = (uint64_t)g_nt4[sq[i]];
        if (deleting) {
            if (del_len < o.indel_min || g_rng.next() < o.indel_extend) {
                if (deleting & 1) s1[i] |= T_DELETE | c;
                if (deleting & 2) s2[i] |= T_DELETE | c;
                ev.push_back(i);
                del_len++;
                continue;
            }
            deleting = del_len = 0;
        }
        if (c >= 4) continue;
        g_rng.x = (g_rng.x * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
        if (g_rng.x >= thr) continue;                                      // drand48() < mut_rate is false
        ev.push_back(i);
        if (g_rng.next() >= o.indel_frac) {
            const double r = g_rng.next();
            c = (c + (uint64_t)(r * 3.0 + 1)) & 3;
            if (o.is_hap || g_rng.next() < 0.333333) s1[i] = s2[i] = T_SUBST | c;
            else ret[g_rng.next() < 0.5 ? 0 : 1]->s[i] = T_SUBST | c;
        } else if (g_rng.next() < 0.5) {
            if (o.is_hap || g_rng.next() < 0.3333333) { s1[i] = s2[i] = T_DELETE | c; deleting = 3; }
            else { deleting = g_rng.next() < 0.5 ? 1 : 2; ret[deleting - 1]->s[i] = T_DELETE | c; }
            del_len = 1;
        } else add_insertion(o, h1, h2, i, c);
    }
    left_justify(seq, h1, h2, events ? &ev : nullptr, touched);         // (without a caller's list: the reference's full pass)
    if (events && touched && !touched->empty()) {                        // candidates for the writers: events + touched, ascending
        ev.insert(ev.end(), touched->begin(), touched->end());
        std::sort(ev.begin(), ev.end());
        ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
    }
}

// mut_diref for long contigs on several threads, same bytes as diref() above.
// The draws of mut_diref are the iterates X_1, X_2, ... of one LCG; which of them are BELOW the mutation rate does not
// depend on the bases, so that question is answered for a whole range of iterates in parallel (phase B: thread t jumps to
// its sub-range and lists the hits with the state before each).  Filling the dense arrays with the plain bases (16 B per
// base, the memory-bound bulk of the work) is parallel too (phase A, which also counts the A/C/G/T bases per block: an
// N consumes no draw).  What stays sequential is one step per mutation (phase C): from (base i, n draws consumed) the
// next mutated base is the k-th A/C/G/T base from i on, k = next hit - n; the mutation itself -- its extra draws and a
// deletion's extension -- runs through the scalar code of diref() with the generator set to the state before the hit.
struct DirefHit { uint64_t n, x_before; };                      // iterate index (1-based) and the state before it
uint64_t lcg_jump(uint64_t x, uint64_t n)                        // x advanced by n steps
{
    uint64_t a = 0x5DEECE66Dull, c = 0xBull;
    const uint64_t M = 0xFFFFFFFFFFFFull;
    for (; n; n >>= 1) {
        if (n & 1) x = (a * x + c) & M;
        c = (a * c + c) & M; a = (a * a) & M;
    }
    return x;
}

void diref_parallel(const Options &o, const std::vector<uint8_t> &seq, Hap &h1, Hap &h2, std::vector<int64_t> *events,
                    std::vector<int64_t> *touched, int n_threads)
{
    const int64_t l = (int64_t)seq.size();
    h1.reset((size_t)l); h2.reset((size_t)l);
    Hap *ret[2] = {&h1, &h2};
    const uint64_t thr = (uint64_t)std::ceil(std::ldexp(o.mut_rate, 48)), M = 0xFFFFFFFFFFFFull;
    std::vector<int64_t> local;
    std::vector<int64_t> &ev = events ? *events : local;
    ev.clear();
    if (touched) touched->clear();
    const uint8_t *sq = seq.data();
    uint64_t *s1 = h1.s.data(), *s2 = h2.s.data();
    constexpr int kBlk = 256;                                    // bases per block of the A/C/G/T census
    const int64_t nblk = (l + kBlk - 1) / kBlk;
    std::vector<uint32_t> blk_acgt((size_t)nblk);
    const int T = std::max(1, n_threads);
    auto parallel = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto &x : th) x.join();
    };
    const bool timing = getenv("DWGSIM_DIREF_TIMING") != nullptr;
    auto clk = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tA = clk();
    // phase A: plain bases into both haplotypes, A/C/G/T census per block
    parallel([&](int t) {
        const int64_t b0 = nblk * t / T, b1 = nblk * (t + 1) / T;
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t i0 = b * kBlk, i1 = std::min<int64_t>(l, i0 + kBlk);
            uint32_t cnt = 0;
            for (int64_t i = i0; i < i1; ++i) { const uint64_t c = g_nt4[sq[i]]; s1[i] = c; s2[i] = c; cnt += c < 4; }
            blk_acgt[(size_t)b] = cnt;
        }
    });
    uint64_t total_acgt = 0;
    for (uint32_t c : blk_acgt) total_acgt += c;
    // phase B: the hits among the iterates 1 .. n_max (more are listed on demand: deletions draw on N bases too, and every
    // mutation draws a few times)
    const uint64_t x0 = g_rng.x;
    std::vector<DirefHit> hits;
    uint64_t listed = 0;                                         // iterates 1 .. listed are covered by `hits`
    auto list_hits = [&](uint64_t upto) {
        if (upto <= listed) return;
        const uint64_t from = listed, span = upto - from;
        std::vector<std::vector<DirefHit>> part((size_t)T);
        parallel([&](int t) {
            const uint64_t a = from + span * (uint64_t)t / (uint64_t)T, b = from + span * (uint64_t)(t + 1) / (uint64_t)T;
            uint64_t x = lcg_jump(x0, a), n = a;
            std::vector<DirefHit> &out = part[(size_t)t];
            while (n + 8 <= b) {                                 // eight independent jump-ahead steps at a time
                uint64_t y[9];
                y[0] = x;
                uint64_t any = 0;
                for (int k = 1; k <= 8; k++) { y[k] = (kJump.a[k] * x + kJump.c[k]) & M; any |= (uint64_t)(y[k] < thr); }
                if (any) for (int k = 1; k <= 8; k++) if (y[k] < thr) out.push_back(DirefHit{n + (uint64_t)k, y[k - 1]});
                x = y[8]; n += 8;
            }
            for (; n < b; ++n) { const uint64_t y = (x * 0x5DEECE66Dull + 0xBull) & M; if (y < thr) out.push_back(DirefHit{n + 1, x}); x = y; }
        });
        for (auto &v : part) hits.insert(hits.end(), v.begin(), v.end());
        listed = upto;
    };
    const double tB = clk();
    list_hits(total_acgt + total_acgt / 64 + 4096);
    const double tC = clk();
    // phase C: one step per mutation
    int64_t i = 0;                                               // next base to look at
    uint64_t n = 0;                                              // draws consumed so far
    size_t hp = 0;
    // the k-th (k >= 1) A/C/G/T base at or after i, or -1 when fewer remain
    auto kth_acgt = [&](int64_t from, uint64_t k) -> int64_t {
        int64_t p = from;
        // the rest of from's block, then whole blocks, then inside the block that holds it
        while (p < l) {
            const int64_t b = p / kBlk, bend = std::min<int64_t>(l, (b + 1) * kBlk);
            if (p == b * kBlk) {
                const uint32_t c = blk_acgt[(size_t)b];
                if (c < k) { k -= c; p = bend; continue; }
                if (c == (uint32_t)(bend - p)) return p + (int64_t)k - 1;        // no N in the block
            }
            for (; p < bend; ++p) if (g_nt4[sq[p]] < 4 && --k == 0) return p;
        }
        return -1;
    };
    for (;;) {
        while (hp < hits.size() && hits[hp].n <= n) ++hp;
        if (hp == hits.size()) {
            // no listed hit left: either the contig ends first, or more iterates have to be looked at (a long insertion or
            // deletion can also have drawn past the listed range)
            uint64_t rest = 0;
            for (int64_t p = i; p < l; ++p) rest += g_nt4[sq[p]] < 4;
            if (rest == 0) break;
            if (listed >= n + rest) { n += rest; break; }        // the remaining bases draw listed iterates: none of them hits
            if (listed < n) listed = n;                          // (iterates already consumed need no listing)
            list_hits(std::min<uint64_t>(n + rest, listed + std::max<uint64_t>(listed / 8, 1 << 16)));
            continue;
        }
        const int64_t t = kth_acgt(i, hits[hp].n - n);
        if (t < 0) { for (int64_t p = i; p < l; ++p) n += g_nt4[sq[p]] < 4; break; }      // the contig ends before the hit
        // base t draws iterate hits[hp].n, which is below the rate: the scalar code of diref() from here, until the mutation
        // (and the deletion it may start) is over
        g_rng.x = hits[hp].x_before;
        uint64_t drawn = hits[hp].n - 1;
        auto next = [&]() { ++drawn; return g_rng.next(); };
        int deleting = 0, del_len = 0;
        int64_t p = t;
        for (; p < l; ++p) {
            uint64_t c = (uint64_t)g_nt4[sq[p]];
            if (deleting) {
                if (del_len < o.indel_min || next() < o.indel_extend) {
                    if (deleting & 1) s1[p] |= T_DELETE | c;
                    if (deleting & 2) s2[p] |= T_DELETE | c;
                    ev.push_back(p);
                    del_len++;
                    continue;
                }
                deleting = del_len = 0;
            }
            if (p != t) break;                                   // the mutation is over: back to skipping along the hit list
            next();                                              // the draw that hit
            ev.push_back(p);
            if (next() >= o.indel_frac) {
                const double r = next();
                c = (c + (uint64_t)(r * 3.0 + 1)) & 3;
                if (o.is_hap || next() < 0.333333) s1[p] = s2[p] = T_SUBST | c;
                else ret[next() < 0.5 ? 0 : 1]->s[p] = T_SUBST | c;
            } else if (next() < 0.5) {
                if (o.is_hap || next() < 0.3333333) { s1[p] = s2[p] = T_DELETE | c; deleting = 3; }
                else { deleting = next() < 0.5 ? 1 : 2; ret[deleting - 1]->s[p] = T_DELETE | c; }
                del_len = 1;
            } else {
                const uint64_t before = g_rng.x;
                add_insertion(o, h1, h2, p, c);
                for (uint64_t y = before; y != g_rng.x; y = (y * 0x5DEECE66Dull + 0xBull) & M) ++drawn;   // its draws, counted
            }
        }
        i = p; n = drawn;
    }
    g_rng.x = lcg_jump(x0, n);
    const double tD = clk();
    left_justify(seq, h1, h2, events ? &ev : nullptr, touched);
    if (events && touched && !touched->empty()) {
        ev.insert(ev.end(), touched->begin(), touched->end());
        std::sort(ev.begin(), ev.end());
        ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
    }
    if (timing) fprintf(stderr, "[diref_parallel] %lld bases, %d threads: fill %.3f s, hits %.3f s (%zu), mutations %.3f s, justify %.3f s\n",
                        (long long)l, T, tB - tA, tC - tB, hits.size(), tD - tC, clk() - tD);
}

// ---- mutations to replay: -m TXT (src/mut_txt.c), -b BED (src/mut_bed.c), -v VCF (src/mut_vcf.c) ------------------------
struct ContigList { std::vector<std::string> name; std::vector<uint32_t> len; };
struct MutRec {
    int32_t contig; uint32_t start, end;      // BED: zero-based start, end; TXT / VCF: start = one-based position
    int type, is_hap;                         // T_SUBST / T_INSERT / T_DELETE; TXT / VCF: haplotype bits
    std::string bases; bool has_bases_ptr;    // (VCF deletions carry no bases)
};
struct MutsInput {
    int kind = -1;                            // 0 TXT, 1 BED, 2 VCF (the order of the -m / -b / -v switches)
    std::vector<MutRec> recs;
};
[[noreturn]] void die(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fflush(nullptr);
    _exit(1);                                  // (may run on the producer thread: no static destructors under the consumer's feet)
}
char iupac_to_mut(char iupac, char base)                                     // src/dwgsim.c:202-213
{
    static const char *codes = "XACMGRSVTWYHKDBN";
    const int b = g_nt4[(unsigned char)base];
    for (int i = 0; i < 4; i++) if (codes[(1 << (b & 3)) | (1 << (i & 3))] == iupac) return "ACGTN"[i];
    return 'X';
}
int muttype_of(char *str)                                                    // src/dwgsim.c:183-200
{
    for (char *p = str; *p; ++p) *p = (char)tolower((unsigned char)*p);
    const std::string t = str;
    if (t == "snp" || t == "substitute" || t == "sub" || t == "s") return (int)T_SUBST;
    if (t == "insertion" || t == "insert" || t == "ins" || t == "i") return (int)T_INSERT;
    if (t == "deletion" || t == "delet" || t == "del" || t == "d") return (int)T_DELETE;
    return -1;
}
void read_muts_txt(FILE *fp, const ContigList &c, MutsInput &M)              // src/mut_txt.c:39-128
{
    char name[1024], mut[1024], ref;
    uint32_t pos, prev_pos = 0, is_hap;
    size_t i = 0;
    while (0 < fscanf(fp, "%1023s\t%u\t%c\t%1023s\t%d", name, &pos, &ref, mut, &is_hap)) {
        while (i < c.name.size() && c.name[i] != name) { i++; prev_pos = 0; }
        if (i == c.name.size()) die("Error: mutation contig not found or out of order [%s]\n", name);
        if (pos <= 0 || c.len[i] < pos) die("Error: start out of range [%s,%u]\n", name, pos);
        if (pos < prev_pos) die("Error: out of order [%s,%u]\n", name, pos);
        MutRec r{(int32_t)i, pos, 0, 0, (int)is_hap, mut, true};
        if ('-' == ref && '-' != mut[0]) r.type = (int)T_INSERT;
        else if ('-' != ref && '-' == mut[0]) r.type = (int)T_DELETE;
        else if ('-' != ref && '-' != mut[0]) {
            r.type = (int)T_SUBST;
            if (is_hap < 3) {                                                 // heterozygous: IUPAC code of reference + alternative
                if (g_nt4[(unsigned char)r.bases[0]] < 4) die("Error: heterozygous bases must be in IUPAC form\n");
                const char b = iupac_to_mut(r.bases[0], ref);
                if ('X' == b) die("Error: out of range\n");
                r.bases.assign(1, b);
            }
        } else die("Error: out of range\n");
        M.recs.push_back(r);
    }
}
void read_muts_bed(FILE *fp, const ContigList &c, MutsInput &M)              // src/mut_bed.c:37-137
{
    char name[1024], type[1024], bases[1024];
    uint32_t start, end, prev_contig = 0, max_end = 0;
    size_t i = 0;
    while (0 < fscanf(fp, "%1023s\t%u\t%u\t%1023s\t%1023s", name, &start, &end, bases, type)) {
        while (i < c.name.size() && c.name[i] != name) i++;
        if (i == c.name.size()) die("Error: contig not found [%s]\n", name);
        if (c.len[i] <= start) die("Error: start out of range [%s,%u]\n", name, start);
        if (c.len[i] < end) die("Error: end out of range [%s,%u]\n", name, end);
        if (end <= start) die("Error: end <= start [%s,%u,%u]\n", name, start, end);
        if (0 != strcmp("*", bases) && (end - start) != strlen(bases)) die("Error: bases did not match start and end [%s,%u,%u,%s]\n", name, start, end, bases);
        if (prev_contig == (uint32_t)i && start + 1 <= max_end) {
            fprintf(stderr, "Warning: overlapping entries, ignoring entry [%s\t%u\t%u\t%s\t%s]\n", name, start, end, bases, type);
            continue;
        }
        if (prev_contig != (uint32_t)i || max_end < end) { prev_contig = (uint32_t)i; max_end = end; }
        const std::string type_in = type;
        const int t = muttype_of(type);
        if (t == (int)T_INSERT && INS_SHORT_MAX < end - start)
            die("Error: insertion of length %d exceeded the maximum supported length of %d\n", (int)(end - start), (int)INS_SHORT_MAX);
        if (t < 0) die("Error: mutation type unrecognized [%s]\n", type);
        M.recs.push_back(MutRec{(int32_t)i, start, end, t, 0, bases, true});
    }
}
void read_muts_vcf(FILE *fp, const ContigList &c, MutsInput &M)              // src/mut_vcf.c:42-278, line by line
{
    static int warned = 0;
    std::string line;
    char name[1024] = "", id[1024] = "", ref[1024] = "", alt[1025] = "";
    uint32_t pos = 0, prev_pos = 0;
    size_t i = 0;
    int ch;
    do {
        line.clear();
        while (EOF != (ch = fgetc(fp)) && ch != '\n' && ch != '\r') line.push_back((char)ch);
        if (line.empty() || line[0] == '#') continue;
        if (EOF == sscanf(line.c_str(), "%1023s\t%u\t%1023s\t%1023s\t%1024s", name, &pos, id, ref, alt)) die("Error: VCF parsing error\n");
        uint32_t is_hap = 4;
        for (size_t s = 0; s + 4 < line.size(); s++)                         // [\t;]pl=[1-3]
            if (('\t' == line[s] || ';' == line[s]) && 'p' == line[s + 1] && 'l' == line[s + 2] && '=' == line[s + 3]) {
                switch (line[s + 4]) {
                    case '1': is_hap = 1; break;
                    case '2': is_hap = 2; break;
                    case '3': is_hap = 3; break;
                    default: die("Error: Could not determine the strand of the mutation from the 'pl' tag.\n");
                }
                break;
            }
        if (4 == is_hap && 0 == warned) {                                     // NB: later untagged records keep is_hap = 4 (no haplotype)
            fprintf(stderr, "Warning: strand of the mutation not found; please use the 'pl' tag.\n");
            warned = 1; is_hap = 3;
        }
        while (i < c.name.size() && c.name[i] != name) { i++; prev_pos = 0; }
        if (i == c.name.size()) die("Error: contig not found [%s]\n", name);
        if (pos <= 0 || c.len[i] < pos) die("Error: start out of range [%s,%u]\n", name, pos);
        if (pos < prev_pos) die("Error: out of order [%s,%u]\n", name, pos);
        int ref_l = (int)strlen(ref), alt_l = (int)strlen(alt), j;
        if (1 == ref_l && ref[0] == '.') { ref[0] = 0; ref_l = 0; }
        if (1 == alt_l && alt[0] == '.') { alt[0] = 0; alt_l = 0; }
        if (0 == alt_l && 0 == ref_l) die("Error: empty alleles\n");
        for (j = 0; j < alt_l; j++) if (',' == alt[j]) die("Error: multiple alleles are not supported\n");
        for (j = 0; j < ref_l; j++) { ref[j] = "ACGTN"[std::min<int>(g_nt4[(unsigned char)ref[j]], 4)]; if ('N' == ref[j]) die("Error: non-ACGT base found\n"); }
        for (j = 0; j < alt_l; j++) { alt[j] = "ACGTN"[std::min<int>(g_nt4[(unsigned char)alt[j]], 4)]; if ('N' == alt[j]) die("Error: non-ACGT base found\n"); }
        if (ref_l == alt_l) {
            for (j = 0; j < ref_l; j++) M.recs.push_back(MutRec{(int32_t)i, pos + (uint32_t)j, 0, (int)T_SUBST, (int)is_hap, std::string(1, alt[j]), true});
        } else if (ref_l < alt_l) {
            for (j = 0; j < ref_l; j++, pos++) if (ref[j] != alt[j]) break;
            M.recs.push_back(MutRec{(int32_t)i, pos, 0, (int)T_INSERT, (int)is_hap, alt + j, true});
        } else {
            for (j = 0; j < alt_l; j++, pos++) if (ref[j] != alt[j]) break;
            if (j == ref_l) die("Error: no deleted bases\n");
            for (; j < ref_l; j++, pos++) M.recs.push_back(MutRec{(int32_t)i, pos, 0, (int)T_DELETE, (int)is_hap, "", false});
        }
        prev_pos = pos;
    } while (ch != EOF);
}

// The substitution checks of mut_debug (src/mut.c:379-425; the reference calls it before and after mut_left_justify and
// aborts on an `assert`).  Randomly generated mutations always pass; a replayed file can name a substitution that
// does not change the base, or two different heterozygous substitutions at one position.  Here: a message and exit 1.
void check_replayed(const char *contig, const std::vector<uint8_t> &seq, const Hap &h1, const Hap &h2)
{
    const int64_t l = (int64_t)seq.size();
    for (int64_t i = 0; i < l; ++i) {
        const uint64_t c0 = g_nt4[seq[i]], c1 = h1.s[i], c2 = h2.s[i];
        if (c0 >= 4 || ((c1 & TYPE_MASK) == T_NOCHANGE && (c2 & TYPE_MASK) == T_NOCHANGE)) continue;
        bool ok = true;
        if ((c1 & BASE_TYPE_MASK) == (c2 & BASE_TYPE_MASK)) {
            if ((c1 & TYPE_MASK) == T_SUBST) ok = (c0 & 3) != (c1 & 3);
        } else if ((c1 & TYPE_MASK) == T_SUBST || (c2 & TYPE_MASK) == T_SUBST) {
            ok = (c1 & 3) != (c2 & 3) && ((c0 & 3) == (c1 & 3) || (c0 & 3) == (c2 & 3));
        }
        if (!ok) die("[dwgsim_core] Error: inconsistent substitution at %s:%lld in the mutations to replay (it must change the "
                     "base, and two alleles need separate positions); the reference aborts here in mut_debug\n", contig, (long long)i + 1);
    }
}
// mut_diref, replay branches, src/mut.c:644-745, then :752-757
void diref_replay(const Options &o, const std::vector<uint8_t> &seq, Hap &h1, Hap &h2, int contig_i, const MutsInput &M, const char *contig)
{
    const int64_t l = (int64_t)seq.size();
    h1.reset((size_t)l); h2.reset((size_t)l);
    Hap *ret[2] = {&h1, &h2};
    for (int64_t j = 0; j < l; ++j) h1.s[j] = h2.s[j] = (uint64_t)g_nt4[seq[j]];
    if (M.kind == 1) {
        for (const MutRec &r : M.recs) {
            if (r.contig == contig_i) {
                const bool has_bases = r.bases != "*";
                bool is_hom = false;
                int hap, which_hap = 0;
                if (o.is_hap || g_rng.next() < 0.333333) { is_hom = true; hap = 3; }
                else { which_hap = g_rng.next() < 0.5 ? 0 : 1; hap = 1 << which_hap; }
                if (r.type == (int)T_SUBST) {
                    for (uint32_t j = r.start; j < r.end; ++j) {
                        uint64_t c = (uint64_t)g_nt4[seq[j]];
                        if (!has_bases) { const double x = g_rng.next(); c = (c + (uint64_t)(x * 3.0 + 1)) & 3; }
                        else c = (uint64_t)g_nt4[(unsigned char)r.bases[j - r.start]];
                        if (is_hom) h1.s[j] = h2.s[j] = T_SUBST | c;
                        else ret[which_hap]->s[j] = T_SUBST | c;
                    }
                } else if (r.type == (int)T_DELETE) {
                    for (uint32_t j = r.start; j < r.end; ++j) {
                        const uint64_t c = (uint64_t)g_nt4[seq[j]];
                        if (is_hom) h1.s[j] = h2.s[j] = T_DELETE | c;
                        else ret[which_hap]->s[j] = T_DELETE | c;
                    }
                } else if (r.type == (int)T_INSERT) {
                    const uint64_t c = (uint64_t)g_nt4[seq[r.start]];
                    if (!has_bases) add_insertion(o, h1, h2, r.start, c, hap, nullptr, r.end - r.start);
                    else add_insertion(o, h1, h2, r.start, c, hap, r.bases.c_str(), 0);
                }
            } else if (contig_i < r.contig) break;
        }
    } else {
        for (const MutRec &r : M.recs) {
            if (r.contig != contig_i) continue;
            const uint32_t pos = r.start;
            const uint64_t c = (uint64_t)g_nt4[seq[pos - 1]];
            if (r.type == (int)T_DELETE) {
                if (r.is_hap & 1) h1.s[pos - 1] |= T_DELETE | c;
                if (r.is_hap & 2) h2.s[pos - 1] |= T_DELETE | c;
            } else if (r.type == (int)T_SUBST) {
                if (r.is_hap & 1) h1.s[pos - 1] = T_SUBST | g_nt4[(unsigned char)r.bases[0]];
                if (r.is_hap & 2) h2.s[pos - 1] = T_SUBST | g_nt4[(unsigned char)r.bases[0]];
            } else if (r.type == (int)T_INSERT) add_insertion(o, h1, h2, pos - 1, c, r.is_hap, r.bases.c_str(), 0);
        }
    }
    check_replayed(contig, seq, h1, h2);
    left_justify(seq, h1, h2);
    check_replayed(contig, seq, h1, h2);
}

void print_ins(FILE *fp, const Hap &h, int64_t i)                        // src/mut.c:249-279
{
    uint64_t n, ins;
    if (get_ins(h, i, &n, &ins)) { while (n > 0) { fputc(kBase[ins & 3], fp); ins >>= 2; n--; } return; }
    uint32_t num = 0;
    uint8_t *p = long_ins_payload(h.ins[ins], &num);
    int byte_i = (int)((num + 3) >> 2) - 1, bit_i = (int)((num + 3) & 3);
    while (0 < num) { fputc(kBase[(p[byte_i] >> (bit_i << 1)) & 3], fp); if (--bit_i < 0) { bit_i = 3; byte_i--; } num--; }
}
void print_del_vcf(FILE *vcf, const char *name, const std::vector<uint8_t> &seq, const Hap &h1, const Hap &h2, int64_t i, int which)
{
    const int64_t l = (int64_t)seq.size();
    uint64_t c0 = g_nt4[seq[i]], c1 = h1.s[i], c2 = h2.s[i];
    fprintf(vcf, "%s\t%lld\t.\t", name, (long long)i);
    if (0 < i) fputc(kBase[g_nt4[seq[i - 1]]], vcf);
    for (int64_t j = i; j < l; ++j) {
        const bool same = (c1 & BASE_TYPE_MASK) == (c2 & BASE_TYPE_MASK);
        const uint64_t cd = which == 2 ? c2 : c1;
        if (which == 3 ? !same : same) break;
        if ((cd & TYPE_MASK) != T_DELETE) break;
        fputc(kBase[c0], vcf);
        if (j + 1 < l) { c0 = g_nt4[seq[j + 1]]; c1 = h1.s[j + 1]; c2 = h2.s[j + 1]; }
    }
    if (0 < i) fprintf(vcf, "\t%c", kBase[g_nt4[seq[i - 1]]]); else fprintf(vcf, "\t.");
    fprintf(vcf, "\t100\tPASS\tAF=%s;pl=%d;mt=DELETE\n", which == 3 ? "1.0" : "0.5", which);
}
// mut_print, src/mut.c:781-893
// `candidates` (optional): ascending positions that contain every mutated position; the bases in between are NOCHANGE in
// both haplotypes, where the reference's loop only clears `prev`
void print_mutations(const char *name, const std::vector<uint8_t> &seq, const Hap &h1, const Hap &h2, FILE *txt, FILE *vcf,
                     const std::vector<int64_t> *candidates = nullptr)
{
    const int64_t l = (int64_t)seq.size();
    int prev[2] = {0, 0};
    const int64_t n_it = candidates ? (int64_t)candidates->size() : l;
    int64_t last = -1;
    for (int64_t it = 0; it < n_it; ++it) {
        const int64_t i = candidates ? (*candidates)[(size_t)it] : it;
        if (i != last + 1) prev[0] = prev[1] = 0;
        last = i;
        const uint64_t c0 = g_nt4[seq[i]], c1 = h1.s[i], c2 = h2.s[i], t1 = c1 & TYPE_MASK, t2 = c2 & TYPE_MASK;
        if (t1 == T_NOCHANGE && t2 == T_NOCHANGE) { prev[0] = prev[1] = 0; continue; }
        if (c0 < 4) {
            fprintf(txt, "%s\t%lld\t", name, (long long)i + 1);
            if ((c1 & BASE_TYPE_MASK) == (c2 & BASE_TYPE_MASK)) {
                if (t1 == T_SUBST) {
                    fprintf(txt, "%c\t%c\t3\n", kBase[c0], kBase[c1 & 0xf]);
                    fprintf(vcf, "%s\t%lld\t.\t%c\t%c\t100\tPASS\tAF=1.0;pl=3;mt=SUBSTITUTE\n", name, (long long)i + 1, kBase[c0], kBase[c1 & 0xf]);
                } else if (t1 == T_DELETE) {
                    fprintf(txt, "%c\t-\t3\n", kBase[c0]);
                    if (prev[0] == 0 || prev[1] == 0) print_del_vcf(vcf, name, seq, h1, h2, i, 3);
                } else {
                    fprintf(txt, "-\t"); print_ins(txt, h1, i); fprintf(txt, "\t3\n");
                    fprintf(vcf, "%s\t%lld\t.\t%c\t%c", name, (long long)i + 1, kBase[c0], kBase[c0]);
                    print_ins(vcf, h1, i);
                    fprintf(vcf, "\t100\tPASS\tAF=1.0;pl=3;mt=INSERT\n");
                }
            } else if (t1 == T_SUBST || t2 == T_SUBST) {
                const int hap = t1 == T_SUBST ? 1 : 2;
                fprintf(txt, "%c\t%c\t%d\n", kBase[c0], "XACMGRSVTWYHKDBN"[(1 << (c1 & 3)) | (1 << (c2 & 3))], hap);
                fprintf(vcf, "%s\t%lld\t.\t%c\t%c\t100\tPASS\tAF=0.5;pl=%d;mt=SUBSTITUTE\n", name, (long long)i + 1, kBase[c0],
                        kBase[(hap == 1 ? c1 : c2) & 0xf], hap);
            } else if (t1 == T_DELETE) {
                fprintf(txt, "%c\t-\t1\n", kBase[c0]);
                if (prev[0] == 0) print_del_vcf(vcf, name, seq, h1, h2, i, 1);
            } else if (t2 == T_DELETE) {
                fprintf(txt, "%c\t-\t2\n", kBase[c0]);
                if (prev[1] == 0) print_del_vcf(vcf, name, seq, h1, h2, i, 2);
            } else {
                const Hap &h = t1 == T_INSERT ? h1 : h2;
                const int hap = t1 == T_INSERT ? 1 : 2;
                fprintf(txt, "-\t"); print_ins(txt, h, i); fprintf(txt, "\t%d\n", hap);
                fprintf(vcf, "%s\t%lld\t.\t%c\t%c", name, (long long)i + 1, kBase[c0], kBase[c0]);
                print_ins(vcf, h, i);
                fprintf(vcf, "\t100\tPASS\tAF=0.5;pl=%d;mt=INSERT\n", hap);
            }
        }
        prev[0] = t1 != T_NOCHANGE;
        prev[1] = t2 != T_NOCHANGE;
    }
}

// ---- FASTQ writers: plain, or block-parallel gzip (independent gzip members, concatenated) -------------------------
struct Writer {
    FILE *fp[3] = {nullptr, nullptr, nullptr};
    bool gz = true;                 // compress here with zlib
    bool gz_file = true;            // the files are .gz (bytes arrive compressed when gz is false)
    int threads = 1;
    static constexpr size_t kBlock = 1 << 20;
    // Bytes that need no compression here are written by one helper thread per file, so the three files of a batch go out
    // side by side (copying into the page cache was the read loop's longest part).  The buffer must stay valid until the
    // write has been joined: by the next write to the same file, or by flush() -- the host shell calls it after every
    // dwgsim_gpu_run, before the library can reuse its pinned slots.
    std::future<bool> pending[3];
    int64_t offset[3] = {0, 0, 0};  // bytes handed to each file so far (positional writes)
    bool join(int id) { return pending[id].valid() ? pending[id].get() : true; }
    bool flush() { bool ok = true; for (int k = 0; k < 3; k++) ok = join(k) && ok; return ok; }
    bool write(int id, const char *buf, size_t n)
    {
        if (!fp[id] || n == 0) return true;
        if (!gz) {
            // a positional write of the batch, started here and joined by the next write to the file or by flush(): the three
            // files of a batch go out side by side and while the device works on the next batch
            if (!join(id)) return false;
            const int fd = fileno(fp[id]);
            const int64_t at = offset[id];
            offset[id] += (int64_t)n;
            pending[id] = std::async(std::launch::async, [fd, buf, n, at]() { return dwgsim_gpu_pwrite_all(fd, buf, n, at) == 0; });
            return true;
        }
        const size_t nblk = (n + kBlock - 1) / kBlock;
        std::vector<std::vector<uint8_t>> out(nblk);
        std::atomic<size_t> next{0};
        std::atomic<bool> ok{true};
        auto work = [&]() {
            for (;;) {
                const size_t b = next.fetch_add(1);
                if (b >= nblk) return;
                const size_t off = b * kBlock, len = std::min(kBlock, n - off);
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
                out[b].resize(deflateBound(&zs, (uLong)len) + 64);
                zs.next_in = (Bytef *)(buf + off); zs.avail_in = (uInt)len;
                zs.next_out = out[b].data(); zs.avail_out = (uInt)out[b].size();
                if (deflate(&zs, Z_FINISH) != Z_STREAM_END) ok = false;
                out[b].resize(zs.total_out);
                deflateEnd(&zs);
            }
        };
        const int nt = (int)std::min<size_t>((size_t)std::max(threads, 1), nblk);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
        if (!ok) return false;
        for (auto &o : out) if (fwrite(o.data(), 1, o.size(), fp[id]) != o.size()) return false;
        return true;
    }
    void close_all()
    {
        flush();
        for (auto &f : fp) if (f) {
            const int k = (int)(&f - fp);
            if (gz_file && ftell(f) == 0 && offset[k] == 0) {      // an empty gzip member, like gzclose on an untouched gzFile
                gzFile g = gzdopen(dup(fileno(f)), "wb");
                if (g) gzclose(g);
            }
            fclose(f); f = nullptr;
        }
    }
};
int sink_cb(void *user, int id, const char *buf, size_t n) { return ((Writer *)user)->write(id, buf, n) ? 0 : 1; }

FILE *xopen(const std::string &fn, const char *mode)
{
    FILE *fp = fopen(fn.c_str(), mode);
    if (!fp) { fprintf(stderr, "[dwgsim] fail to open file '%s'. Abort!\n", fn.c_str()); exit(1); }
    return fp;
}

// ---- -x regions (src/regions_bed.c) ------------------------------------------------------------------------------------
struct Regions {
    std::vector<uint32_t> contig, start, end;
    // regions_bed_init, src/regions_bed.c:43-115: contigs in FASTA order, starts sorted, overlapping regions merged
    void read(const std::string &path, const ContigList &c)
    {
        FILE *fp = xopen(path, "r");
        char name[1024];
        uint32_t s, e;
        size_t i = 0;
        long prev_contig = -1, prev_start = -1, prev_end = -1;
        while (0 < fscanf(fp, "%1023s\t%u\t%u", name, &s, &e)) {
            while (i < c.name.size() && c.name[i] != name) i++;
            if (i == c.name.size()) { fprintf(stderr, "Error: contig not found: %s.  Are you sure your BED is coordinate sorted?\n", name); exit(1); }
            if (c.len[i] < s || c.len[i] < e) { fprintf(stderr, "Error: start/end was out of range\n"); exit(1); }
            if (e < s) { fprintf(stderr, "Error: end < start: [%s,%u,%u]\n", name, s, e); exit(1); }
            if (prev_contig == (long)i && (long)s < prev_start) { fprintf(stderr, "Error: the input was not sorted: [%s,%u,%u]\n", name, s, e); exit(1); }
            if (prev_contig == (long)i && (long)s <= prev_end && prev_start <= (long)s) {
                if (prev_end < (long)e) { end.back() = e; prev_end = e; }
            } else {
                prev_contig = (long)i; prev_start = s; prev_end = e;
                contig.push_back((uint32_t)i); start.push_back(s); end.push_back(e);
            }
            int b;
            while (EOF != (b = fgetc(fp))) if ('\n' == b || '\r' == b) break;
        }
        fclose(fp);
    }
};

}  // namespace

int main(int argc, char **argv)
{
    nt4_init();
    Options o;
    int first = 0;
    if (!parse_options(o, argc, argv, &first)) return usage(o);
    const std::string fn_fa = argv[first], prefix = argv[first + 1];
    Fasta fa;
    if (!fa.open(fn_fa.c_str())) { fprintf(stderr, "[main] fail to open file '%s'. Abort!\n", fn_fa.c_str()); return 1; }
    FILE *fp_fai = fopen((fn_fa + ".fai").c_str(), "r");
    FILE *fp_txt = nullptr, *fp_vcf = nullptr;
    if (o.output_type != 1) { fp_txt = xopen(prefix + ".mutations.txt", "w"); fp_vcf = xopen(prefix + ".mutations.vcf", "w"); }
    Writer wr;
    wr.gz = !o.uncompressed && o.host_gzip;          // zlib on the host only when asked; default: gzip members from the GPU
    wr.gz_file = !o.uncompressed;
    wr.threads = o.threads > 0 ? o.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const char *ext = o.uncompressed ? "" : ".gz";
    dwgsim_gpu_t *gpu = nullptr;
    if (o.output_type != 2) {
        if (o.reads_output_type != 1) wr.fp[2] = xopen(prefix + ".bfast.fastq" + ext, "wb");
        if (o.reads_output_type != 2) { wr.fp[0] = xopen(prefix + ".bwa.read1.fastq" + ext, "wb"); wr.fp[1] = xopen(prefix + ".bwa.read2.fastq" + ext, "wb"); }
    }
    // the device handle is created when the first contig with pairs arrives (a -C 0 run never needs a GPU)
    auto gpu_open = [&]() {
        dwgsim_gpu_params_t p;
        memset(&p, 0, sizeof p);
        for (int i = 0; i < 2; i++) { p.e_start[i] = o.e_start[i]; p.e_by[i] = o.e_by[i]; p.length[i] = o.length[i]; }
        p.is_inner = o.is_inner; p.dist = o.dist; p.std_dev = o.std_dev; p.mut_freq = o.mut_freq; p.rand_read = o.rand_read;
        p.max_n = o.max_n; p.data_type = o.data_type; p.strandedness = o.strandedness; p.read_one_strand = o.read_one_strand;
        p.flow_order = o.flow_codes.empty() ? nullptr : o.flow_codes.data(); p.flow_order_len = (int32_t)o.flow_codes.size();
        p.seed = o.seed == -1 ? (int32_t)time(nullptr) : o.seed;
        p.fixed_quality = o.has_fixed_quality ? (unsigned char)o.fixed_quality[0] : 0;
        p.quality_std = o.quality_std; p.read_prefix = o.has_prefix ? o.read_prefix.c_str() : nullptr;
        p.reads_output_type = o.reads_output_type; p.amplicons = o.amplicons;
        int want = o.gpus;
        if (want == 0) if (const char *e = getenv("DWGSIM_GPUS")) want = atoi(e);
        if (want == 0) want = 1;
        std::vector<int32_t> devs;
        if (want < 0) want = 64;                                         // all there are
        for (int d = o.device; (int)devs.size() < want; ++d) devs.push_back(d);
        if (const char *e = getenv("DWGSIM_DEVICES")) {                  // explicit list "0,1,3" (ids may repeat)
            devs.clear();
            for (const char *q = e; *q;) { devs.push_back((int32_t)strtol(q, (char **)&q, 10)); while (*q == ',' || *q == ' ') ++q; }
            if (devs.empty()) devs.push_back(o.device);
        }
        int rc = DWGSIM_GPU_ENODEV;
        // (-1 / more than the box has: shrink to the devices that exist)
        for (; !devs.empty(); devs.pop_back()) {
            rc = devs.size() == 1 ? dwgsim_gpu_create(&gpu, &p, devs[0]) : dwgsim_gpu_create_group(&gpu, &p, devs.data(), (int32_t)devs.size());
            if (rc != DWGSIM_GPU_ENODEV || (o.gpus > 0 || devs.size() == 1)) break;
        }
        if (rc != DWGSIM_GPU_OK) { fprintf(stderr, "\n[dwgsim_core] Error: %s\n", dwgsim_gpu_strerror(rc)); exit(1); }
        // 2^18 pairs per device batch (the per-batch launch costs are small against it); with four or more devices half of
        // that, so that page-locking the devices' rings stays a matter of seconds
        dwgsim_gpu_set_batch(gpu, o.batch > 0 ? o.batch : (dwgsim_gpu_group_size(gpu) >= 4 ? (1 << 17) : (1 << 18)), 3);
        if (!o.uncompressed && !o.host_gzip) dwgsim_gpu_set_compression(gpu, 1);
    };

    // With reads to simulate, the device handle, its batch workspace and the pinned ring are set up on a helper thread while
    // the census and the first contig's prologue run (CUDA start-up and page-locking the ring take about two seconds)
    std::future<void> gpu_ready;
    if (o.output_type != 2 && (o.N > 0 || o.C > 0) && !getenv("DWGSIM_LAZY_GPU"))
        gpu_ready = std::async(std::launch::async, [&]() {
            gpu_open();
            const int rc = dwgsim_gpu_warm(gpu);
            if (rc != DWGSIM_GPU_OK) { fprintf(stderr, "\n[dwgsim_core] Error: %s: %s\n", dwgsim_gpu_strerror(rc), dwgsim_gpu_last_error(gpu)); exit(1); }
        });

    // census, src/dwgsim.c:465-492
    std::vector<uint8_t> seq;
    std::string name;
    uint64_t tot_len = 0;
    int n_ref = 0;
    ContigList contigs;
    if (!o.fn_regions_bed.empty()) fclose(xopen(o.fn_regions_bed, "r"));   // fail before the census, like src/dwgsim.c:460
    if (fp_vcf) fprintf(fp_vcf, "##fileformat=VCFv4.1\n");
    if (fp_fai) {
        char nm[1024];
        int l, d0, d1, d2;
        while (0 < fscanf(fp_fai, "%1023s\t%d\t%d\t%d\t%d", nm, &l, &d0, &d1, &d2)) {
            fprintf(stderr, "[dwgsim_core] %s length: %d\n", nm, l);
            tot_len += (uint64_t)l; ++n_ref;
            contigs.name.push_back(nm); contigs.len.push_back((uint32_t)l);
            if (fp_vcf) fprintf(fp_vcf, "##contig=<ID=%s,length=%d>\n", nm, l);
        }
        fclose(fp_fai);
    } else {
        int64_t l;
        while ((l = fa.next(seq, name, true)) >= 0) {
            fprintf(stderr, "[dwgsim_core] %s length: %lld\n", name.c_str(), (long long)l);
            tot_len += (uint64_t)l; ++n_ref;
            contigs.name.push_back(name); contigs.len.push_back((uint32_t)l);
            if (fp_vcf) fprintf(fp_vcf, "##contig=<ID=%s,length=%d>\n", name.c_str(), (int)l);
        }
    }
    fprintf(stderr, "[dwgsim_core] %d sequences, total length: %llu\n", n_ref, (unsigned long long)tot_len);
    fa.pos = 0;
    if (fp_vcf) {                                                        // src/mut.c:765-771
        fprintf(fp_vcf, "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency\">\n");
        fprintf(fp_vcf, "##INFO=<ID=pl,Number=1,Type=Integer,Description=\"Phasing: 1 - HET contig 1, #2 - HET contig #2, 3 - HOM both contigs\">\n");
        fprintf(fp_vcf, "##INFO=<ID=mt,Number=1,Type=String,Description=\"Variant Type: SUBSTITUTE/INSERT/DELETE\">\n");
        fprintf(fp_vcf, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n");
    }
    MutsInput muts;
    if (o.muts_input_type >= 0) {                                        // src/dwgsim.c:494-497
        FILE *fp = xopen(o.fn_muts_input, "r");
        muts.kind = o.muts_input_type;
        if (muts.kind == 0) read_muts_txt(fp, contigs, muts);
        else if (muts.kind == 1) read_muts_bed(fp, contigs, muts);
        else read_muts_vcf(fp, contigs, muts);
        fclose(fp);
        // the reference shrinks its record array to the number of records read; with none, realloc(p, 0) returns NULL
        // and it exits with this message (src/mut_txt.c:117-125, src/mut_bed.c:126-134, src/mut_vcf.c:267-275)
        if (muts.recs.empty()) die("Error: memory allocation failed in muts_%s_init\n", muts.kind == 0 ? "txt" : (muts.kind == 1 ? "bed" : "vcf"));
    }
    Regions regions;
    const bool use_regions = !o.fn_regions_bed.empty();
    if (use_regions) {                                                   // src/dwgsim.c:499-506
        regions.read(o.fn_regions_bed, contigs);
        tot_len = 0;
        for (size_t i = 0; i < regions.start.size(); i++) tot_len += regions.end[i] - regions.start[i];
    }
    fprintf(stderr, o.output_type != 2 ? "[dwgsim_core] Currently on: \n0" : "[dwgsim_core] Currently on:");

    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mut = 0, t_print = 0, t_gpu = 0, t_pack = 0, t_kernels = 0;
    const double t_begin = now();
    long long bytes_out = 0, bases_in = 0;
    long long n_sim = 0;
    unsigned long long ctr = 0;
    int rc_exit = 0;
    const int maxlen = std::max(o.length[0], o.length[1]);

    // One contig's prologue (src/dwgsim.c:519-632): budget, skip rules, mut_diref.  With reads to simulate it runs on a
    // producer thread, one contig ahead of the consumer below (mutation files, packing, GPU read loop), so the host's
    // serial part -- the drand48 stream of mut_diref must be replayed in contig order -- overlaps the rest.
    struct Job {
        std::string name;
        std::vector<uint8_t> seq;
        Hap h1, h2;
        int seq_l = 0, l = 0, contig_i = 0;
        long long n_pairs = 0;
        std::vector<uint32_t> reg_start, reg_end;
        std::vector<int64_t> events, touched;     // mutated positions (random mut_diref only): the writers need not scan the contig
        bool have_events = false;
        double t_mut = 0;
    };
    struct Producer {
        int contig_i = 0, prev_skip = 0, n_ref = 0;
        long long n_sim = 0;                      // pairs budgeted so far (what the consumer's n_sim will be)
    } P;
    P.n_ref = n_ref;
    std::mutex pool_mu;
    std::vector<std::unique_ptr<Job>> pool;         // finished jobs: their vectors keep capacity (and mapped pages) for the next contig
    auto recycle = [&](std::unique_ptr<Job> j) { std::lock_guard<std::mutex> g(pool_mu); pool.push_back(std::move(j)); };
    // mut_diref of long contigs runs on several threads (diref_parallel); DWGSIM_DIREF_THREADS / DWGSIM_DIREF_PAR_MIN override
    int diref_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / 2));
    if (const char *e = getenv("DWGSIM_DIREF_THREADS")) diref_threads = std::max(1, atoi(e));
    int64_t diref_par_min = 1 << 20;
    if (const char *e = getenv("DWGSIM_DIREF_PAR_MIN")) diref_par_min = atoll(e);
    auto produce = [&]() -> std::unique_ptr<Job> {   // next simulated contig, or null at the end of the FASTA
        std::unique_ptr<Job> j;
        {
            std::lock_guard<std::mutex> g(pool_mu);
            if (!pool.empty()) { j = std::move(pool.back()); pool.pop_back(); }
        }
        if (!j) j.reset(new Job);
        int64_t l64;
        while ((l64 = fa.next(j->seq, j->name)) >= 0) {                   // src/dwgsim.c:519-1106
            const std::string &name = j->name;
            const std::vector<uint8_t> &seq = j->seq;
            const int seq_l = (int)l64;
            int l = seq_l;                                                // with -x: the total length of the contig's regions
            long long n_pairs = 0;
            std::vector<uint32_t> &reg_start = j->reg_start, &reg_end = j->reg_end;
            reg_start.clear(); reg_end.clear();
            const int contig_i = P.contig_i;
            P.n_ref--;
            if (o.output_type == 2) fprintf(stderr, "\r[dwgsim_core] Currently on: %s", name.c_str());
            else {
                if (use_regions)
                    for (size_t i = 0; i < regions.start.size(); i++)
                        if ((uint32_t)contig_i == regions.contig[i]) { reg_start.push_back(regions.start[i]); reg_end.push_back(regions.end[i]); }
                if (0 == P.n_ref && o.C < 0) n_pairs = o.N - P.n_sim;     // NB: the last contig keeps its full length, also with -x
                else if (use_regions && [&]() {                           // src/dwgsim.c:539-581
                         int m = 0, num_n = 0;
                         for (size_t i = 0; i < reg_start.size(); i++) m += (int)(reg_end[i] - reg_start[i]);
                         if (0 == m) { fprintf(stderr, "[dwgsim_core] #0 skip sequence '%s' as it is not in the targeted region\n", name.c_str()); return true; }
                         l = m;
                         for (size_t i = 0; i < reg_start.size(); i++)
                             for (uint32_t q = reg_start[i]; q <= reg_end[i]; q++) {     // the reference reads seq[q-1] for q in [start, end]
                                 const int ch = q >= 1 ? seq[q - 1] : 'N';
                                 switch (ch) { case 'a': case 'A': case 'c': case 'C': case 'g': case 'G': case 't': case 'T': break; default: num_n++; }
                             }
                         if (0.95 < num_n / (double)l) { fprintf(stderr, "[dwgsim_core] #1 skip sequence '%s' as %d out of %d bases are non-ACGT\n", name.c_str(), num_n, l); return true; }
                         return false;
                     }()) { P.contig_i++; continue; }
                else if (0 < o.N) {
                    n_pairs = (long long)(uint64_t)((long double)l / tot_len * o.N + 0.5);
                    if (o.N - P.n_sim < n_pairs) n_pairs = o.N - P.n_sim;
                } else n_pairs = (long long)(uint64_t)(l * o.C / ((long double)(o.length[0] + o.length[1])) / (1.0 - o.rand_read) + 0.5);
                auto skip = [&]() { if (0 == P.prev_skip) fprintf(stderr, "\n"); P.prev_skip = 1; };
                if (o.amplicons == 1) {
                    if (l < maxlen) { skip(); fprintf(stderr, "[dwgsim_core] #2 skip sequence '%s' as it is shorter than the read length %d < %d!\n", name.c_str(), l, maxlen); P.contig_i++; continue; }
                } else if (0 < o.length[1] && l < o.dist + 3 * o.std_dev) {
                    skip(); fprintf(stderr, "[dwgsim_core] #3 skip sequence '%s' as it is shorter than %f!\n", name.c_str(), o.dist + 3 * o.std_dev); P.contig_i++; continue;
                } else if (l < o.length[0] || (0 < o.length[1] && l < o.length[1])) {
                    skip(); fprintf(stderr, "[dwgsim_core] #4 skip sequence '%s' as it is shorter than %d!\n", name.c_str(), l < o.length[0] ? o.length[0] : o.length[1]); P.contig_i++; continue;
                } else if (n_pairs < 0) { fprintf(stderr, "[dwgsim_core] #5 skip sequence '%s' as not enough pairs found\n", name.c_str()); continue; }
                P.prev_skip = 0;
            }
            j->seq_l = seq_l; j->l = l; j->contig_i = contig_i; j->n_pairs = n_pairs;
            if (o.output_type != 2 && n_pairs > 0) P.n_sim += n_pairs;
            P.contig_i++;
            return j;
        }
        return nullptr;
    };
    // mut_diref of one contig (the drand48 stream: strictly in contig order, on one thread at a time)
    auto mutate = [&](Job &jb) {
        Job *j = &jb;
        const double t0 = now();
        j->have_events = muts.kind < 0 && !getenv("DWGSIM_FULL_SCAN");
        if (muts.kind >= 0) diref_replay(o, j->seq, j->h1, j->h2, j->contig_i, muts, j->name.c_str());
        else if (j->have_events && diref_threads > 1 && (int64_t)j->seq.size() >= diref_par_min)
            diref_parallel(o, j->seq, j->h1, j->h2, &j->events, &j->touched, diref_threads);
        else if (j->have_events) diref(o, j->seq, j->h1, j->h2, &j->events, &j->touched);
        else diref(o, j->seq, j->h1, j->h2);
        j->t_mut = now() - t0;
    };
    // The consumer side in two halves so they can run on two threads: prepare() writes the contig's mutation records and packs
    // it for the device (host-only work on the dense arrays, which go back to the pool right after); execute() queues the
    // packed contig, runs the read loop on the device(s) and writes the FASTQ files.
    struct RunJob {
        std::string name;
        int seq_l = 0, l = 0, contig_i = 0;
        long long n_pairs = 0;
        std::vector<uint32_t> reg_start, reg_end;
        dwgsim_gpu_packed_t *packed = nullptr;
        double t_mut = 0, t_print = 0, t_pack = 0;
        bool failed = false;
    };
    std::mutex gpu_mu;
    auto prepare = [&](Job &j) -> std::unique_ptr<RunJob> {
        std::unique_ptr<RunJob> r(new RunJob);
        r->name = j.name; r->seq_l = j.seq_l; r->l = j.l; r->contig_i = j.contig_i; r->n_pairs = j.n_pairs; r->t_mut = j.t_mut;
        r->reg_start = j.reg_start; r->reg_end = j.reg_end;
        double t0 = now();
        if (o.output_type != 1) print_mutations(j.name.c_str(), j.seq, j.h1, j.h2, fp_txt, fp_vcf, j.have_events ? &j.events : nullptr);
        r->t_print = now() - t0;
        if (o.output_type != 2 && j.n_pairs > 0) {
            { std::lock_guard<std::mutex> g(gpu_mu); if (gpu_ready.valid()) gpu_ready.get(); if (!gpu) gpu_open(); }
            t0 = now();
            const int rc = dwgsim_gpu_pack_contig(gpu, j.contig_i, j.name.c_str(), j.seq.data(), j.seq_l, j.h1.s.data(), j.h2.s.data(), j.h1.ins.data(),
                                                  (int32_t)j.h1.ins.size(), j.h2.ins.data(), (int32_t)j.h2.ins.size(), j.n_pairs, &r->packed);
            r->t_pack = now() - t0;
            if (rc != DWGSIM_GPU_OK) {
                fprintf(stderr, "\r[dwgsim_core] %s: packing '%s' for the device failed\n", dwgsim_gpu_strerror(rc), j.name.c_str());
                r->failed = true;
            }
        }
        return r;
    };
    auto execute = [&](RunJob &j) -> bool {                              // false: stop (the GPU path reported an error)
        t_mut += j.t_mut; t_print += j.t_print;
        bases_in += j.seq_l;
        if (j.failed) { rc_exit = 1; return false; }
        if (j.packed) {
            const double t0 = now();
            int rc = dwgsim_gpu_add_packed(gpu, j.packed);
            j.packed = nullptr;
            if (rc == DWGSIM_GPU_OK && use_regions) rc = dwgsim_gpu_set_regions(gpu, j.reg_start.data(), j.reg_end.data(), (int32_t)j.reg_start.size(), j.l);
            dwgsim_gpu_stats_t st;
            memset(&st, 0, sizeof st);
            if (rc == DWGSIM_GPU_OK) rc = dwgsim_gpu_run(gpu, sink_cb, &wr, &st);
            if (!wr.flush() && rc == DWGSIM_GPU_OK) { fprintf(stderr, "\r[dwgsim_core] Error: writing the FASTQ files failed\n"); rc_exit = 1; return false; }
            if (rc != DWGSIM_GPU_OK) {
                fprintf(stderr, "\r[dwgsim_core] %s%s%s\n", dwgsim_gpu_strerror(rc), *dwgsim_gpu_last_error(gpu) ? ": " : "", dwgsim_gpu_last_error(gpu));
                rc_exit = 1;
                return false;
            }
            t_gpu += now() - t0; t_pack += j.t_pack; t_kernels += (st.ms_simulate + st.ms_layout + st.ms_format) * 1e-3;
            bytes_out += st.bytes[0] + st.bytes[1] + st.bytes[2];
            ctr += (unsigned long long)j.n_pairs; n_sim += j.n_pairs;
            fprintf(stderr, "\r[dwgsim_core] %llu", ctr);
        }
        return true;
    };
    // The prologue is memory-bound (17 B/base of dense arrays): more threads only pay when there is GPU work to hide
    // behind it.  DWGSIM_PIPELINE=0/1 overrides.
    const char *pipe_env = getenv("DWGSIM_PIPELINE");
    const bool pipelined = pipe_env ? atoi(pipe_env) != 0 : (o.output_type != 2 && (o.N > 0 || o.C > 0));
    if (!pipelined) {
        for (;;) {
            std::unique_ptr<Job> j = produce();
            if (!j) break;
            mutate(*j);
            std::unique_ptr<RunJob> r = prepare(*j);
            recycle(std::move(j));
            if (!execute(*r)) break;
        }
    } else {
        // Four stages, one contig each: FASTA record + budget / skip rules (reader thread) -> mut_diref (mutator thread; the
        // drand48 stream stays in contig order) -> mutation files + packing (packer thread) -> device read loop + FASTQ files
        // (this thread).  Bounded hand-over: one finished item waits per stage, so at most four contigs hold dense arrays.
        struct Slot {
            std::mutex mu;
            std::condition_variable cv;
            bool full = false, done = false, stop = false;
        } sr, sa, sb;
        std::unique_ptr<Job> slot_r, slot_a;
        std::unique_ptr<RunJob> slot_b;
        auto put = [](Slot &sl, auto &slot, auto item) -> bool {      // false: the pipeline was stopped
            std::unique_lock<std::mutex> lk(sl.mu);
            sl.cv.wait(lk, [&]() { return !sl.full || sl.stop; });
            if (sl.stop) return false;
            slot = std::move(item); sl.full = true;
            sl.cv.notify_all();
            return true;
        };
        auto finish = [](Slot &sl) { std::unique_lock<std::mutex> lk(sl.mu); sl.done = true; sl.cv.notify_all(); };
        auto take = [](Slot &sl, auto &slot, auto &item) -> int {     // 1: item taken, 0: upstream finished, -1: stopped
            std::unique_lock<std::mutex> lk(sl.mu);
            sl.cv.wait(lk, [&]() { return sl.full || sl.done || sl.stop; });
            if (sl.stop) return -1;
            if (!sl.full) return 0;
            item = std::move(slot); sl.full = false;
            sl.cv.notify_all();
            return 1;
        };
        auto halt = [](Slot &sl) { std::unique_lock<std::mutex> lk(sl.mu); sl.stop = true; sl.cv.notify_all(); };
        std::thread reader([&]() {
            for (;;) {
                std::unique_ptr<Job> j = produce();
                if (!j) { finish(sr); return; }
                if (!put(sr, slot_r, std::move(j))) return;
            }
        });
        std::thread mutator([&]() {
            for (;;) {
                std::unique_ptr<Job> j;
                const int got = take(sr, slot_r, j);
                if (got < 0) return;
                if (got == 0) { finish(sa); return; }
                mutate(*j);
                if (!put(sa, slot_a, std::move(j))) return;
            }
        });
        std::thread packer([&]() {
            for (;;) {
                std::unique_ptr<Job> j;
                const int got = take(sa, slot_a, j);
                if (got < 0) return;
                if (got == 0) { finish(sb); return; }
                std::unique_ptr<RunJob> r = prepare(*j);
                recycle(std::move(j));
                dwgsim_gpu_packed_t *pk = r->packed;
                if (!put(sb, slot_b, std::move(r))) { if (pk) dwgsim_gpu_packed_free(pk); return; }
            }
        });
        for (;;) {
            std::unique_ptr<RunJob> r;
            const int got = take(sb, slot_b, r);
            if (got <= 0) break;
            if (!execute(*r)) { halt(sb); halt(sa); halt(sr); break; }
        }
        packer.join();
        mutator.join();
        reader.join();
        if (slot_b && slot_b->packed) dwgsim_gpu_packed_free(slot_b->packed);
    }
    if (!rc_exit) fprintf(stderr, "\n[dwgsim_core] Complete!\n");
    if (getenv("DWGSIM_STATS")) {
        const double total = now() - t_begin;
        fprintf(stderr, "[dwgsim_b200] bases %lld pairs %lld fastq_bytes %lld | total %.3f s: mut_diref %.3f s (%.1f ns/base), mut_print %.3f s, "
                        "read loop %.3f s (host pack %.3f s, kernels %.3f s, rest = D2H + sink%s) | %.3f Mpairs/s overall, %.3f Mpairs/s in the read loop\n",
                bases_in, n_sim, bytes_out, total, t_mut, bases_in ? 1e9 * t_mut / bases_in : 0.0, t_print, t_gpu, t_pack, t_kernels,
                o.uncompressed ? "" : " incl. gzip", total > 0 ? n_sim / total / 1e6 : 0.0, t_gpu > 0 ? n_sim / t_gpu / 1e6 : 0.0);
    }
    if (gpu_ready.valid()) gpu_ready.get();
    if (fp_txt) fclose(fp_txt);
    if (fp_vcf) fclose(fp_vcf);
    wr.close_all();
    // every file is closed: leave without tearing the CUDA contexts and the pinned rings down page by page (seconds with
    // several devices); DWGSIM_CLEAN_EXIT=1 keeps the orderly shutdown (leak checkers)
    if (gpu && !getenv("DWGSIM_CLEAN_EXIT")) { fflush(stdout); fflush(stderr); _exit(rc_exit); }
    if (gpu) dwgsim_gpu_destroy(gpu);
    return rc_exit;
}
