/*
 * dwgsim_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the nh13/DWGSIM read-pair path (dwgsim_core, src/dwgsim.c:419-1121 of the
 * reference) and of the host-side producers that feed it (seq_read_fasta, mut_diref,
 * mut_left_justify, mut_print; src/mut.c).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (dwgsim_b200/, include/dwgsim_gpu.h) never links, imports or executes anything under oracle/.
 *
 * One driver, two random-number backends:
 *   ORC_RNG_DRAND48  every draw comes from one global glibc-compatible drand48 stream in the exact
 *                    call order of the reference.  In this mode the oracle is byte-identical to
 *                    the reference binary (pinned against the reference's five golden files,
 *                    testdata/ there, and against oracle/_ref/dwgsim_ref on every platform mode).
 *   ORC_RNG_PHILOX   the read loop draws from Philox4x32-10 addressed by
 *                    (seed, global pair index, attempt, stream, index) -- the specification the
 *                    CUDA kernels implement bit-for-bit (DESIGN.md "RNG addressing").  Mutation
 *                    generation still uses drand48, so .mutations.* equal the reference run with
 *                    -C 0.
 * Everything that is not a random draw (read placement, the walk through the mutation arrays,
 * N filtering, colour encoding, flow-space error bookkeeping, Phred arithmetic, record grammar)
 * is shared by both backends, so byte parity of the first backend with the reference carries
 * over to the second.
 */
#ifndef DWGSIM_ORACLE_H
#define DWGSIM_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RNG_DRAND48 = 0, ORC_RNG_PHILOX = 1 };
enum { ORC_ILLUMINA = 0, ORC_SOLID = 1, ORC_IONTORRENT = 2 };

/* mirrors dwgsim_opt_t (src/dwgsim_opt.h:21-60) minus file handles and -m/-b/-v/-x inputs */
typedef struct {
    double  e_start[2], e_end[2], e_by[2];  /* e_by is filled by orc_opt_finalize */
    int32_t is_inner;
    int32_t dist;
    double  std_dev;
    int64_t N;
    double  C;
    int32_t length[2];
    double  mut_rate, mut_freq, indel_frac, indel_extend;
    int32_t indel_min;
    double  rand_read;
    int32_t max_n;
    int32_t data_type;
    int32_t strandedness;
    int32_t read_one_strand;
    int32_t flow_order_len;
    int8_t  flow_order[1024];               /* ASCII on input; codes 0..3 after orc_opt_finalize */
    int32_t use_base_error;
    int32_t is_hap;
    int32_t seed;
    int32_t fixed_quality;                  /* 0 = none, else the character */
    double  quality_std;
    int32_t has_read_prefix;
    char    read_prefix[256];
    int32_t reads_output_type;              /* 0 all, 1 bwa, 2 bfast */
    int32_t output_type;                    /* 0 all, 1 reads, 2 mutations */
    int32_t amplicons;
    int32_t finalized;
    char    fn_regions_bed[1024];           /* -x: BED of regions to cover ("" = none)              */
    int32_t muts_input_type;                /* -1 none, 0 = -b BED, 1 = -m TXT, 2 = -v VCF (src/mut_input.h:33-37) */
    char    fn_muts_input[1024];            /* the file of mutations to replay                      */
} orc_opt_t;

/* tables the Philox backend (and the GPU) samples from instead of calling log/sqrt per draw */
typedef struct {
    uint64_t thr_genomic;       /* genomic pair iff (uint64)u32 >= thr_genomic                    */
    uint64_t thr_hap0;          /* haplotype 0 iff (uint64)u32 < thr_hap0                         */
    int32_t  isize_lo;          /* insert size = isize_lo + #{j : u32 >= isize_cdf[j]}            */
    int32_t  isize_n;           /* number of thresholds                                           */
    uint32_t *isize_cdf;
    int32_t  qdelta_lo;         /* quality noise = qdelta_lo + #{j : u32 >= qdelta_cdf[j]}        */
    int32_t  qdelta_n;
    uint32_t *qdelta_cdf;
    int32_t  n_cycles[2];       /* table lengths (read length, or grown for Ion Torrent)          */
    uint32_t *err_gap[2];       /* candidate gap = #{g : u32 >= err_gap[end][g]}, g < read length  */
    uint32_t *err_acc[2];       /* candidate at cycle i becomes an error iff u32 < err_acc[end][i] */
    uint8_t  *qbase[2];         /* Phred before noise, 0..40, per cycle                           */
    uint32_t flow_thr[2];       /* Ion Torrent: per-flow error probability of each end as a 32-bit threshold */
    uint32_t *flow_gap[2];      /* Ion Torrent: failures before the next success of the per-flow coin =          */
                                /* #{g < ORC_FLOW_GAP_N : u32 >= flow_gap[end][g]}; ORC_FLOW_GAP_N = "no success yet" */
} orc_tables_t;
#define ORC_FLOW_GAP_N 4096

typedef struct {
    int64_t n_pairs_total;      /* pairs written (genomic + random)                               */
    int64_t n_random;           /* random pairs                                                   */
    int64_t n_failed_attempts;  /* rejected genomic attempts                                      */
    int64_t n_contigs, n_contigs_skipped;
    int64_t bytes_bwa1, bytes_bwa2, bytes_bfast;
    int32_t error;              /* 0 ok; 1 = 10001-failure abort; 2 = io; 4 = Ion read overflow   */
} orc_stats_t;

typedef struct orc_session orc_session_t;

void orc_opt_init(orc_opt_t *opt);                       /* defaults: src/dwgsim_opt.c:40-80      */
/* range checks + drand48 seeding + Ion flow order + -B calibration + slope:
 * src/dwgsim_opt.c:307-469.  Returns 1 if ok, 0 if the reference would print usage and stop. */
int  orc_opt_finalize(orc_opt_t *opt);
int  orc_parse_error_rate(const char *str, double *start, double *end); /* src/dwgsim_opt.c:162-179 */

/* Run the whole driver over a FASTA file.  out_prefix may be NULL (no files written).  Files are
 * written UNCOMPRESSED as <prefix>.bwa.read1.fastq, .bwa.read2.fastq, .bfast.fastq,
 * .mutations.txt, .mutations.vcf (the reference gzips the FASTQs; goldens are compared
 * after gunzip).  keep_contigs != 0 retains every simulated contig's sequence and mutation arrays
 * for the accessors below (this is what a reference host would hand to the C-ABI at the seam). */
orc_session_t *orc_run(const orc_opt_t *opt, const char *fasta_path, const char *out_prefix,
                       int rng_mode, int keep_contigs);
void orc_close(orc_session_t *s);
const orc_stats_t *orc_stats(const orc_session_t *s);

int32_t      orc_n_contigs(const orc_session_t *s);      /* kept (non-skipped) contigs            */
const char  *orc_contig_name(const orc_session_t *s, int32_t k);
int32_t      orc_contig_index(const orc_session_t *s, int32_t k);   /* contig_i in the FASTA      */
int32_t      orc_contig_len(const orc_session_t *s, int32_t k);
int64_t      orc_contig_n_pairs(const orc_session_t *s, int32_t k);
const uint8_t  *orc_contig_seq(const orc_session_t *s, int32_t k);  /* ASCII                      */
const uint64_t *orc_contig_hap(const orc_session_t *s, int32_t k, int32_t hap); /* mut_t[len]     */
int32_t      orc_contig_n_ins(const orc_session_t *s, int32_t k, int32_t hap);
/* -x regions of the kept contig (merged, sorted): returns their number, fills the two pointers */
/* the length the position sampler uses: the contig's, or with -x the total length of its regions */
int32_t      orc_contig_sample_len(const orc_session_t *s, int32_t k);
int32_t      orc_contig_regions(const orc_session_t *s, int32_t k, const uint32_t **start, const uint32_t **end);
uint8_t *const *orc_contig_ins(const orc_session_t *s, int32_t k, int32_t hap);

/* derived tables (Philox backend) */
orc_tables_t *orc_tables_build(const orc_opt_t *opt);
void          orc_tables_free(orc_tables_t *t);

/* primitives exposed for known-answer tests */
void     orc_srand48(int32_t seed);                      /* src/dwgsim_opt.c:387-394              */
double   orc_drand48(void);
uint64_t orc_drand48_state(void);
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u32 draw #idx of (stream,end) for pair gidx / attempt: the addressing the kernels use */
uint32_t orc_philox_draw(int32_t seed, uint64_t gidx, uint32_t attempt, uint32_t stream,
                         uint32_t end, uint32_t idx);
/* Ion Torrent flow model on one read (src/dwgsim.c:246-417) driven by the drand48 stream */
int32_t  orc_generate_errors_flows(const orc_opt_t *opt, uint8_t *seq, int32_t cap, uint8_t *mask,
                                   int32_t len, int32_t strand, double e, int32_t *n_err);

#ifdef __cplusplus
}
#endif
#endif
