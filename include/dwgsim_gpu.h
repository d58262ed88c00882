/*
 * dwgsim_gpu.h -- C ABI of libdwgsim_b200.so: the dwgsim_core read-pair loop on a B200.
 *
 * The reference has no plugin / FFI surface for this path; the loop is inlined in
 * dwgsim_core (reference src/dwgsim.c:636-1099).  The seam is cut where a maintainer would cut it:
 * once per contig, right after mut_diref + mut_print (src/dwgsim.c:628-632) and before
 * mutseq_destroy (src/dwgsim.c:1103-1104).  Everything crossing the boundary is plain C: POD
 * structs, pointers and sizes.  No torch / CUDA types appear in any signature.
 *
 * Error convention: the reference prints to stderr and exit(1)s (src/dwgsim.c:177-180,837-840);
 * this library never exits: every call returns DWGSIM_GPU_OK (0) or a negative code, and
 * dwgsim_gpu_strerror() gives the text the host shell prints before exiting 1.
 *
 * Threading: callable from the reference's single thread.  The library owns its CUDA streams,
 * pinned rings and device memory, and never touches libc's drand48 state (the host's mut_diref
 * depends on it).
 */
#ifndef DWGSIM_GPU_H
#define DWGSIM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWGSIM_GPU_ABI_VERSION 2

enum {
    DWGSIM_GPU_OK = 0,
    DWGSIM_GPU_EINVAL = -1,        /* bad argument / option out of the range the reference accepts   */
    DWGSIM_GPU_ENODEV = -2,        /* no usable CUDA device: there is NO CPU fallback                  */
    DWGSIM_GPU_ECUDA = -3,         /* CUDA runtime error (text via dwgsim_gpu_last_error)              */
    DWGSIM_GPU_ENOMEM = -4,
    DWGSIM_GPU_ETRIALS = -5,       /* "failed to generate a read after 10001 trials" src/dwgsim.c:837 */
    DWGSIM_GPU_ESINK = -6,         /* the sink callback returned non-zero                              */
    DWGSIM_GPU_EUNSUPPORTED = -7,  /* option combination not implemented on the device path            */
    DWGSIM_GPU_EOVERFLOW = -8,     /* Ion Torrent read grew past the device bound (2*len+64)           */
    DWGSIM_GPU_ESTATE = -9         /* call order violated                                              */
};

/* file ids handed to the sink: the three gzFiles of dwgsim_opt_t (src/dwgsim_opt.h:51-53) */
enum { DWGSIM_GPU_FILE_BWA1 = 0, DWGSIM_GPU_FILE_BWA2 = 1, DWGSIM_GPU_FILE_BFAST = 2 };

typedef struct dwgsim_gpu dwgsim_gpu_t;

/* POD copy of the dwgsim_opt_t fields the loop reads (src/dwgsim_opt.h:21-60), after
 * dwgsim_opt_parse (so e_by is (end-start)/length, src/dwgsim_opt.c:459-460, and flow_order holds
 * codes 0..3, src/dwgsim_opt.c:404-407). */
typedef struct {
    double  e_start[2], e_by[2];   /* error_t of each end, src/dwgsim_opt.h:17-19                 */
    int32_t is_inner;              /* -i                                                           */
    int32_t dist;                  /* -d                                                           */
    double  std_dev;               /* -s                                                           */
    int32_t length[2];             /* -1 / -2 (length[1] == 0: single end)                         */
    double  mut_freq;              /* -F                                                           */
    double  rand_read;             /* -y                                                           */
    int32_t max_n;                 /* -n                                                           */
    int32_t data_type;             /* -c 0 Illumina, 1 SOLiD, 2 Ion Torrent                        */
    int32_t strandedness;          /* -S                                                           */
    int32_t read_one_strand;       /* -A                                                           */
    const int8_t *flow_order;      /* -f as codes 0..3 (NULL unless Ion Torrent)                   */
    int32_t flow_order_len;
    int32_t seed;                  /* -z (keys the Philox streams; drand48 is never used)          */
    int32_t fixed_quality;         /* -q: 0 = none, else the character                             */
    double  quality_std;           /* -Q                                                           */
    const char *read_prefix;       /* -P or NULL                                                   */
    int32_t reads_output_type;     /* -o 0 all, 1 bwa only, 2 bfast only                           */
    int32_t amplicons;             /* -a                                                           */
} dwgsim_gpu_params_t;

typedef struct {
    int64_t n_pairs;               /* pairs written (genomic + random)                             */
    int64_t n_random;              /* of which random (src/dwgsim.c:983-1097)                      */
    int64_t n_failed_attempts;     /* rejected genomic attempts (src/dwgsim.c:833-842)             */
    int64_t bytes[3];              /* FASTQ bytes handed to the sink per file id                   */
    int64_t h2d_bytes, d2h_bytes;  /* bytes copied over PCIe by this call                          */
    double  ms_simulate, ms_layout, ms_format;   /* device time per kernel group (CUDA events)     */
    double  ms_pack, ms_total;     /* host packing time; wall time of the call                     */
    int32_t n_launches;            /* kernels launched by this call                                */
    int32_t n_batches;
    int64_t raw_bytes[3];          /* uncompressed FASTQ bytes per file id (== bytes[] without compression) */
    double  ms_compress;           /* device time of the gzip kernels                               */
} dwgsim_gpu_stats_t;

/* receives FASTQ bytes strictly in pair-index order per file id.  buf points into the library's pinned ring and stays
 * valid until the NEXT sink call for the same file id or the return of dwgsim_gpu_run, whichever comes first (the ring
 * has at least two slots and a slot is reused only after the following batch was handed over): a sink may return at
 * once and finish writing in the background, provided it joins that write at its next call for the file.
 * Return 0 to continue, non-zero to abort the run (-> DWGSIM_GPU_ESINK). */
typedef int (*dwgsim_gpu_sink_fn)(void *user, int file_id, const char *buf, size_t n);

/* -- lifecycle ------------------------------------------------------------------------------ */
int  dwgsim_gpu_abi_version(void);
int  dwgsim_gpu_create(dwgsim_gpu_t **h, const dwgsim_gpu_params_t *p, int device);
/* One handle over several devices of the box (the loop being sharded: src/dwgsim.c:636; its two running counters: :423,
 * :1096).  Every entry point below works on the group: add_contig packs once, run() copies the packed genome to the other
 * devices (device to device: NVLink between peers), splits the pair-index space into batches, gives batch b to device
 * b % n_devices, and hands the batches to the sink in order -- the bytes are those of a single device.  set_shard /
 * set_exchange are not available on a group (it is its own set of ranks).  A device id may repeat (its ranks share the
 * device). */
int  dwgsim_gpu_create_group(dwgsim_gpu_t **h, const dwgsim_gpu_params_t *p, const int32_t *devices, int32_t n_devices);
int  dwgsim_gpu_group_size(const dwgsim_gpu_t *h);      /* devices behind the handle (1 for dwgsim_gpu_create) */
void dwgsim_gpu_destroy(dwgsim_gpu_t *h);
const char *dwgsim_gpu_strerror(int code);
const char *dwgsim_gpu_last_error(const dwgsim_gpu_t *h);

/* -- the seam (replaces the body of `for (ii...)`, src/dwgsim.c:636-1099) ---------------------- */
/* Queue one contig: name, seq_t (ASCII, src/mut.h:12-15), both mutseq_t (src/mut.h:42-47: s[len]
 * of 64-bit mut_t plus the long-insertion byte strings) and the pair budget computed by the
 * driver (src/dwgsim.c:535-591).  Everything is packed and copied before the call returns, because
 * the caller frees / reallocs these per contig (src/dwgsim.c:1103-1104). */
int dwgsim_gpu_add_contig(dwgsim_gpu_t *h, int32_t contig_i, const char *name,
                          const uint8_t *seq_ascii, int32_t len,
                          const uint64_t *hap1, const uint64_t *hap2,
                          uint8_t *const *ins1, int32_t ins1_n,
                          uint8_t *const *ins2, int32_t ins2_n,
                          int64_t n_pairs);
/* The same in two calls, for hosts that pipeline: pack_contig does all the host work of add_contig on the caller's arrays
 * (they may be freed when it returns) and may run on another thread while dwgsim_gpu_run is in flight on the handle;
 * add_packed queues the result (and takes ownership of it) between two runs. */
typedef struct dwgsim_gpu_packed dwgsim_gpu_packed_t;
int dwgsim_gpu_pack_contig(const dwgsim_gpu_t *h, int32_t contig_i, const char *name,
                           const uint8_t *seq_ascii, int32_t len, const uint64_t *hap1, const uint64_t *hap2,
                           uint8_t *const *ins1, int32_t ins1_n, uint8_t *const *ins2, int32_t ins2_n,
                           int64_t n_pairs, dwgsim_gpu_packed_t **out);
int dwgsim_gpu_add_packed(dwgsim_gpu_t *h, dwgsim_gpu_packed_t *p);
void dwgsim_gpu_packed_free(dwgsim_gpu_packed_t *p);
/* Optional: allocate the batch workspace and the pinned ring now (they are allocated by the first run() otherwise; page-locking
 * the ring takes about a second per GB).  Call it after set_batch / set_compression, from the thread that will call run() or
 * strictly before it; a host can hide it behind its own start-up work. */
int dwgsim_gpu_warm(dwgsim_gpu_t *h);
/* threads the packer may use per contig (0 = default: the hardware threads, at most 32; several ranks on one host
 * should share the cores) */
int dwgsim_gpu_set_host_threads(dwgsim_gpu_t *h, int32_t n);
/* -x (targeted regions) for the contig just queued by add_contig: its regions as regions_bed_init leaves them
 * (src/regions_bed.c:43-115: BED half-open, sorted by start, overlapping ones merged) and `sample_len`, the `l`
 * the reference's sampler draws in at that point (src/dwgsim.c:539-553: the total region length, except for the
 * last contig under -N, which keeps its full length).  The device then restates src/dwgsim.c:677-713: the position
 * is drawn in [0, sample_len - d], mapped through the regions, and the draw is repeated (as a new attempt) unless
 * one region holds [pos, pos + d] (regions_bed_query, src/regions_bed.c:117-141).  n = 0 is allowed and, like the
 * reference, can never place a pair (ETRIALS instead of the reference's endless loop).  n_pairs passed to
 * add_contig must already be the region-based budget. */
int dwgsim_gpu_set_regions(dwgsim_gpu_t *h, const uint32_t *start, const uint32_t *end, int32_t n, int32_t sample_len);
/* Simulate every queued pair, stream FASTQ to the sink in pair order, then drop the queued
 * contigs.  Pair indices (the Philox key), `ctr` and `rand_ii` (src/dwgsim.c:423) carry over to
 * the next add_contig/run round, so calling run() once per contig or once at the end yields the
 * same bytes. */
int dwgsim_gpu_run(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats);

/* -- knobs ---------------------------------------------------------------------------------- */
/* pairs per device batch (default 1<<17); ring = number of pinned output slots (default 3) */
int dwgsim_gpu_set_batch(dwgsim_gpu_t *h, int64_t pairs_per_batch, int32_t ring_slots);
/* 0 (default): the sink receives FASTQ text.  1: the sink receives gzip (RFC 1952) bytes -- every batch of every
 * stream is a run of complete gzip members written on the device (64 KiB of FASTQ each, literal-only dynamic-Huffman
 * block with a per-stream code fitted to the first batch), so appending them to <prefix>.*.fastq.gz gives the file
 * the reference writes through gzFile (src/dwgsim.c:1151-1157), at a fraction of the PCIe bytes and no host zlib. */
int dwgsim_gpu_set_compression(dwgsim_gpu_t *h, int32_t mode);
/* shard the pair-index space: dwgsim_gpu_run simulates the batches b with b % world == rank and hands
 * only those to the sink (in order); concatenating the ranks' batches round-robin gives the bytes of
 * the unsharded run.  rand_ii (src/dwgsim.c:1096) is a running count over ALL pairs, so every round
 * the ranks exchange how many random pairs their batch held: */
int dwgsim_gpu_set_shard(dwgsim_gpu_t *h, int32_t rank, int32_t world);
/* called once per round by every rank (a collective: e.g. an all-gather of my_random over NCCL);
 * must return in *before_me the sum over the lower ranks and in *round_total the sum over all ranks.
 * Return 0 on success. */
typedef int (*dwgsim_gpu_exchange_fn)(void *user, int64_t round, int64_t my_random,
                                      int64_t *before_me, int64_t *round_total);
int dwgsim_gpu_set_exchange(dwgsim_gpu_t *h, dwgsim_gpu_exchange_fn fn, void *user);
/* the first pair index / random-pair serial this handle starts from (default 0 / 0) */
int dwgsim_gpu_set_origin(dwgsim_gpu_t *h, int64_t first_pair_index, int64_t first_rand_serial);

/* -- device-resident interface (multi-GPU broadcast, benchmarks) ------------------------------ */
/* Finish packing the queued contigs into ONE position-independent blob in HBM (2-bit reference,
 * N mask, per-haplotype sparse mutation tables + block index, contig table).  Rank 0 exports it,
 * the other ranks import a byte-identical copy (e.g. after an NCCL broadcast into their own
 * device memory): no host data is needed on those ranks. */
int dwgsim_gpu_genome_finalize(dwgsim_gpu_t *h);
int dwgsim_gpu_genome_blob(const dwgsim_gpu_t *h, uint64_t *device_ptr, uint64_t *n_bytes);
int dwgsim_gpu_genome_import(dwgsim_gpu_t *h, uint64_t device_ptr, uint64_t n_bytes, int32_t take_ownership);
int64_t dwgsim_gpu_genome_pairs(const dwgsim_gpu_t *h);   /* total pairs queued over all contigs */

typedef struct {
    uint64_t dev_ptr[3];           /* device addresses of the three packed FASTQ streams           */
    uint64_t n_bytes[3];
    int64_t  n_pairs, n_random, n_failed_attempts;
    double   ms_simulate, ms_layout, ms_format;   /* CUDA-event time of each kernel group          */
    int32_t  n_launches;
} dwgsim_gpu_batch_t;
/* Simulate pairs [first, first+n) of the resident genome into device memory and leave them
 * there (no D2H).  rand_serial_base = number of random pairs before `first`.  Synchronous. */
int dwgsim_gpu_simulate_resident(dwgsim_gpu_t *h, int64_t first, int64_t n, int64_t rand_serial_base,
                                 dwgsim_gpu_batch_t *out);
/* the same in two phases, for sharded runs: begin() simulates and reports the batch's random-pair
 * count (pass NULL to skip the host sync), finish() lays the records out from rand_serial_base */
int dwgsim_gpu_resident_begin(dwgsim_gpu_t *h, int64_t first, int64_t n, int64_t *n_random);
int dwgsim_gpu_resident_finish(dwgsim_gpu_t *h, int64_t rand_serial_base, dwgsim_gpu_batch_t *out);
/* The same exchange without a host round trip (sharded runs over NCCL): after begin(.., NULL) the batch's random-pair
 * count sits in device memory at *count_device_ptr (one uint64, valid in the order of dwgsim_gpu_cuda_stream), and
 * finish_dev() reads rand_serial_base (one uint64) from device memory when its kernels run.  The caller enqueues the
 * collective and the prefix arithmetic between the two calls on that stream. */
int dwgsim_gpu_resident_count_ptr(dwgsim_gpu_t *h, uint64_t *count_device_ptr);
int dwgsim_gpu_resident_finish_dev(dwgsim_gpu_t *h, uint64_t rand_serial_base_device_ptr, dwgsim_gpu_batch_t *out);
/* Batches queued back to back, no host round trip between them.  enqueue() = begin + finish on the library's stream, returning
 * at once; the running count of random pairs (rand_ii, src/dwgsim.c:1096) lives in device memory: set_running() sets it,
 * every enqueue()d batch starts from it and adds its own.  finish_async() is finish_dev() without the wait (sharded runs: the
 * caller's collective wrote rand_serial_base to device memory; the library's counter is left alone).  wait() waits for
 * everything queued and describes the LAST batch (sizes, pointers, timings); n_launches and the error status cover every
 * batch since the previous wait.  Each batch overwrites the previous one's device buffers. */
int dwgsim_gpu_resident_set_running(dwgsim_gpu_t *h, int64_t rand_serial);
int dwgsim_gpu_resident_enqueue(dwgsim_gpu_t *h, int64_t first, int64_t n);
int dwgsim_gpu_resident_finish_async(dwgsim_gpu_t *h, uint64_t rand_serial_base_device_ptr);
/* the same with the prefix arithmetic done by the library: counts_device_ptr = the all-gathered random-pair counts of the round
 * (uint64[world], rank order = batch order, on the library's stream); this rank's batch starts at the running count + the counts
 * of the ranks before it, and the running count advances by all of them */
int dwgsim_gpu_resident_finish_gathered(dwgsim_gpu_t *h, uint64_t counts_device_ptr, int32_t world, int32_t rank);
int dwgsim_gpu_resident_wait(dwgsim_gpu_t *h, dwgsim_gpu_batch_t *out);
/* copy one stream of the last resident batch to host memory (tests) */
int dwgsim_gpu_copy_stream(dwgsim_gpu_t *h, int file_id, char *dst, uint64_t cap);
/* Queue a synthetic genome built procedurally inside the library (benchmarks only; no dense
 * host arrays are needed): contigs "chr1".. of the given lengths, i.i.d. uniform ACGT, N runs
 * covering n_frac of each contig (10 kb telomeres + one long run), SNP / deletion / insertion
 * events at mut_rate with 1/3 homozygous; coverage gives the pair budget (src/dwgsim.c:589). */
int dwgsim_gpu_genome_synthetic(dwgsim_gpu_t *h, int32_t n_contigs, const int32_t *lengths,
                                uint64_t seed, double mut_rate, double indel_frac, double n_frac,
                                double coverage);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on */
void *dwgsim_gpu_cuda_stream(const dwgsim_gpu_t *h);

/* Host-side encoder of the gzip member format the device writer emits (one literal-only dynamic-Huffman
 * block per 64 KiB member, code fitted to the data's byte histogram).  Exists so the Huffman / header / CRC
 * tables can be validated against zlib without a GPU; not used on the product path. */
int dwgsim_gpu_gz_host_encode(const uint8_t *data, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n);

/* -- ready-made sinks for dwgsim_gpu_run ---------------------------------------------------------- */
/* user = int64_t[4]: bytes per file id and the number of calls */
int dwgsim_gpu_sink_count(void *user, int file_id, const char *buf, size_t n);
/* user = int[3]: a file descriptor per file id (-1 discards); write(2)s every chunk */
int dwgsim_gpu_sink_fd(void *user, int file_id, const char *buf, size_t n);

/* File sink with one writer thread per file id: user = the handle of dwgsim_gpu_file_sink_open (fd[k] = -1 discards,
 * offset[k] = where the first byte of file k goes; NULL: 0).  A chunk is written positionally in the background and joined
 * at the next chunk for the same file or by close(), which also reports the end offsets and returns non-zero if any write
 * failed.  Call close() after dwgsim_gpu_run returned and before the next run on the same handle. */
typedef struct dwgsim_gpu_file_sink dwgsim_gpu_file_sink_t;
dwgsim_gpu_file_sink_t *dwgsim_gpu_file_sink_open(const int32_t fd[3], const int64_t offset[3]);
int dwgsim_gpu_sink_files(void *user, int file_id, const char *buf, size_t n);
int dwgsim_gpu_file_sink_close(dwgsim_gpu_file_sink_t *f, int64_t offset_out[3]);
int dwgsim_gpu_pwrite_all(int fd, const char *buf, size_t n, int64_t offset);   /* 0 when every byte was written */

/* -- derived tables (exposed so tests can compare them with the oracle's) ---------------------- */
typedef struct {
    uint64_t thr_genomic, thr_hap0;
    int32_t  isize_lo, isize_n;
    const uint32_t *isize_cdf;
    int32_t  qdelta_lo, qdelta_n;
    const uint32_t *qdelta_cdf;
    int32_t  n_cycles[2];
    const uint32_t *err_gap[2], *err_acc[2];
    const uint8_t  *qbase[2];
    uint32_t flow_thr[2];
    const uint32_t *flow_gap[2];   /* Ion Torrent: geometric gap CDF of the per-flow error coin (flow_gap_n entries; 1 otherwise) */
    int32_t  flow_gap_n[2];
} dwgsim_gpu_tables_t;
int dwgsim_gpu_tables(const dwgsim_gpu_t *h, dwgsim_gpu_tables_t *out);   /* host copies */

#ifdef __cplusplus
}
#endif
#endif
