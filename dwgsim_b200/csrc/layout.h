// layout.h -- HBM data layout shared by the host packer and the sm_100a kernels (DESIGN.md "Data layout").
#pragma once
#include <stdint.h>

namespace dwg {

// ---- genome blob --------------------------------------------------------------------------------
// One position-independent allocation: [BlobHeader][ContigDesc x n][names][per-contig sections...].
// All offsets are bytes from the blob base, 256-byte aligned.
constexpr uint32_t kBlobMagic = 0x42475744u;   // "DWGB"
constexpr uint32_t kBlobVersion = 2;
constexpr int kBlkShift = 7;                   // mutation block index granularity: 128 bases

struct BlobHeader {
    uint32_t magic, version;
    uint32_t n_contigs, flags;                 // flags bit 0: some contig carries -x regions
    uint64_t n_bytes;
    uint64_t contigs_off;                      // ContigDesc[n_contigs]
    uint64_t names_off;                        // concatenated contig names
    int64_t  total_pairs;
    int64_t  total_len;
};

struct ContigDesc {
    int32_t  len;                              // bases (the reference's `l`, src/dwgsim.c:519)
    int32_t  contig_i;
    int64_t  pair_base;                        // first queue-relative pair index of this contig
    int64_t  n_pairs;
    uint64_t ref2_off;                         // uint32 words, 16 bases each, base p at bits 2*(p&15)
    uint64_t nmask_off;                        // uint32 words, bit p&31 set when the symbol is not ACGT
    uint64_t ev_off[2];                        // Event[n_ev[h]] sorted by pos, one per non-plain position
    uint64_t blk_off[2];                       // uint32[(len >> kBlkShift) + 2]: first event with pos >= b*128
    uint64_t pool_off[2];                      // long insertions, 2-bit packed, forward order
    uint32_t n_ev[2];
    uint32_t name_off, name_len;
    // -x (src/dwgsim.c:539-581,677-713): sample_len > 0 switches the position sampler to region space
    uint64_t reg_off;                          // Region[n_reg] sorted by start, disjoint (src/regions_bed.c:82-97 merges)
    uint32_t n_reg;
    int32_t  sample_len;                       // the `l` the sampler draws in: 0 = no -x (use len)
};

struct Region {
    uint32_t start, end;                       // BED half-open
    uint32_t cum;                              // total length of the contig's regions before this one
    uint32_t pad;
};

// One entry of a haplotype's sparse mutation table = one mut_t that differs from the plain
// NOCHANGE|base entry (reference encoding: src/mut.h:25-47, src/mut.c:98-117).
//   meta bits 0-1: 0 base override (NOCHANGE whose base differs from the reference symbol, a
//                  side effect of mut_left_justify next to N runs), 1 INSERT, 2 SUBSTITUTE, 3 DELETE
//   meta bits 2-4: the entry's base code (c & 7)
//   meta bits 5-31: insertion length n
//   payload: n <= 32: the inserted bases, 2 bits each, first inserted base in the lowest bits;
//            n  > 32: base offset of the insertion in the haplotype's pool
struct Event {
    uint32_t pos;
    uint32_t meta;
    uint64_t payload;
};
constexpr uint32_t kEvOverride = 0, kEvInsert = 1, kEvSubst = 2, kEvDelete = 3;
constexpr uint32_t kInlineInsMax = 32;

// ---- per-pair intermediate record (simulate kernel -> layout + format kernels) --------------------
struct PairRec {                               // 32 bytes
    uint32_t pos[2];                           // ext_coor + 1 (src/dwgsim.c:926); 0 for random pairs
    uint16_t len[2];                           // emitted read lengths (Ion Torrent reads change length)
    uint16_t n_err[2], n_sub[2], n_indel[2];
    uint16_t n_indel_first[2];                 // insertions crossed (src/dwgsim.c:97-98), SOLiD bwa names
    uint8_t  n_err_first;                      // bit j: error on the first colour of end j
    uint8_t  flags;                            // see below
    uint16_t attempt;                          // attempt index that succeeded (keys the quality draws)
};
static_assert(sizeof(PairRec) == 32, "the format kernel reads words 2 and 7 of a PairRec");
constexpr uint8_t kRecRandom = 1, kRecStrand0 = 2, kRecStrand1 = 4, kRecHap1 = 8, kRecFailed = 0x80;

// ---- Philox addressing (DESIGN.md "RNG addressing"; oracle/dwgsim_oracle.c restates it) ---------
constexpr uint32_t kPhiloxKey1 = 0x44574753u;  // "DWGS"
enum : uint32_t { kStPair = 0, kStRandBase = 1, kStErr = 2, kStFlowU = 3, kStQual = 4, kStFlow = 5 };   // kStFlow: gaps of the
                                               // Ion Torrent error coin, kStFlowU: the uniforms of its events
enum : uint32_t { kPairGate = 0, kPairIsize = 1, kPairPosHi = 2, kPairPosLo = 3, kPairHap = 4, kPairStrand = 5 };

// ---- kernel parameters ---------------------------------------------------------------------------
struct SimParams {
    int32_t  len[2];                           // requested read lengths
    int32_t  cap[2];                           // storage per end (== len, or 2*len+64 for Ion Torrent)
    int32_t  is_inner, max_n, data_type, strandedness, read_one_strand, amplicons;
    int32_t  regions;                          // the resident genome carries -x regions (BlobHeader.flags bit 0)
    uint32_t seed;
    uint64_t thr_genomic, thr_hap0;
    int32_t  isize_lo, isize_n;
    int32_t  qdelta_lo, qdelta_n;
    int32_t  q_wrap;                           // the quality sum can leave the int8 range (emulate the reference's char arithmetic)
    uint32_t inv_name_chunks;                  // 2^32 / (name_cap / 16) + 1
    int32_t  fixed_quality, out_bwa, out_bfast;
    int32_t  prefix_len;                       // strlen(read_prefix)+1 ("pfx_"), 0 if none
    int32_t  flow_order_len;
    uint32_t flow_thr[2];
    int32_t  nw[2];                            // 32-bit words of nibble-packed read codes per end (8 codes per word);
                                               // stored word-major: word w of pair p at seqw[(w0[end] + w) * n + p]
    int32_t  row_stride;                       // words between the staged rows of two threads in the simulate kernel: nw0+nw1 when
                                               // that keeps shared-memory conflicts at two ways (vector flush), else padded to odd
    int32_t  win_slots;                        // 8-byte words of the simulate kernel's shared-memory reference window per thread
                                               // (0: reads go straight to HBM / L2; chosen on the host by occupancy)
    int32_t  tile_pairs;                       // pairs per warp mini-tile of the format kernel
    int32_t  fmt_warps;                        // warps per CTA of the format kernel (16; fewer when long reads need the shared memory)
    uint32_t inv_nw, inv_groups;               // 2^32 / (nw0+nw1) + 1 and 2^32 / (groups per pair) + 1: divisions by multiply-high
    int32_t  name_cap;                         // bytes reserved per read name in shared memory
    int32_t  rec_cap[3];                       // upper bound of a pair's bytes per output stream
    // device tables
    const uint32_t *isize_cdf, *qdelta_cdf;
    const uint32_t *qguide;                    // [1024][2] one-load guide into qdelta_cdf (see qdelta_rank)
    const uint8_t  *qtab;                      // [2^kQTabBits] noise rank by the upper bits of a draw; bit 7: the cell holds a threshold
    uint32_t qkey[10];                         // Philox round keys of the seed: seed + r * 0x9E3779B9
    int32_t  fmt_v2;                           // 1: format_fastq2_kernel (word-granular), 0: format_fastq_kernel (int8 wrap / wide noise)
    int32_t  tp_tables;                        // simulate kernel, sampling tables copied to shared memory: 3 all, 2 without the three
                                               // guides, 1 also without the error gap / accept tables, 0 none (chosen by occupancy)
    const uint16_t *isize_guide, *gap_guide[2]; // [1025] the same for isize_cdf and err_gap[end] (first len[end] entries)
    const uint32_t *err_gap[2], *err_acc[2];   // substitution errors by thinning (DESIGN.md "RNG addressing")
    const uint8_t  *qbase[2];
    const int8_t   *flow_order;
    const uint32_t *flow_gap[2];               // Ion Torrent: [kFlowGapN] geometric gap CDF of the per-flow error coin of each end
    const char     *prefix;
};
constexpr int kFlowGapN = 4096;                // a draw beyond the table: kFlowGapN failures, then a new draw (oracle: ORC_FLOW_GAP_N)

}  // namespace dwg
