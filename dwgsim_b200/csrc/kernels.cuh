// kernels.cuh -- sm_100a kernels of the dwgsim_core read-pair path.
//
//   simulate_pairs_tp_kernel   Illumina / SOLiD, one THREAD per read pair (no work is replicated across lanes):
//                           reference src/dwgsim.c:649-882 and :983-1001 (gate, insert size, position,
//                           haplotype, strands, the walk through the mutation table = __gen_read
//                           src/dwgsim.c:75-153 eight bases at a time with bit-parallel 2-bit -> nibble
//                           expansion, N filter, colour encoding, substitution errors)
//                           -> 32-byte PairRec + nibble-packed read codes (word-major, coalesced)
//   simulate_pairs_kernel   Ion Torrent, one warp per read pair: same sampling and walk, then the
//                           flow-space error model (generate_errors_flows, src/dwgsim.c:246-417)
//   layout_* kernels        exclusive scans that turn per-pair record lengths (a function of the name
//                           fields, src/dwgsim.c:923-929) and random-pair flags (rand_ii, :1096) into
//                           byte offsets of every record in the three output streams
//   format_fastq_kernel     one CTA per tile of pairs, one thread per 8-base group: quality strings
//                           (src/dwgsim.c:899-918) and the three record layouts (src/dwgsim.c:920-980)
//                           staged in shared memory and copied out with aligned 16-byte stores
//
// Integer / byte work only: no tensor cores, no floating point on the device (every probability is a
// 32-bit threshold built on the host, DESIGN.md "RNG addressing").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.h"
#include "flow_model.h"

namespace dwg {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
#ifndef DWG_SCAN_ITEMS
#define DWG_SCAN_ITEMS 1
#endif
constexpr int kScanItems = DWG_SCAN_ITEMS;       // pairs per thread of the layout kernels (1: most threads in flight; the kernels wait on loads)
constexpr int kScanTile = kThreads * kScanItems;   // pairs per layout block
constexpr int kMaxTrials = 10000;                  // src/dwgsim.c:837

// ---- Philox4x32-10 ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct PairKey {                       // what addresses the draws of one attempt at one pair
    uint32_t seed, lo, hi, attempt;
};
__device__ __forceinline__ uint4 draw_block(const PairKey &k, uint32_t stream, uint32_t end, uint32_t blk)
{
    return philox4x32_10(k.lo, k.hi, (k.attempt & 0xFFFFu) | (stream << 16) | (end << 24), blk, k.seed, kPhiloxKey1);
}
__device__ __forceinline__ uint32_t word_of(const uint4 &v, uint32_t w)
{
    return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w));
}

// #{j < n : u >= cdf[j]} for a non-decreasing table
__device__ __forceinline__ int table_rank(const uint32_t *__restrict__ cdf, int n, uint32_t u)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (u >= __ldg(cdf + mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the same rank through a guide: guide[g] = rank of (g << 22), g = 0..1024, so the search is confined to the
// thresholds inside u's 2^22-wide bucket (none or one in the bulk of a distribution, many only in its tails)
__device__ __forceinline__ int guided_rank(const uint32_t *cdf, const uint16_t *guide, uint32_t u)
{
    int lo = guide[u >> 22], hi = guide[(u >> 22) + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u >= cdf[mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- genome access ------------------------------------------------------------------------------------
struct ContigView {
    int len;
    const uint32_t *ref2, *nmask;
    const Event *ev[2];
    const uint32_t *blk[2];
    const uint8_t *pool[2];
    int n_ev[2];
};
__device__ __forceinline__ const ContigDesc *find_contig(const uint8_t *blob, int64_t q, int *index)
{
    const BlobHeader *hd = reinterpret_cast<const BlobHeader *>(blob);
    const ContigDesc *cd = reinterpret_cast<const ContigDesc *>(blob + hd->contigs_off);
    int lo = 0, hi = (int)hd->n_contigs - 1;
    while (lo < hi) {                  // last contig with pair_base <= q
        int mid = (lo + hi + 1) >> 1;
        if (cd[mid].pair_base <= q) lo = mid; else hi = mid - 1;
    }
    *index = lo;
    return cd + lo;
}
__device__ __forceinline__ ContigView view_of(const uint8_t *blob, const ContigDesc *cd)
{
    ContigView v;
    v.len = cd->len;
    v.ref2 = reinterpret_cast<const uint32_t *>(blob + cd->ref2_off);
    v.nmask = reinterpret_cast<const uint32_t *>(blob + cd->nmask_off);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        v.ev[h] = reinterpret_cast<const Event *>(blob + cd->ev_off[h]);
        v.blk[h] = reinterpret_cast<const uint32_t *>(blob + cd->blk_off[h]);
        v.pool[h] = blob + cd->pool_off[h];
        v.n_ev[h] = (int)cd->n_ev[h];
    }
    return v;
}
__device__ __forceinline__ uint32_t ref_code(const ContigView &c, int p)
{
    uint32_t w = __ldg(c.ref2 + (p >> 4));
    uint32_t m = __ldg(c.nmask + (p >> 5));
    return ((m >> (p & 31)) & 1u) ? 4u : ((w >> ((p & 15) << 1)) & 3u);
}
// -x: map a position drawn in region space to the contig and require one region to hold the whole fragment
// (src/dwgsim.c:695-713 with regions_bed_query, src/regions_bed.c:117-141).  Returns false where the reference
// repeats its draw (returned as -1); here that is a rejected attempt (draws are addressed by attempt).
// Both helpers look the contig up again instead of keeping its region fields live across the read walk.
__device__ __noinline__ int region_sample_len(const uint8_t *blob, int64_t q)
{
    int index;
    const ContigDesc *cd = find_contig(blob, q, &index);
    return cd->sample_len > 0 ? cd->sample_len : cd->len;
}
__device__ __noinline__ int map_to_regions(const uint8_t *blob, int64_t q, int pos, int d)
{
    int index;
    const ContigDesc *cd = find_contig(blob, q, &index);
    if (cd->sample_len <= 0) return pos;                               // a contig queued without -x
    struct { const Region *reg; int n_reg, len; } c{reinterpret_cast<const Region *>(blob + cd->reg_off), (int)cd->n_reg, cd->len};
    int lo = 0, hi = c.n_reg - 1, k = -1;
    while (lo <= hi) {                 // last region with cum <= pos
        const int mid = (lo + hi) >> 1;
        if (__ldg(&c.reg[mid].cum) <= (uint32_t)pos) { k = mid; lo = mid + 1; } else hi = mid - 1;
    }
    int total = 0;
    if (c.n_reg > 0) { const uint4 r = __ldg(reinterpret_cast<const uint4 *>(c.reg + c.n_reg - 1)); total = (int)(r.z + (r.y - r.x)); }
    if (k >= 0) {
        const uint4 r = __ldg(reinterpret_cast<const uint4 *>(c.reg + k));
        const int off = pos - (int)r.z;
        if (off < (int)(r.y - r.x)) pos = (int)r.x + off - 1;          // "zero-based", src/dwgsim.c:700
        else pos -= total;                                              // past the last region: left unmapped
    }
    if (pos < 0 || pos >= c.len || pos + d - 1 >= c.len) return -1;
    lo = 0; hi = c.n_reg - 1; k = -1;
    while (lo <= hi) {                 // last region with start <= pos
        const int mid = (lo + hi) >> 1;
        if (__ldg(&c.reg[mid].start) <= (uint32_t)pos) { k = mid; lo = mid + 1; } else hi = mid - 1;
    }
    return (k >= 0 && (uint32_t)(pos + d) <= __ldg(&c.reg[k].end)) ? pos : -1;
}
__device__ __forceinline__ uint4 load_event(const Event *ev, int e)
{
    return __ldg(reinterpret_cast<const uint4 *>(ev + e));
}
__device__ __forceinline__ uint32_t ins_code(const uint4 &ev, const uint8_t *pool, uint32_t n, uint32_t j)
{
    uint64_t payload = (uint64_t)ev.z | ((uint64_t)ev.w << 32);
    if (n <= kInlineInsMax) return (uint32_t)(payload >> (2 * j)) & 3u;
    uint64_t at = payload + j;
    return (__ldg(pool + (at >> 2)) >> ((at & 3) << 1)) & 3u;
}

// ---- the walk (__gen_read, src/dwgsim.c:75-153) over the sparse table -----------------------------------
// Scalar state (position, emitted count, counters) is warp-uniform; lanes share the emission of each
// run of plain reference bases and of each insertion.  Returns false when the reference would leave
// ext_coor negative (ran off the contig, or the predicted left end went below 0 on the minus strand).
struct Walk {
    int ext, n_sub, n_indel, n_indel_first;
};
__device__ __forceinline__ bool gen_read(const ContigView &c, int h, int start, int strand, int s,
                                         uint8_t *out, int lane, Walk &w)
{
    const int dir = strand ? -1 : 1;
    const Event *ev = c.ev[h];
    const int n_ev = c.n_ev[h];
    w.ext = -10; w.n_sub = w.n_indel = w.n_indel_first = 0;
    int i = start;
    if (i < 0 || i >= c.len) return false;
    int e;
    if (dir > 0) {
        e = (int)__ldg(c.blk[h] + (i >> kBlkShift));
        while (e < n_ev && (int)__ldg(&ev[e].pos) < i) ++e;
    } else {
        e = (int)__ldg(c.blk[h] + (i >> kBlkShift) + 1) - 1;
        while (e >= 0 && (int)__ldg(&ev[e].pos) > i) --e;
    }
    bool have = dir > 0 ? (e < n_ev) : (e >= 0);
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (have) cur = load_event(ev, e);
    // a read may only start on a NOCHANGE / SUBSTITUTE position (src/dwgsim.c:78-82)
    while (have && (int)cur.x == i) {
        uint32_t t = cur.y & 3u;
        if (t != kEvInsert && t != kEvDelete) break;
        i += dir; e += dir;
        if (i < 0 || i >= c.len) return false;
        have = dir > 0 ? (e < n_ev) : (e >= 0);
        if (have) cur = load_event(ev, e);
    }
    int ext = i - (strand ? s - 1 : 0);
    if (ext < 0) return false;
    int k = 0;
    while (k < s) {
        const int pe = have ? (int)cur.x : (dir > 0 ? c.len : -1);
        int run = dir > 0 ? pe - i : i - pe;
        if (run > s - k) run = s - k;
        for (int j = lane; j < run; j += 32) out[k + j] = (uint8_t)ref_code(c, i + dir * j);
        k += run; i += dir * run;
        if (k == s) break;
        if (!have) return false;                             // walked off the contig
        const uint32_t t = cur.y & 3u, base = (cur.y >> 2) & 7u, n = cur.y >> 5;
        if (t == kEvSubst || t == kEvOverride) {
            if (lane == 0) out[k] = (uint8_t)base;
            ++k;
            if (t == kEvSubst) ++w.n_sub;
        } else if (t == kEvDelete) {
            ++w.n_indel;
            if (strand && --ext < 0) return false;
        } else {
            ++w.n_indel; ++w.n_indel_first;
            if (!strand) {
                if (lane == 0) out[k] = (uint8_t)base;
                ++k;
                int m = (int)n < s - k ? (int)n : s - k;
                for (int j = lane; j < m; j += 32) out[k + j] = (uint8_t)ins_code(cur, c.pool[h], n, (uint32_t)j);
                k += m;
            } else {
                int m = (int)n < s - k ? (int)n : s - k;
                ext += m;
                for (int j = lane; j < m; j += 32) out[k + j] = (uint8_t)ins_code(cur, c.pool[h], n, n - 1u - (uint32_t)j);
                k += m;
                if (k < s) { if (lane == 0) out[k] = (uint8_t)base; ++k; }
            }
        }
        i += dir; e += dir;
        have = dir > 0 ? (e < n_ev) : (e >= 0);
        if (have) cur = load_event(ev, e);
    }
    __syncwarp();
    if (strand)
        for (int j = lane; j < s; j += 32) { uint8_t v = out[j]; out[j] = v < 4 ? (uint8_t)(3 - v) : (uint8_t)4; }
    __syncwarp();
    w.ext = ext;
    return true;
}

// colour encoding with adaptor base 0 (src/dwgsim.c:845-858); chunks are processed high to low so that
// every predecessor is still a base when it is read
__device__ __forceinline__ void colour_encode(uint8_t *seq, int n, int lane)
{
    for (int r = (n - 1) >> 5; r >= 0; --r) {
        int k = (r << 5) + lane;
        uint32_t cur = 0, prev = 0;
        if (k < n) { cur = seq[k]; prev = k ? seq[k - 1] : 0u; }
        __syncwarp();
        if (k < n) seq[k] = (uint8_t)((cur >= 4 || prev >= 4) ? 4u : (cur ^ prev));
        __syncwarp();
    }
}

// ---- Ion Torrent flow model (generate_errors_flows, src/dwgsim.c:246-417) -------------------------------
// The model is sequential per read (every error shifts the rest of the read and the flow phase), so the
// scalar control flow runs warp-uniform (all lanes execute the same decisions on the same sequential FLOW
// draws) and the lanes share the work that is parallel: N removal, reversal, and the shifts of the read tail.
struct FlowRng {
    PairKey key;
    uint32_t end, next;
    uint4 blk;
    int have;
    // the per-flow error coin ("drand48() < e", src/dwgsim.c:290,372) as a Bernoulli process generated from its geometric
    // gaps: `left` failures are still to come before the next success (succ) or the next draw (!succ); one FLOW draw per
    // success instead of one per trial (oracle: flow_coin)
    int left = 0;
    bool succ = false, need = true;
    __device__ __forceinline__ uint32_t draw()                 // next word of the gap stream
    {
        const int b = (int)(next >> 2);
        if (b != have) { have = b; blk = draw_block(key, kStFlow, end, (uint32_t)b); }
        return word_of(blk, next++ & 3u);
    }
    uint32_t unext = 0;
    __device__ __forceinline__ uint32_t unif()                 // next word of the events' uniform stream (oracle: flow_unif)
    {
        const uint4 b = draw_block(key, kStFlowU, end, unext >> 2);
        return word_of(b, unext++ & 3u);
    }
    __device__ __forceinline__ bool coin(const uint32_t *__restrict__ gap)
    {
        for (;;) {
            if (need) { const int g = table_rank(gap, kFlowGapN, draw()); left = g; succ = g < kFlowGapN; need = false; }
            if (left > 0) { --left; return false; }
            need = true;
            if (succ) return true;
        }
    }
};
__device__ __forceinline__ void warp_reverse(uint8_t *seq, int len, int lane)
{
    for (int j = lane; j < (len >> 1); j += 32) { uint8_t a = seq[j], b = seq[len - 1 - j]; seq[j] = b; seq[len - 1 - j] = a; }
    __syncwarp();
}
// seq[j + n] = seq[j] for j in [i, len), highest chunk first
__device__ __forceinline__ void warp_shift_up(uint8_t *seq, int i, int len, int n, int lane)
{
    for (int top = len; top > i; top -= 32) {
        const int j = top - 1 - lane;
        uint8_t v = 0;
        if (j >= i) v = seq[j];
        __syncwarp();
        if (j >= i) seq[j + n] = v;
        __syncwarp();
    }
}
// seq[j] = seq[j + n] for j in [i, len - n), lowest chunk first
__device__ __forceinline__ void warp_shift_down(uint8_t *seq, int i, int len, int n, int lane)
{
    for (int a = i; a < len - n; a += 32) {
        const int j = a + lane;
        uint8_t v = 0;
        if (j < len - n) v = seq[j + n];
        __syncwarp();
        if (j < len - n) seq[j] = v;
        __syncwarp();
    }
}
// returns the new length, -1 when the first base is not in the flow order (src/dwgsim.c:275-278);
// *overflow is set when the read would grow past `cap` symbols
__device__ __forceinline__ int flow_errors(uint8_t *seq, int len, int cap, int strand, const uint32_t *__restrict__ gap, const int8_t *fo,
                                           int fl, uint8_t *mask, FlowRng &rng, int *n_err_out, int *overflow, int lane)
{
    for (int j = lane; j < len; j += 32) if (seq[j] >= 4) seq[j] = 0;          // src/dwgsim.c:253-257
    for (int j = lane; j < fl; j += 32) mask[j] = 0;                            // per-read mask
    __syncwarp();
    if (strand) warp_reverse(seq, len, lane);
    int i, flow_i;
    {
        const int c = len > 0 ? seq[0] : 0;
        for (i = 0; i < fl; ++i) if (c == fo[i]) break;
        if (i == fl) return -1;
    }
    flow_i = i;
    int prev_c = 4;
    for (i = 0; i < len; ++i) {                                                  // src/dwgsim.c:281-364
        const int c = seq[i];
        while (c != fo[flow_i]) { if (lane == 0) mask[flow_i] = 0; flow_i = flow_i + 1 == fl ? 0 : flow_i + 1; }
        if (prev_c != c) {
            if (lane == 0) mask[flow_i] = 0;
            int n_err = 0;
            while (rng.coin(gap)) ++n_err;
            if (n_err > 0) {
                if (!(rng.unif() >> 31)) {                                       // U < 0.5: insertion
                    if (len + n_err >= cap) { *overflow = 1; return len; }
                    __syncwarp();
                    warp_shift_up(seq, i, len, n_err, lane);
                    for (int j = i + lane; j < i + n_err; j += 32) seq[j] = (uint8_t)c;
                    __syncwarp();
                    len += n_err;
                } else {                                                         // deletion, bounded by the homopolymer
                    int hp_l = 0, next_c = 4;
                    for (int j = i; j < len; ++j, ++hp_l) { next_c = seq[j]; if (c != next_c) break; }
                    if (hp_l < n_err) n_err = hp_l;
                    __syncwarp();
                    warp_shift_down(seq, i, len, n_err, lane);
                    len -= n_err;
                    if (lane == 0) mask[flow_i] = 1;
                    if (n_err == hp_l && (i == 0 || prev_c == next_c)) {         // dot-fill, src/dwgsim.c:342-358
                        int j = 0;
                        while (next_c != fo[(flow_i + j) % fl]) ++j;
                        if (j <= 0) { *overflow = 2; return len; }
                        const int k = (int)__umulhi(rng.unif(), (uint32_t)j);
                        if (len + 1 >= cap) { *overflow = 1; return len; }
                        warp_shift_up(seq, i, len, 1, lane);
                        if (lane == 0) seq[i] = (uint8_t)fo[(flow_i + k) % fl];
                        __syncwarp();
                        len += 1;
                    }
                }
                *n_err_out += n_err;
            }
            prev_c = c;
        }
    }
    __syncwarp();
    for (i = 0; i < len; ++i) {                                                  // src/dwgsim.c:366-406
        const int c = seq[i];
        while (c != fo[flow_i]) {
            int n_err = 0;
            while (rng.coin(gap)) ++n_err;
            if (n_err > 0 && mask[flow_i] == 0) {
                if (len + n_err >= cap) { *overflow = 1; return len; }
                __syncwarp();
                warp_shift_up(seq, i, len, n_err, lane);
                for (int j = i + lane; j < i + n_err; j += 32) seq[j] = (uint8_t)fo[flow_i];
                __syncwarp();
                len += n_err;
                *n_err_out += n_err;
            }
            flow_i = flow_i + 1 == fl ? 0 : flow_i + 1;
        }
    }
    __syncwarp();
    if (strand) warp_reverse(seq, len, lane);
    return len;
}

// ---- kernel A: simulate ----------------------------------------------------------------------------------
// status[0]: error bits (1 = a pair exhausted its 10001 trials), status[1]: rejected attempts
__global__ void __launch_bounds__(kThreads)
simulate_pairs_kernel(const SimParams P, const uint8_t *__restrict__ blob, int64_t first, int64_t gidx_origin, int n,
                      PairRec *__restrict__ recs, uint32_t *__restrict__ seqw, unsigned long long *__restrict__ status)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap0 = (P.cap[0] + 15) & ~15, cap1 = (P.cap[1] + 15) & ~15;
    const int flr = (P.flow_order_len + 15) & ~15;
    uint8_t *code[2];
    code[0] = smem + (size_t)warp * (cap0 + cap1 + flr);
    code[1] = code[0] + cap0;
    uint8_t *flow_mask = code[1] + cap1;
    int8_t *flow_order = reinterpret_cast<int8_t *>(smem + (size_t)kWarpsPerBlock * (cap0 + cap1 + flr));
    for (int j = threadIdx.x; j < P.flow_order_len; j += kThreads) flow_order[j] = P.flow_order[j];
    __syncthreads();
    const int warps_total = gridDim.x * kWarpsPerBlock;

    for (int p = blockIdx.x * kWarpsPerBlock + warp; p < n; p += warps_total) {
        const int64_t q = first + p;
        int contig_index;
        const ContigDesc *cd = find_contig(blob, q, &contig_index);
        const ContigView cv = view_of(blob, cd);
        const uint64_t gidx = (uint64_t)(gidx_origin + q);
        PairKey key{P.seed, (uint32_t)gidx, (uint32_t)(gidx >> 32), 0u};
        PairRec rec;
        int s[2] = {P.len[0], P.len[1]};
        int strand[2] = {0, 0};
        bool done = false, random_pair = false;
        int hap = 0;
        Walk w[2];
        unsigned failed = 0;

        for (int attempt = 0; attempt <= kMaxTrials && !done; ++attempt) {
            key.attempt = (uint32_t)attempt;
            const uint4 b0 = draw_block(key, kStPair, 0, 0);
            if ((uint64_t)b0.x < P.thr_genomic) { random_pair = true; done = true; break; }   // src/dwgsim.c:649
            int d, pos;
            if (P.amplicons) { pos = 0; d = cv.len; }                                          // src/dwgsim.c:650-653
            else {
                const int slen = P.regions ? region_sample_len(blob, q) : cv.len;
                if (s[1] > 0) {                                                                // src/dwgsim.c:656-664
                    d = P.isize_lo + table_rank(P.isize_cdf, P.isize_n, b0.y);
                    const int min_dist = s[0] + s[1];
                    if (d < min_dist) d = min_dist;
                    if (d > slen) d = slen;
                } else d = 0;
                const uint64_t range = (uint64_t)((int64_t)slen - d + 1);
                pos = (int)__umul64hi(range, ((uint64_t)b0.z << 32) | b0.w);                   // src/dwgsim.c:671
                if (P.regions && (pos = map_to_regions(blob, q, pos, d)) < 0) { ++failed; continue; }
            }
            const uint4 b1 = draw_block(key, kStPair, 0, 1);
            hap = ((uint64_t)b1.x < P.thr_hap0) ? 0 : 1;                                       // src/dwgsim.c:716
            strand[0] = P.read_one_strand == 0 ? ((b1.y >> 31) ? 0 : 1) : (P.read_one_strand == 1 ? 0 : 1);
            if (P.strandedness == 0) strand[1] = (P.data_type == 0) ? 1 - strand[0] : strand[0];
            else strand[1] = (P.strandedness == 1) ? strand[0] : 1 - strand[0];
            // read placement, src/dwgsim.c:745-821
            int st0, st1 = 0;
            const int last = cv.len - 1;
            if (s[1] > 0) {
                if (strand[0] == strand[1]) {
                    if (strand[0] == 0) {
                        st0 = P.amplicons ? last : (P.is_inner ? pos + s[1] + d - 1 : pos + d - s[0]);
                        st1 = pos;
                    } else {
                        st0 = pos + s[0] - 1;
                        st1 = P.amplicons ? last : (P.is_inner ? pos + s[0] + d + s[1] - 1 : pos + d - 1);
                    }
                } else if (strand[0] == 0) {
                    st0 = pos;
                    st1 = P.amplicons ? last : (P.is_inner ? pos + s[0] + d + s[1] - 1 : pos + d - 1);
                } else {
                    st0 = P.amplicons ? last : (P.is_inner ? pos + s[1] + d + s[0] - 1 : pos + d - 1);
                    st1 = pos;
                }
            } else {
                st0 = strand[0] == 0 ? pos : (P.amplicons ? last : pos + s[0] - 1);
            }
            bool ok = gen_read(cv, hap, st0, strand[0], s[0], code[0], lane, w[0]);
            if (s[1] > 0) {
                // the reference always walks both ends before testing (src/dwgsim.c:759-760,833)
                bool ok1 = gen_read(cv, hap, st1, strand[1], s[1], code[1], lane, w[1]);
                ok = ok && ok1;
            } else { w[1].ext = 0; w[1].n_sub = w[1].n_indel = w[1].n_indel_first = 0; }
            if (ok) {                                                                          // src/dwgsim.c:823-833
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int nn = 0;
                    for (int k = lane; k < s[j]; k += 32) nn += code[j][k] == 4;
                    nn = __reduce_add_sync(0xffffffffu, nn);
                    if (nn > P.max_n) ok = false;
                }
            }
            if (ok) done = true; else ++failed;
        }

        rec.attempt = (uint16_t)key.attempt;
        rec.n_err_first = 0;
        if (!done) {                                     // 10001 rejected attempts, src/dwgsim.c:837-840
            if (lane == 0) atomicOr(status, 1ull);
            random_pair = true;                          // keep the streams well-formed; the host reports the error
            rec.flags = kRecFailed;
        } else rec.flags = 0;

        if (random_pair) {                               // src/dwgsim.c:983-1001
            rec.flags |= kRecRandom;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                rec.pos[j] = 0; rec.len[j] = (uint16_t)s[j];
                rec.n_err[j] = rec.n_sub[j] = rec.n_indel[j] = rec.n_indel_first[j] = 0;
                uint4 blk = make_uint4(0, 0, 0, 0);
                int have_blk = -1;
                for (int k = lane; k < s[j]; k += 32) {
                    if ((k >> 6) != have_blk) { have_blk = k >> 6; blk = draw_block(key, kStRandBase, j, have_blk); }
                    code[j][k] = (uint8_t)((word_of(blk, (k >> 4) & 3) >> ((k & 15) << 1)) & 3u);
                }
            }
            __syncwarp();
            if (P.data_type == 1) { colour_encode(code[0], s[0], lane); colour_encode(code[1], s[1], lane); }
        } else {
            if (P.data_type == 1) { colour_encode(code[0], s[0], lane); colour_encode(code[1], s[1], lane); }
            rec.flags |= (strand[0] ? kRecStrand0 : 0) | (strand[1] ? kRecStrand1 : 0) | (hap ? kRecHap1 : 0);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                rec.pos[j] = (uint32_t)(w[j].ext + 1);
                rec.len[j] = (uint16_t)s[j];
                rec.n_sub[j] = (uint16_t)w[j].n_sub;
                rec.n_indel[j] = (uint16_t)w[j].n_indel;
                rec.n_indel_first[j] = (uint16_t)w[j].n_indel_first;
                if (P.data_type == 2) {                      // Ion Torrent, src/dwgsim.c:861-864
                    int nerr = 0, ovf = 0, nl = 0;
                    if (s[j] > 0) {
                        FlowRng rng{key, (uint32_t)j, 0u, make_uint4(0, 0, 0, 0), -1};
                        nl = flow_errors(code[j], s[j], P.cap[j], strand[j], P.flow_gap[j], flow_order, P.flow_order_len,
                                         flow_mask, rng, &nerr, &ovf, lane);
                    }
                    if (ovf && lane == 0) atomicOr(status, 2ull);
                    s[j] = nl > 0 ? nl : 0;
                    rec.len[j] = (uint16_t)s[j];
                    rec.n_err[j] = (uint16_t)nerr;
                    continue;
                }
                rec.n_err[j] = 0;                            // substitution errors: simulate_pairs_tp_kernel
            }
            rec.n_err_first = (uint8_t)__shfl_sync(0xffffffffu, (int)rec.n_err_first, 0);   // k == 0 lives in lane 0
            __syncwarp();
        }
        if (lane == 0) {
            recs[p] = rec;
            if (failed) atomicAdd(status + 1, (unsigned long long)failed);
        }
        // nibble-pack the read codes: 8 per 32-bit word, word-major
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t *d = seqw + (size_t)p * (P.nw[0] + P.nw[1]) + (j ? P.nw[0] : 0);
            const int nw = (s[j] + 7) >> 3;
            for (int wi = lane; wi < nw; wi += 32) {
                uint32_t v = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int k = wi * 8 + t;
                    uint32_t c = k < s[j] ? code[j][k] : 0u;
                    v |= c << (4 * t);
                }
                d[wi] = v;
            }
        }
        __syncwarp();
    }
}

// ---- explicit shared-memory accesses (32-bit shared-window addresses; the staging areas are selected at run time, and a
// generic pointer there would turn every byte store into a 64-bit generic ST) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int kOff>
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(kOff) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// a value the compiler must keep in an ordinary register (uniform registers do not survive the divergent code around
// the table lookups, and rebuilding a shared-window address costs four instructions each time)
__device__ __forceinline__ uint32_t in_register(uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }

// ---- kernel A (Illumina / SOLiD): one thread per pair ------------------------------------------------------
// The walk reads the 2-bit reference from a per-thread window in shared memory that cp.async fills before the walk
// starts: the 8-byte words (32 bases each) that cover the read plus kWindowSlack positions in walk direction.  The
// copies land without occupying a register and the walk never waits on a dependent reference load; a walk that strays
// outside its window (long deletions) falls back to direct loads.  The N mask is looked at only when the read's
// neighbourhood holds an N at all (RefWindow::has_n, decided by a few 16-byte loads before the walk).  Word indices just
// outside a contig's section are inside the blob (sections are 256-byte aligned and never last), and the bases read
// from there are never emitted.
constexpr int kWindowSlotStride = 128 * 8; // [slots][kTpThreads] x 8 bytes: lanes of a warp hit consecutive 8-byte words
constexpr int kWindowSlack = 32;
__host__ __device__ inline int window_slots(int len) { return (len + kWindowSlack + 31) / 32 + 1; }
struct RefWindow {
    uint32_t base;                     // shared-window address of this thread's slot 0
    int first, slots;                  // the window holds words first .. first + slots - 1
    bool has_n, fresh;
};
constexpr int kNCheckSlack = 64;       // the N pre-check covers the read plus this many positions in walk direction
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
// start the copies of the words a walk of s bases from `start` will read
__device__ __forceinline__ void window_prime(const ContigView &c, RefWindow &R, int start, int dir, int s, int max_slots)
{
    if (max_slots <= 0) return;                                            // no window: every group reads HBM / L2 directly
    cp_async_wait<0>();                                                    // copies of a window that was never read
    R.slots = min(window_slots(s), max_slots);
    R.first = dir > 0 ? (start >> 5) : ((start >> 5) - R.slots + 2);        // backward: the group at `start` may need word (start >> 5) + 1
    const uint2 *src = reinterpret_cast<const uint2 *>(c.ref2) + R.first;
    for (int t = 0; t < R.slots; ++t) cp_async8(R.base + (uint32_t)t * kWindowSlotStride, src + t);
    cp_async_commit();
    R.fresh = true;
}
// the N pre-check of a read of s bases starting at `start`: OR of the N bits of the 128-base-aligned blocks around
// [start, start +- (s + kNCheckSlack)]; the walk switches the mask loads on when it strays further (long deletions)
__device__ __forceinline__ uint32_t n_precheck(const ContigView &c, int start, int dir, int s)
{
    int lo = dir > 0 ? start : start - s - kNCheckSlack, hi = dir > 0 ? start + s + kNCheckSlack : start;
    lo = lo < 0 ? 0 : lo; hi = hi >= c.len ? c.len - 1 : hi;
    uint32_t any = 0;
    const uint4 *m4 = reinterpret_cast<const uint4 *>(c.nmask);
    for (int b = lo >> 7; b <= (hi >> 7); ++b) { const uint4 v = __ldg(m4 + b); any |= v.x | v.y | v.z | v.w; }
    return any;
}
// 8 consecutive bases of the 2-bit reference -> 8 nibble codes (0-3, 4 = N), optionally reversed + complemented
__device__ __forceinline__ uint32_t fetch_codes(const ContigView &c, RefWindow &R, int i, int dir, int m)
{
    const int j = dir > 0 ? i : i - m + 1;                    // lowest position of the group
    const int t = (j >> 5) - R.first;
    if (R.fresh) { cp_async_wait<0>(); R.fresh = false; }     // first use of the window
    uint2 lo, hi;
    if (t >= 0 && t + 1 < R.slots) { lo = lds64(R.base + (uint32_t)t * kWindowSlotStride); hi = lds64(R.base + (uint32_t)(t + 1) * kWindowSlotStride); }
    else { const uint2 *src = reinterpret_cast<const uint2 *>(c.ref2) + (j >> 5); lo = __ldg(src); hi = __ldg(src + 1); }
    const int pos = j & 31;
    const uint32_t wa = pos < 16 ? lo.x : lo.y, wb = pos < 16 ? lo.y : hi.x;
    uint32_t x = __funnelshift_r(wa, wb, (pos << 1) & 31) & ((1u << (2 * m)) - 1u);
    // minus strand: reverse the order of the m bases and complement them (computed for every lane, then selected)
    uint32_t xr = __brev(x) >> (32 - 2 * m);
    xr = (((xr >> 1) & 0x5555u) | ((xr & 0x5555u) << 1)) ^ ((1u << (2 * m)) - 1u);
    x = dir < 0 ? xr : x;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u;
    if (R.has_n) {                                            // the N mask of the group, same treatment, merged as code 4
        const uint32_t n0 = __ldg(c.nmask + (j >> 5)), n1 = __ldg(c.nmask + (j >> 5) + 1);
        uint32_t y = __funnelshift_r(n0, n1, j & 31) & ((1u << m) - 1u);
        const uint32_t yr = __brev(y) >> (32 - m);
        y = dir < 0 ? yr : y;
        y = (y | (y << 12)) & 0x000F000Fu; y = (y | (y << 6)) & 0x03030303u; y = (y | (y << 3)) & 0x11111111u;
        x = (x & ~(y * 3u)) | (y << 2);
    }
    return x;
}

template <int kStride>
struct EmitT {                         // emission state of one read
    uint32_t *dst;                     // this thread's row: word w at dst[w * kStride] (the CTA's shared staging tile, stride 1;
                                       //   Ion Torrent: a row in HBM interleaved with those of the warp's other lanes, stride 32)
    uint64_t acc; int na, w;           // pending nibbles, their count, next word index
    int k, nN;                         // symbols emitted, N bases seen (src/dwgsim.c:823-831)
    int solid; uint32_t prev;          // colour space: previous base, adaptor = 0 (src/dwgsim.c:845-858)
};
struct TpTables {                     // shared-memory copies of the sampling tables of the thread-per-pair kernel
    const uint32_t *isize_cdf; const uint16_t *isize_guide;
    const uint32_t *gap[2], *acc[2]; const uint16_t *gap_guide[2];
    const int8_t *flow_order; uint32_t *flow_mask;
    const uint16_t *flow_nd;           // [flow_order_len][4] steps to a base's next flow (flow_model.h)
    uint16_t *flow_q;                  // this thread's kFlowGapsAhead gaps drawn ahead
};
template <int kStride>
__device__ __forceinline__ void emit_begin(EmitT<kStride> &E, uint32_t *dst, int solid)
{
    E.dst = dst; E.acc = 0; E.na = 0; E.w = 0; E.k = 0; E.nN = 0;
    E.solid = solid; E.prev = 0;
}
// append m <= 8 base codes (nibbles in the low bits of `codes`, zero above)
template <int kStride>
__device__ __forceinline__ void emit_group(EmitT<kStride> &E, uint32_t codes, int m)
{
    E.nN += __popc(codes & 0x44444444u);
    if (E.solid) {
        const uint32_t prevs = (codes << 4) | E.prev;
        E.prev = (codes >> (4 * (m - 1))) & 15u;
        const uint32_t x = (codes ^ prevs) & 0x33333333u, nf = (codes | prevs) & 0x44444444u;
        codes = ((x & ~((nf >> 2) * 3u)) | nf) & (m == 8 ? 0xFFFFFFFFu : ((1u << (4 * m)) - 1u));
    }
    E.acc |= (uint64_t)codes << (4 * E.na);
    E.na += m; E.k += m;
    if (E.na >= 8) { E.dst[E.w * kStride] = (uint32_t)E.acc; ++E.w; E.acc >>= 32; E.na -= 8; }
}
template <int kStride>
__device__ __forceinline__ void emit_end(EmitT<kStride> &E)
{
    if (E.na > 0) { E.dst[E.w * kStride] = (uint32_t)E.acc; ++E.w; }
}

// The walk of gen_read() above executed by one thread as ONE loop: every iteration either emits up to eight
// plain reference bases or consumes one mutation event, so the lanes of a warp stay in the same loop even
// when their reads cross different events.
// first load of the walk, issued early by the caller: the block-index entry that brackets `start` (0 when off the contig)
__device__ __forceinline__ int walk_hint(const ContigView &c, int h, int start, int strand)
{
    if (start < 0 || start >= c.len) return 0;
    return (int)__ldg(c.blk[h] + (start >> kBlkShift) + (strand ? 1 : 0));
}
template <int kStride>
__device__ __forceinline__ bool walk_thread(const ContigView &c, int h, int start, int strand, int s, EmitT<kStride> &E, Walk &w, int hint,
                                            RefWindow &C)
{
    const int dir = strand ? -1 : 1;
    const Event *ev = c.ev[h];
    const int n_ev = c.n_ev[h];
    w.ext = -10; w.n_sub = w.n_indel = w.n_indel_first = 0;
    int i = start;
    if (i < 0 || i >= c.len) return false;
    int e;
    if (dir > 0) {
        e = hint;
        while (e < n_ev && (int)__ldg(&ev[e].pos) < i) ++e;
    } else {
        e = hint - 1;
        while (e >= 0 && (int)__ldg(&ev[e].pos) > i) --e;
    }
    bool have = dir > 0 ? (e < n_ev) : (e >= 0);
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (have) cur = load_event(ev, e);
    while (have && (int)cur.x == i) {                            // src/dwgsim.c:78-82
        const uint32_t t = cur.y & 3u;
        if (t != kEvInsert && t != kEvDelete) break;
        i += dir; e += dir;
        if (i < 0 || i >= c.len) return false;
        have = dir > 0 ? (e < n_ev) : (e >= 0);
        if (have) cur = load_event(ev, e);
    }
    int ext = i - (strand ? s - 1 : 0);
    if (ext < 0) return false;
    int strayed = dir > 0 ? i - start : start - i;                 // positions beyond the N pre-check's plain reach
    if (strayed > kNCheckSlack - 2) C.has_n = true;
    const uint32_t comp = strand ? 3u : 0u;                       // base b < 4 -> b ^ comp
    int k = 0;
    int ins_left = 0, ins_at = 0;                                  // insertion being emitted: bases left, next index
    bool ins_ref_pending = false;                                  // minus strand: reference base follows the insertion
    bool ok = true;
    while (k < s) {
        if (ins_left > 0) {                                        // inside an insertion: up to 8 inserted bases
            const uint32_t n = cur.y >> 5;
            const int m = ins_left < 8 ? ins_left : 8;
            uint32_t codes = 0;
            for (int j = 0; j < m; ++j) {
                const uint32_t idx = strand ? n - 1u - (uint32_t)(ins_at + j) : (uint32_t)(ins_at + j);
                codes |= (ins_code(cur, c.pool[h], n, idx) ^ comp) << (4 * j);
            }
            emit_group(E, codes, m);
            k += m; ins_left -= m; ins_at += m;
            if (ins_left > 0) continue;
            if (ins_ref_pending) {
                ins_ref_pending = false;
                if (k < s) { uint32_t base = (cur.y >> 2) & 7u; emit_group(E, base < 4 ? base ^ comp : 4u, 1); ++k; }
            }
            i += dir; e += dir;
            have = dir > 0 ? (e < n_ev) : (e >= 0);
            if (have) cur = load_event(ev, e);
            continue;
        }
        const int pe = have ? (int)cur.x : (dir > 0 ? c.len : -1);
        int run = dir > 0 ? pe - i : i - pe;
        if (run > s - k) run = s - k;
        if (run > 0) {                                             // plain reference bases up to the next event
            const int m = run < 8 ? run : 8;
            emit_group(E, fetch_codes(c, C, i, dir, m), m);
            i += dir * m; k += m;
            continue;
        }
        if (!have) { ok = false; break; }                         // walked off the contig
        const uint32_t t = cur.y & 3u, n = cur.y >> 5;
        uint32_t base = (cur.y >> 2) & 7u;
        base = base < 4 ? base ^ comp : 4u;
        if (t == kEvInsert) {
            ++w.n_indel; ++w.n_indel_first;
            if (!strand) { emit_group(E, base, 1); ++k; }
            const int m = (int)n < s - k ? (int)n : s - k;
            if (strand) { ext += m; ins_ref_pending = true; }
            ins_left = m; ins_at = 0;
            if (m > 0) continue;
            // nothing of the insertion fits (the read ended on the reference base, plus strand): fall through
            ins_ref_pending = false;
        } else if (t == kEvDelete) {
            ++w.n_indel;
            if (++strayed > kNCheckSlack - 2) C.has_n = true;
            if (strand && --ext < 0) { ok = false; break; }
        } else {
            emit_group(E, base, 1); ++k;
            if (t == kEvSubst) ++w.n_sub;
        }
        i += dir; e += dir;
        have = dir > 0 ? (e < n_ev) : (e >= 0);
        if (have) cur = load_event(ev, e);
    }
    w.ext = ext;
    return ok;
}

// ---- Ion Torrent flow model, one thread per read, on a nibble-packed row (thread-per-pair kernel) ---------------
// Same algorithm and draw order as flow_errors() above (src/dwgsim.c:246-417); the read lives in the thread's row of
// the shared staging tile (8 symbols per word), so insertions / deletions are word-wise funnel shifts.

__device__ __forceinline__ uint32_t nib_get(const uint32_t *r, int k) { return (r[k >> 3] >> ((k & 7) << 2)) & 15u; }
__device__ __forceinline__ void nib_set(uint32_t *r, int k, uint32_t v)
{
    const int sh = (k & 7) << 2;
    r[k >> 3] = (r[k >> 3] & ~(15u << sh)) | (v << sh);
}
__device__ __forceinline__ uint32_t nib_mask_ge(int k) { return (k & 7) ? ~0u << ((k & 7) << 2) : ~0u; }   // nibbles >= k&7 of k's word
// the draw sources of flow_model.h's FlowCoin for one (pair, attempt, end): sequential words of the FLOW stream (gaps of the
// error coin) and of the FLOWU stream (uniforms of the events)
struct FlowDraw {
    PairKey key;
    uint32_t end, next, unext;
    __device__ __forceinline__ uint32_t gap_word()
    {
        const uint4 b = draw_block(key, kStFlow, end, next >> 2);
        return word_of(b, next++ & 3u);
    }
    __device__ __forceinline__ uint32_t unif_word()
    {
        const uint4 b = draw_block(key, kStFlowU, end, unext >> 2);
        return word_of(b, unext++ & 3u);
    }
};
constexpr int kFlowGapsAhead = 16;     // gaps every lane draws (four Philox blocks, sixteen table searches) before its read starts

// Substitution errors (__gen_errors_mismatches, src/dwgsim.c:233-244) by thinning (DESIGN.md "RNG addressing"), applied to
// the read's staged row after the walk: candidate n of a read sits 1 + rank(gap table, x_n) symbols after candidate
// n - 1 and is kept when the symbol is not N and y_n is below the cycle's acceptance threshold.  Running this after the
// walk lets the lanes of a warp draw their candidates in lockstep (one Philox block per candidate) instead of each
// lane drawing inside its own group of the walk loop.
__device__ __forceinline__ void apply_errors(uint32_t *row, int s, const TpTables &T, const PairKey &key, int end, int &n_err, int &err_first)
{
    const uint32_t *gap = T.gap[end], *accp = T.acc[end];
    const uint16_t *gguide = T.gap_guide[end];
    int pos = -1;
    n_err = 0; err_first = 0;
    for (uint32_t cand = 0;; ++cand) {
        const uint4 cur = draw_block(key, kStErr, (uint32_t)end, cand);
        pos += 1 + guided_rank(gap, gguide, cur.x);
        if (pos >= s) break;
        const uint32_t c = nib_get(row, pos);
        if (c < 4 && cur.y < accp[pos]) {
            nib_set(row, pos, (c + 1u + __umulhi(cur.z, 3u)) & 3u);
            ++n_err;
            if (pos == 0) err_first = 1;
        }
    }
}

constexpr int kTpThreads = 128;
constexpr int kTpWarps = kTpThreads / 32;
#ifndef DWG_ISIZE_SMEM_MAX
#define DWG_ISIZE_SMEM_MAX 8192
#endif
constexpr int kIsizeSmemMax = DWG_ISIZE_SMEM_MAX;  // insert-size CDFs up to this many entries are copied to shared memory
#ifndef DWG_TP_MIN_BLOCKS
#define DWG_TP_MIN_BLOCKS 6
#endif
constexpr int kTpMinBlocks = DWG_TP_MIN_BLOCKS;   // 5: <= 102 registers per thread, 20 warps per SM

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the sectors of the packed reference that the window of a read of s bases starting at `start` will touch
__device__ __forceinline__ void prefetch_read(const ContigView &c, int start, int strand, int s)
{
    int lo = strand ? start - s - kWindowSlack : start, hi = strand ? start + 32 : start + s + kWindowSlack;
    lo = lo < 0 ? 0 : lo; hi = hi >= c.len ? c.len - 1 : hi;
    for (int w = lo >> 4; w <= (hi >> 4); w += 8) prefetch_l2(c.ref2 + w);      // one per 32-byte sector
    prefetch_l2(c.ref2 + (hi >> 4));
}

// Job lists of the simulate passes (DESIGN.md "Kernels"): an attempt that is rejected (N filter, contig end, -x miss)
// and a pair whose gate draw says "random" are not finished by the lane that found out -- 31 finished lanes would wait
// for it -- but appended to a list in HBM and executed by the next pass with full warps:
//   pass 0  one attempt 0 at every pair of the batch          -> lists F1 (retry, attempt 1) and R1 (random pairs)
//   pass 1  F1 and R1; every lane loops until its pair is done (2 % of the retries are rejected again)
// The draws of a pair are addressed by (pair, attempt), so the result does not depend on which pass executes a job.
struct JobLists {
    uint2 *retry, *random;             // (pair index in the batch, attempt | failed << 16)
    unsigned long long *count;         // [2]: |F1|, |R1|
};
// append the items of the lanes with `push` set (warp-converged call): one atomic per warp
__device__ __forceinline__ void list_push(uint2 *list, unsigned long long *count, bool push, uint2 item, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, push);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (push) list[base + __popc(m & ((1u << lane) - 1u))] = item;
}

template <bool kIon, int kTables /* = SimParams.tp_tables: which sampling tables are copied to shared memory */>
__global__ void __launch_bounds__(kTpThreads, kTpMinBlocks)
simulate_pairs_tp_kernel(const SimParams P, const uint8_t *__restrict__ blob, int64_t first, int64_t gidx_origin, int n, int pass,
                         JobLists J, PairRec *__restrict__ recs, uint32_t *__restrict__ seqw, unsigned long long *__restrict__ status,
                         uint32_t *__restrict__ flow_scratch /* Ion Torrent: [warps of the grid][max(nw0, nw1)][32] words */)
{
    // the jobs of this pass: fresh pairs, or the two lists of the previous pass back to back
    const int n_f = pass == 0 ? 0 : (int)J.count[0], n_r = pass == 0 ? 0 : (int)J.count[1];
    const int n_jobs = pass == 0 ? n : n_f + n_r;
    if ((int)(blockIdx.x * kTpThreads) >= n_jobs) return;
    // staging tile: one row of nw0+nw1 words per thread, odd row stride => conflict-free; every warp flushes its own 32
    // rows to HBM (pair-major) with coalesced stores
    extern __shared__ __align__(16) uint32_t tile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NW = P.nw[0] + P.nw[1], RS = P.row_stride;
    uint32_t *row = tile + (size_t)threadIdx.x * RS;
    uint32_t *wtile = tile + (size_t)warp * 32 * RS;
    // sampling tables behind the tile: insert-size CDF + guide, per end: gap CDF (len entries) + guide, accept thresholds
    // then the reference window, [slot][thread] x 8 bytes
    uint2 *const win_mem = reinterpret_cast<uint2 *>(tile + (((size_t)kTpThreads * RS + 3) & ~(size_t)3));
    TpTables T;
    {
        uint32_t *p32 = reinterpret_cast<uint32_t *>(win_mem + (size_t)P.win_slots * kTpThreads);
        // (kTables: which of them are copied; the others are read where they lie, a few L1 / L2 hits per pair, when that
        //  lets one more CTA fit the SM.  A template parameter: the loads keep their address space)
        const bool isz_smem = P.isize_n <= kIsizeSmemMax && kTables >= 1;  // wider insert-size tables stay in HBM / L2 (one lookup per pair)
        constexpr bool err_smem = !kIon && kTables >= 2, guide_smem = kTables >= 3;
        uint32_t *isz = p32; p32 += isz_smem ? ((P.isize_n + 1) & ~1) : 0;
        uint32_t *gp[2], *ac[2];
        // (the substitution-error tables are not used by the Ion Torrent flow model: no room taken for them there)
        for (int e = 0; e < 2; ++e) { gp[e] = p32; p32 += err_smem ? (P.len[e] + 1) & ~1 : 0; ac[e] = p32; p32 += err_smem ? (P.len[e] + 1) & ~1 : 0; }
        uint16_t *p16 = reinterpret_cast<uint16_t *>(p32);
        uint16_t *ig = p16; p16 += guide_smem ? 1026 : 0;
        uint16_t *gg[2] = {p16, kIon ? p16 : p16 + 1026};
        if (guide_smem && !kIon) p16 += 2 * 1026;
        if (isz_smem) for (int j = threadIdx.x; j < P.isize_n; j += kTpThreads) isz[j] = P.isize_cdf[j];
        if (guide_smem) for (int j = threadIdx.x; j < 1025; j += kTpThreads) ig[j] = P.isize_guide[j];
        if (!kIon)
            for (int e = 0; e < 2; ++e) {
                if (err_smem) for (int j = threadIdx.x; j < P.len[e]; j += kTpThreads) { gp[e][j] = P.err_gap[e][j]; ac[e][j] = P.err_acc[e][j]; }
                if (guide_smem) for (int j = threadIdx.x; j < 1025; j += kTpThreads) gg[e][j] = P.gap_guide[e][j];
            }
        T.isize_cdf = isz_smem ? isz : P.isize_cdf; T.isize_guide = guide_smem ? ig : P.isize_guide;
        for (int e = 0; e < 2; ++e) {
            T.gap[e] = err_smem ? gp[e] : P.err_gap[e]; T.acc[e] = err_smem ? ac[e] : P.err_acc[e];
            T.gap_guide[e] = guide_smem ? gg[e] : P.gap_guide[e];
        }
        // Ion Torrent: flow order (codes) and one flow-mask bit vector per thread
        T.flow_order = nullptr; T.flow_mask = nullptr; T.flow_nd = nullptr; T.flow_q = nullptr;
        if (kIon) {
            int8_t *fo = reinterpret_cast<int8_t *>(p16);           // (behind the insert-size guide)
            for (int j = threadIdx.x; j < P.flow_order_len; j += kTpThreads) fo[j] = P.flow_order[j];
            T.flow_order = fo;
            uint32_t *masks = reinterpret_cast<uint32_t *>(fo + ((P.flow_order_len + 15) & ~15));
            T.flow_mask = masks + (size_t)threadIdx.x * ((P.flow_order_len + 31) >> 5);
            uint16_t *nd = reinterpret_cast<uint16_t *>(masks + (size_t)kTpThreads * ((P.flow_order_len + 31) >> 5));
            for (int j = threadIdx.x; j < 4 * P.flow_order_len; j += kTpThreads) nd[j] = (uint16_t)fm_build_nd_entry(P.flow_order, P.flow_order_len, j >> 2, j & 3);
            T.flow_nd = nd;
            T.flow_q = nd + ((4 * P.flow_order_len + 7) & ~7) + (size_t)threadIdx.x * kFlowGapsAhead;
        }
        __syncthreads();                                                 // the only CTA-wide barrier
    }
    constexpr bool ion = kIon;
    // Ion Torrent rounds are long (the flow model runs per read): a second pass of a few warps costs more than repeating
    // a walk inside the lane, so there every lane finishes its own pair
    const bool defer = pass == 0 && !kIon;
    const int solid = P.data_type == 1;
    const int s0 = P.len[0], s1 = P.len[1];
    uint32_t *dst0 = row, *dst1 = row + P.nw[0];
    // Ion Torrent: the rows live in HBM / L2 instead, word w of lane l at warp_rows[w * 32 + l] (flow_model.h streams every read
    // through its row and a scratch row; in shared memory the two would leave room for eight warps per SM)
    constexpr int kRowStride = kIon ? 32 : 1;
    const int max_nw = P.nw[0] > P.nw[1] ? P.nw[0] : P.nw[1];
    uint32_t *const warp_rows = kIon ? flow_scratch + (size_t)(blockIdx.x * kTpWarps + warp) * (NW + max_nw) * 32 : nullptr;
    if (kIon) { dst0 = warp_rows + lane; dst1 = dst0 + P.nw[0] * 32; }
    const uint32_t ring_base = smem_addr(win_mem) + threadIdx.x * 8;
    unsigned failed_total = 0;

    for (int jbase = (blockIdx.x * kTpWarps + warp) * 32; jbase < n_jobs; jbase += gridDim.x * kTpThreads) {
        const int job = jbase + lane;
        int p = -1, kind = 0;                                            // kind 0: an attempt, 1: a random pair
        uint32_t attempt = 0, failed_flag = 0;
        if (job < n_jobs) {
            if (pass == 0) p = job;
            else {
                const uint2 it = job < n_f ? J.retry[job] : J.random[job - n_f];
                kind = job < n_f ? 0 : 1;
                p = (int)it.x; attempt = it.y & 0xFFFFu; failed_flag = it.y >> 16;
            }
        }
        int staged = -1;
        bool push_retry = false, push_random = false, walked = false;    // walked: both ends of an accepted attempt are staged
        Walk w0, w1;
        int strand0 = 0, strand1 = 0, hap = 0;
        w0.ext = w1.ext = 0; w0.n_sub = w0.n_indel = w0.n_indel_first = w1.n_sub = w1.n_indel = w1.n_indel_first = 0;
        while (p >= 0) {                                                 // one trip when retries are deferred
            const int64_t q = first + p;
            const uint64_t gidx = (uint64_t)(gidx_origin + q);
            PairKey key{P.seed, (uint32_t)gidx, (uint32_t)(gidx >> 32), attempt};
            if (kind == 1) {
                PairRec rec;
                rec.attempt = (uint16_t)attempt;
                rec.n_err_first = 0;                                             // random pair, src/dwgsim.c:983-1001
                rec.flags = (uint8_t)(kRecRandom | (failed_flag ? kRecFailed : 0));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int s = j ? s1 : s0;
                    rec.pos[j] = 0; rec.len[j] = (uint16_t)s;
                    rec.n_err[j] = rec.n_sub[j] = rec.n_indel[j] = rec.n_indel_first[j] = 0;
                    if (s <= 0) continue;
                    EmitT<kRowStride> E;
                    emit_begin(E, j ? dst1 : dst0, solid);
                    for (int k = 0; k < s; k += 64) {
                        const uint4 blk = draw_block(key, kStRandBase, j, (uint32_t)(k >> 6));
#pragma unroll
                        for (int wd = 0; wd < 4; ++wd) {
                            const uint32_t word = word_of(blk, wd);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const int k0 = k + wd * 16 + half * 8;
                                if (k0 >= s) break;
                                const int m = s - k0 < 8 ? s - k0 : 8;
                                uint32_t x = (word >> (16 * half)) & ((1u << (2 * m)) - 1u);
                                x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u;
                                emit_group(E, x, m);
                            }
                        }
                    }
                    emit_end(E);
                }
                recs[p] = rec;
                staged = p;
                break;
            }
            const uint4 b0 = draw_block(key, kStPair, 0, 0);
            if ((uint64_t)b0.x < P.thr_genomic) {                             // src/dwgsim.c:649
                kind = 1;
                if (defer) { push_random = true; break; }
                continue;
            }
            int contig_index;
            const ContigDesc *cd = find_contig(blob, q, &contig_index);
            const ContigView cv = view_of(blob, cd);
            int d, pos;
            bool ok = true;
            if (P.amplicons) { pos = 0; d = cv.len; }
            else {
                const int slen = P.regions ? region_sample_len(blob, q) : cv.len;
                if (s1 > 0) {
                    // (the guide holds 16-bit ranks: tables past 65,535 steps, -s above about 4,000, are searched whole)
                    d = P.isize_lo + (P.isize_n <= 65535 ? guided_rank(T.isize_cdf, T.isize_guide, b0.y) : table_rank(P.isize_cdf, P.isize_n, b0.y));
                    const int min_dist = s0 + s1;
                    if (d < min_dist) d = min_dist;
                    if (d > slen) d = slen;
                } else d = 0;
                const uint64_t range = (uint64_t)((int64_t)slen - d + 1);
                pos = (int)__umul64hi(range, ((uint64_t)b0.z << 32) | b0.w);
                if (P.regions && (pos = map_to_regions(blob, q, pos, d)) < 0) ok = false;
            }
            EmitT<kRowStride> E0, E1;
            if (ok) {
                const uint4 b1 = draw_block(key, kStPair, 0, 1);
                hap = ((uint64_t)b1.x < P.thr_hap0) ? 0 : 1;
                strand0 = P.read_one_strand == 0 ? ((b1.y >> 31) ? 0 : 1) : (P.read_one_strand == 1 ? 0 : 1);
                if (P.strandedness == 0) strand1 = (P.data_type == 0) ? 1 - strand0 : strand0;
                else strand1 = (P.strandedness == 1) ? strand0 : 1 - strand0;
                int st0, st1 = 0;
                const int last = cv.len - 1;
                if (s1 > 0) {                                             // src/dwgsim.c:745-810
                    if (strand0 == strand1) {
                        if (strand0 == 0) { st0 = P.amplicons ? last : (P.is_inner ? pos + s1 + d - 1 : pos + d - s0); st1 = pos; }
                        else { st0 = pos + s0 - 1; st1 = P.amplicons ? last : (P.is_inner ? pos + s0 + d + s1 - 1 : pos + d - 1); }
                    } else if (strand0 == 0) { st0 = pos; st1 = P.amplicons ? last : (P.is_inner ? pos + s0 + d + s1 - 1 : pos + d - 1); }
                    else { st0 = P.amplicons ? last : (P.is_inner ? pos + s1 + d + s0 - 1 : pos + d - 1); st1 = pos; }
                } else st0 = strand0 == 0 ? pos : (P.amplicons ? last : pos + s0 - 1);
                // start the memory accesses of both ends before walking the first one
                RefWindow R;
                R.base = ring_base; R.slots = 0; R.first = 0; R.fresh = false;
                int hint1 = -1;
                uint32_t any1 = 0;
                const bool in0 = st0 >= 0 && st0 < cv.len, in1 = s1 > 0 && st1 >= 0 && st1 < cv.len;
                if (in0) window_prime(cv, R, st0, strand0 ? -1 : 1, s0, P.win_slots);
                if (in1) prefetch_read(cv, st1, strand1, s1);                // end 1's window is primed after walk 0: have it in L2 by then
                if (s1 > 0) { hint1 = walk_hint(cv, hap, st1, strand1); if (in1) any1 = n_precheck(cv, st1, strand1 ? -1 : 1, s1); }
                const int hint0 = walk_hint(cv, hap, st0, strand0);
                const uint32_t any0 = in0 ? n_precheck(cv, st0, strand0 ? -1 : 1, s0) : 0u;
                R.has_n = any0 != 0;
                emit_begin(E0, dst0, solid);
                ok = walk_thread(cv, hap, st0, strand0, s0, E0, w0, hint0, R);
                if (ok) { emit_end(E0); ok = E0.nN <= P.max_n; }
                if (s1 > 0) {
                    bool ok1 = false;
                    if (ok) {                                              // a rejected end 0 already rejects the pair
                        emit_begin(E1, dst1, solid);
                        R.slots = 0; R.fresh = false;
                        if (in1) window_prime(cv, R, st1, strand1 ? -1 : 1, s1, P.win_slots);
                        R.has_n = any1 != 0;
                        ok1 = walk_thread(cv, hap, st1, strand1, s1, E1, w1, hint1, R);
                        if (ok1) { emit_end(E1); ok1 = E1.nN <= P.max_n; }
                    }
                    ok = ok && ok1;
                } else { w1.ext = 0; w1.n_sub = w1.n_indel = w1.n_indel_first = 0; }
            }
            if (!ok) {                                                    // src/dwgsim.c:833-842
                ++failed_total;
                if (attempt >= (uint32_t)kMaxTrials) {                    // 10001 rejected attempts: the host reports it
                    atomicOr(status, 1ull);
                    kind = 1; failed_flag = 1;
                    if (defer) { push_random = true; break; }
                    continue;
                }
                ++attempt;
                if (defer) { push_retry = true; break; }
                continue;
            }
            walked = true;
            break;
        }
        // the accepted attempt's epilogue runs after the loop, where the lanes of the warp are together again (inside it,
        // a lane that retries would send the others through the error models a second time)
        if (kIon) __syncwarp();                                          // (without it the lanes arrive here in the groups they left the loop in)
        if (walked) {
            const uint64_t gidx = (uint64_t)(gidx_origin + first + p);
            const PairKey key{P.seed, (uint32_t)gidx, (uint32_t)(gidx >> 32), attempt};
            PairRec rec;
            rec.attempt = (uint16_t)attempt;
            int n_err0 = 0, n_err1 = 0, err_first0 = 0, err_first1 = 0;
            if (!ion) {                                                   // src/dwgsim.c:866-881
                apply_errors(dst0, s0, T, key, 0, n_err0, err_first0);
                if (s1 > 0) apply_errors(dst1, s1, T, key, 1, n_err1, err_first1);
            }
            rec.flags = (uint8_t)((strand0 ? kRecStrand0 : 0) | (strand1 ? kRecStrand1 : 0) | (hap ? kRecHap1 : 0));
            rec.pos[0] = (uint32_t)(w0.ext + 1); rec.pos[1] = (uint32_t)(w1.ext + 1);
            rec.len[0] = (uint16_t)s0; rec.len[1] = (uint16_t)s1;
            rec.n_sub[0] = (uint16_t)w0.n_sub; rec.n_sub[1] = (uint16_t)w1.n_sub;
            rec.n_indel[0] = (uint16_t)w0.n_indel; rec.n_indel[1] = (uint16_t)w1.n_indel;
            rec.n_indel_first[0] = (uint16_t)w0.n_indel_first; rec.n_indel_first[1] = (uint16_t)w1.n_indel_first;
            rec.n_err[0] = (uint16_t)n_err0; rec.n_err[1] = (uint16_t)(s1 > 0 ? n_err1 : 0);
            rec.n_err_first = (uint8_t)((err_first0 ? 1 : 0) | ((s1 > 0 && err_first1) ? 2 : 0));
            if constexpr (kIon) {                                       // flow-space errors, src/dwgsim.c:861-864
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    const int s = j ? s1 : s0;
                    if (s <= 0) continue;
                    int nerr = 0, ovf = 0;
                    // the first gaps of the error coin, drawn while the lanes of the warp are together (a draw inside the
                    // model runs for one lane at a time: a Philox block and a 12-step table search each)
                    for (int b = 0; b < kFlowGapsAhead / 4; ++b) {
                        const uint4 blk = draw_block(key, kStFlow, (uint32_t)j, (uint32_t)b);
                        T.flow_q[4 * b + 0] = (uint16_t)table_rank(P.flow_gap[j], kFlowGapN, blk.x);
                        T.flow_q[4 * b + 1] = (uint16_t)table_rank(P.flow_gap[j], kFlowGapN, blk.y);
                        T.flow_q[4 * b + 2] = (uint16_t)table_rank(P.flow_gap[j], kFlowGapN, blk.z);
                        T.flow_q[4 * b + 3] = (uint16_t)table_rank(P.flow_gap[j], kFlowGapN, blk.w);
                    }
                    FlowCoin<FlowDraw> rng(FlowDraw{key, (uint32_t)j, (uint32_t)kFlowGapsAhead, 0u}, P.flow_gap[j], T.flow_q, kFlowGapsAhead);
                    // the read streams between its row in shared memory and a scratch row in HBM / L2 whose words are interleaved
                    // with those of the warp's other lanes (coalesced: the lanes advance through their reads together)
                    const FlowRow ra{j ? dst1 : dst0, 32}, rb{warp_rows + NW * 32 + lane, 32};
                    const int nl = flow_model_rows(ra, rb, s, P.cap[j], j ? strand1 : strand0, T.flow_order,
                                                   P.flow_order_len, T.flow_nd, T.flow_mask, rng, &nerr, &ovf);
                    if (ovf) atomicOr(status, 2ull);
                    rec.len[j] = (uint16_t)(nl > 0 ? nl : 0);
                    rec.n_err[j] = (uint16_t)nerr;
                }
            }
            recs[p] = rec;
            staged = p;
        }
        __syncwarp();
        if (defer) {
            const uint2 item = make_uint2((uint32_t)p, attempt | (failed_flag << 16));
            list_push(J.retry, J.count, push_retry, item, lane);
            list_push(J.random, J.count + 1, push_random, item, lane);
        }
        // flush the warp's staged rows (pair-major in HBM, NW words per pair).  The rows of a pass-0 round are contiguous in
        // HBM; without row padding they are contiguous in shared memory too and leave as 16-byte vectors (rows of deferred
        // pairs go along and are overwritten by pass 1)
        if (kIon) {
            // the rows of the warp, interleaved in HBM / L2 -> pair-major rows of seqw (strided reads, coalesced writes)
            for (int x = lane; x < 32 * NW; x += 32) {
                const int r = (int)__umulhi((uint32_t)x, P.inv_nw), dp = __shfl_sync(0xffffffffu, staged, r);
                if (dp >= 0) seqw[(size_t)dp * NW + (x - r * NW)] = warp_rows[(x - r * NW) * 32 + r];
            }
        } else if (RS == NW && pass == 0 && jbase + 32 <= n_jobs) {
            const uint4 *src = reinterpret_cast<const uint4 *>(wtile);
            uint4 *dst = reinterpret_cast<uint4 *>(seqw + (size_t)jbase * NW);
            for (int x = lane; x < 8 * NW; x += 32) dst[x] = src[x];
        } else {
            // tile index of linear word x is x + (x / NW) * (RS - NW); x / NW by a multiply-high with P.inv_nw = 2^32/NW + 1
            for (int x = lane; x < 32 * NW; x += 32) {
                const int r = (int)__umulhi((uint32_t)x, P.inv_nw), dp = __shfl_sync(0xffffffffu, staged, r);
                if (dp >= 0) seqw[(size_t)dp * NW + (x - r * NW)] = wtile[x + r * (RS - NW)];
            }
        }
        __syncwarp();
    }
    if (failed_total) atomicAdd(status + 1, (unsigned long long)failed_total);
}

// the result words of a batch to mapped host memory in one launch: totals[0..3] -> dst[0..3], status[0..1] -> dst[8..9]
// `queue` (batches queued without a host sync in between, dwgsim_gpu_resident_enqueue): [0] the running count of random
// pairs, advanced by this batch when `advance`; [1] the error bits of all batches since the last wait -> dst[10]
__global__ void publish_batch_kernel(unsigned long long *__restrict__ dst_host, const unsigned long long *__restrict__ totals,
                                     const unsigned long long *__restrict__ status, unsigned long long *__restrict__ queue, int advance)
{
    if (threadIdx.x < 4) dst_host[threadIdx.x] = totals[threadIdx.x];
    else if (threadIdx.x < 6) dst_host[8 + threadIdx.x - 4] = status[threadIdx.x - 4];
    else if (threadIdx.x == 6 && queue) {
        if (advance) queue[0] += totals[0];
        const unsigned long long bits = queue[1] | status[0];
        queue[1] = bits;
        dst_host[10] = bits;
    }
    __threadfence_system();
}
// sharded runs: from the all-gathered random-pair counts of a round (counts[world], rank order = batch order) the count before
// this rank's batch -> queue[2], and the running count of the whole job queue[0] += all of them
__global__ void gathered_base_kernel(const unsigned long long *__restrict__ counts, int world, int rank, unsigned long long *__restrict__ queue)
{
    unsigned long long before = 0, all = 0;
    for (int r = threadIdx.x; r < world; r += 32) { const unsigned long long c = counts[r]; all += c; if (r < rank) before += c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    if (threadIdx.x == 0) { queue[2] = queue[0] + before; queue[0] += all; }
}
// a few 64-bit results to mapped host memory (no copy engine involved)
__global__ void publish_words_kernel(unsigned long long *__restrict__ dst_host, const unsigned long long *__restrict__ src, int n)
{
    if ((int)threadIdx.x < n) dst_host[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// ---- record geometry shared by the layout and format kernels --------------------------------------------
__device__ __forceinline__ int ndigits10(uint32_t v)
{
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
         : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
__device__ __forceinline__ int ndigits16(uint64_t v) { return v ? (67 - __clzll((long long)v)) >> 2 : 1; }

// the 13 numeric fields of a read name (src/dwgsim.c:923-929); variant 1 = SOLiD bwa counts (:945-946)
__device__ __forceinline__ uint64_t name_field(const PairRec &r, uint64_t serial, int f, int variant)
{
    const bool rnd = r.flags & kRecRandom;
    switch (f) {
        case 0: return r.pos[0];
        case 1: return r.pos[1];
        case 2: return (r.flags & kRecStrand0) ? 1 : 0;
        case 3: return (r.flags & kRecStrand1) ? 1 : 0;
        case 4: case 5: return rnd ? 1 : 0;
        case 6: return r.n_err[0] - (variant ? (r.n_err_first & 1) : 0);
        case 7: return r.n_sub[0];
        case 8: return r.n_indel[0] - (variant ? r.n_indel_first[0] : 0);
        case 9: return r.n_err[1] - (variant ? ((r.n_err_first >> 1) & 1) : 0);
        case 10: return r.n_sub[1];
        case 11: return r.n_indel[1] - (variant ? r.n_indel_first[1] : 0);
        default: return serial;
    }
}
// '@' + prefix + contig + 13 x (separator + digits)
__device__ __forceinline__ int name_length(const SimParams &P, const PairRec &r, uint64_t serial, int contig_name_len,
                                           int variant)
{
    int n = 1 + P.prefix_len + ((r.flags & kRecRandom) ? 4 : contig_name_len) + 13;
#pragma unroll
    for (int f = 0; f < 12; ++f) n += ndigits10((uint32_t)name_field(r, serial, f, variant));
    return n + ndigits16(serial);
}
// bytes of the records of one pair in the three streams (src/dwgsim.c:920-980)
__device__ __forceinline__ void record_lengths(const SimParams &P, const PairRec &r, uint64_t serial, int contig_name_len,
                                               uint32_t len[3])
{
    const bool solid = P.data_type == 1;
    const int name_full = name_length(P, r, serial, contig_name_len, 0);
    const int name_bwa = solid ? name_length(P, r, serial, contig_name_len, 1) : name_full;
    len[0] = len[1] = len[2] = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int L = r.len[j];
        if (L <= 0) continue;
        if (P.out_bwa) len[j] = (uint32_t)(name_bwa + 3 + (solid ? 2 * (L - 1) : 2 * L) + 4);
        if (P.out_bfast) len[2] += (uint32_t)(name_full + 1 + 2 * L + 4 + (solid ? 1 : 0));
    }
}

// '@' prefix contig _pos1_pos2_s1_s2_r_r_e:s:i_e:s:i_hex   (src/dwgsim.c:923-929), assembled in a 64-bit register and stored
// eight bytes at a time straight into the name row (16-byte aligned, name_cap bytes): no byte stores, no local staging
// buffer.  Values handed to put() carry their first byte lowest and nothing above their n bytes.
struct NameEmit {
    unsigned long long acc;
    int o;
    char *dst;
    __device__ __forceinline__ void put(unsigned long long v, int n)        // 1 <= n <= 8
    {
        const int k = o & 7;
        acc |= v << (8 * k);
        if (k + n >= 8) {
            *reinterpret_cast<unsigned long long *>(dst + (o - k)) = acc;
            acc = k ? v >> (8 * (8 - k)) : 0ull;
        }
        o += n;
    }
    __device__ __forceinline__ void finish() { if (o & 7) *reinterpret_cast<unsigned long long *>(dst + (o & ~7)) = acc; }
};
// the decimal digits of v < 10^8 as ASCII bytes, first digit lowest; *nd = how many (`fixed`: exactly 8, zero padded)
__device__ __forceinline__ unsigned long long dec_bytes(uint32_t v, int *nd, bool fixed = false)
{
    unsigned long long d = 0;
    int n = 0;
    do { d = (d << 8) | (unsigned long long)('0' + v % 10u); v /= 10u; ++n; } while (fixed ? n < 8 : v != 0u);
    *nd = n;
    return d;
}
// separator + decimal number
__device__ __forceinline__ void put_field(NameEmit &E, uint32_t sep, uint32_t v)
{
    if (v < 10u) { E.put(sep | ((unsigned long long)('0' + v) << 8), 2); return; }     // strands, flags and most counts
    int nd;
    if (v >= 100000000u) {
        const unsigned long long hi = dec_bytes(v / 100000000u, &nd);
        E.put(sep | (hi << 8), nd + 1);
        E.put(dec_bytes(v % 100000000u, &nd, true), 8);
        return;
    }
    const unsigned long long d = dec_bytes(v, &nd);
    if (nd < 8) E.put(sep | (d << 8), nd + 1);
    else { E.put(sep, 1); E.put(d, 8); }
}
__device__ __forceinline__ unsigned long long hex_bytes(uint32_t v, int nd)      // nd <= 8 hexadecimal digits of v, first digit lowest
{
    unsigned long long d = 0;
    for (int j = 0; j < nd; ++j) { const uint32_t h = v & 15u; d = (d << 8) | (unsigned long long)('0' + h + (h > 9u ? 39u : 0u)); v >>= 4; }
    return d;
}
__device__ __forceinline__ int write_name_words(const SimParams &P, const PairRec &r, uint64_t serial, const char *cname,
                                                int cname_len, int variant, char *row)
{
    NameEmit E;
    E.acc = 0; E.o = 0; E.dst = row;
    E.put('@', 1);
    for (int j = 0; j < P.prefix_len; ++j) E.put((unsigned long long)(uint8_t)P.prefix[j], 1);
    if (r.flags & kRecRandom) E.put(0x646e6172ull, 4);                          // "rand"
    else for (int j = 0; j < cname_len; ++j) E.put((unsigned long long)(uint8_t)cname[j], 1);
#pragma unroll
    for (int f = 0; f < 12; ++f)
        put_field(E, (f == 7 || f == 8 || f == 10 || f == 11) ? ':' : '_', (uint32_t)name_field(r, serial, f, variant));
    const int nd = ndigits16(serial);
    if (nd > 8) {
        E.put('_' | (hex_bytes((uint32_t)(serial >> 32), nd - 8) << 8), nd - 7);
        E.put(hex_bytes((uint32_t)serial, 8), 8);
    } else if (nd == 8) { E.put('_', 1); E.put(hex_bytes((uint32_t)serial, 8), 8); }
    else E.put('_' | (hex_bytes((uint32_t)serial, nd) << 8), nd + 1);
    E.finish();
    return E.o;
}

// ---- layout kernels --------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem_w /* [kWarpsPerBlock] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) smem_w[warp] = inc;
    __syncthreads();
    T base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kWarpsPerBlock; ++i) { T x = smem_w[i]; if (i < warp) base += x; tot += x; }
    *total = tot;
    return base + inc - v;
}

// block sums of the random-pair flags
__global__ void __launch_bounds__(kThreads)
layout_count_random_kernel(const PairRec *__restrict__ recs, int n, unsigned long long *__restrict__ blk_rand)
{
    __shared__ uint32_t sw[kWarpsPerBlock];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t c = 0;
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) if (base + t < n) c += (recs[base + t].flags & kRecRandom) ? 1u : 0u;
    uint32_t tot;
    block_exclusive_scan<uint32_t>(c, &tot, sw);
    if (threadIdx.x == 0) blk_rand[blockIdx.x] = tot;
}

// exclusive scan of m 64-bit block sums per array (m <= a few thousand; one CTA per array), in place; totals[a] = sum
__global__ void __launch_bounds__(1024)
layout_scan_blocks_kernel(unsigned long long *__restrict__ v, int m, int n_arrays, unsigned long long *__restrict__ totals)
{
    __shared__ unsigned long long sw[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (m + 1023) / 1024;                                // consecutive elements per thread: one pass over the array
    for (int a = blockIdx.x; a < n_arrays; a += gridDim.x) {          // one CTA per array
        unsigned long long *x = v + (size_t)a * m;
        const int i0 = threadIdx.x * per, i1 = min(i0 + per, m);
        unsigned long long mine = 0;
        for (int i = i0; i < i1; ++i) mine += x[i];
        unsigned long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        if (warp == 0) {                                              // scan of the 32 warp sums
            unsigned long long w = sw[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            sw[lane] = winc - w;
            if (lane == 31) totals[a] = winc;
        }
        __syncthreads();
        unsigned long long run = sw[warp] + inc - mine;               // exclusive prefix of this thread's first element
        for (int i = i0; i < i1; ++i) { const unsigned long long t = x[i]; x[i] = run; run += t; }
        __syncthreads();
    }
}

// serial (ii or rand_ii) + record lengths per pair, block sums of the lengths
#ifndef DWG_LAYOUT_MIN_BLOCKS
#define DWG_LAYOUT_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(kThreads, DWG_LAYOUT_MIN_BLOCKS)
layout_lengths_kernel(const SimParams P, const uint8_t *__restrict__ blob, const PairRec *__restrict__ recs, int n,
                      int64_t first, unsigned long long rand_base, const unsigned long long *__restrict__ rand_base_dev,
                      const unsigned long long *__restrict__ blk_rand_excl,
                      unsigned long long *__restrict__ serial, uint32_t *__restrict__ lens /* [3][n] */,
                      unsigned long long *__restrict__ blk_len /* [3][nblk] */,
                      char *__restrict__ names /* [n][nvar][name_cap] */, uint16_t *__restrict__ name_len /* [n][2] */)
{
    __shared__ uint32_t sw[kWarpsPerBlock];
    __shared__ unsigned long long sw64[kWarpsPerBlock];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t c = 0;
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) if (base + t < n) c += (recs[base + t].flags & kRecRandom) ? 1u : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan<uint32_t>(c, &tot, sw);
    unsigned long long rs = rand_base + (rand_base_dev ? *rand_base_dev : 0ull) + blk_rand_excl[blockIdx.x] + ex;
    unsigned long long sum[3] = {0, 0, 0};
#pragma unroll 1
    for (int t = 0; t < kScanItems; ++t) {
        if (base + t >= n) break;
        const int64_t q = first + base + t;
        int ci;
        const ContigDesc *cd = find_contig(blob, q, &ci);
        const PairRec rt = recs[base + t];                    // (L1 hit: read above for the random-pair count)
        unsigned long long ser;
        if (rt.flags & kRecRandom) ser = rs++;
        else ser = (unsigned long long)(q - cd->pair_base);
        serial[base + t] = ser;
        int nl_v[2] = {0, 0};
        {   // the read name(s), written once here and copied by the format kernel
            const BlobHeader *hd = reinterpret_cast<const BlobHeader *>(blob);
            const char *cname = reinterpret_cast<const char *>(blob + hd->names_off + cd->name_off);
            const int nvar = (P.data_type == 1 && P.out_bwa) ? 2 : 1;
            for (int v = 0; v < nvar; ++v) {
                char *dst = names + ((size_t)(base + t) * nvar + v) * P.name_cap;   // name_cap is a multiple of 16
                nl_v[v] = write_name_words(P, rt, ser, cname, (int)cd->name_len, v, dst);
            }
            if (nvar == 1) nl_v[1] = nl_v[0];
            *reinterpret_cast<uint32_t *>(name_len + (size_t)(base + t) * 2) = (uint32_t)nl_v[0] | ((uint32_t)nl_v[1] << 16);
        }
        // bytes of the pair's records in the three streams (src/dwgsim.c:920-980), from the name lengths just written
        uint32_t len[3] = {0, 0, 0};
        {
            const bool solid = P.data_type == 1;
            const int name_full = nl_v[0], name_bwa = nl_v[1];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int L = rt.len[j];
                if (L <= 0) continue;
                if (P.out_bwa) len[j] = (uint32_t)(name_bwa + 3 + (solid ? 2 * (L - 1) : 2 * L) + 4);
                if (P.out_bfast) len[2] += (uint32_t)(name_full + 1 + 2 * L + 4 + (solid ? 1 : 0));
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { lens[(size_t)k * n + base + t] = len[k]; sum[k] += len[k]; }
    }
    const int nblk = gridDim.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long t2;
        block_exclusive_scan<unsigned long long>(sum[k], &t2, sw64);
        if (threadIdx.x == 0) blk_len[(size_t)k * nblk + blockIdx.x] = t2;
    }
}

// lengths -> byte offsets inside the batch (in place)
__global__ void __launch_bounds__(kThreads)
layout_offsets_kernel(int n, const unsigned long long *__restrict__ blk_len_excl /* [3][nblk] */,
                      uint32_t *__restrict__ lens /* [3][n] in, offsets out (low 32 bits) */)
{
    __shared__ unsigned long long sw64[kWarpsPerBlock];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    const int nblk = gridDim.x;
    for (int k = 0; k < 3; ++k) {
        uint32_t l[kScanItems];
        unsigned long long s = 0;
#pragma unroll
        for (int t = 0; t < kScanItems; ++t) { l[t] = base + t < n ? lens[(size_t)k * n + base + t] : 0u; s += l[t]; }
        unsigned long long tot;
        unsigned long long ex = block_exclusive_scan<unsigned long long>(s, &tot, sw64) + blk_len_excl[(size_t)k * nblk + blockIdx.x];
#pragma unroll
        for (int t = 0; t < kScanItems; ++t) {
            if (base + t < n) lens[(size_t)k * n + base + t] = (uint32_t)ex;   // batches are < 4 GiB per stream
            ex += l[t];
        }
    }
}

// ---- kernel B: format ------------------------------------------------------------------------------------
// Every WARP formats its own mini-tiles of P.tile_pairs consecutive pairs, independently of the other warps of the CTA
// (no CTA barrier after the prologue; the warps only share the read-only sampling tables).  The records of consecutive
// pairs are contiguous in each output stream, so a mini-tile owns ONE contiguous byte range per stream: it is
// assembled in the warp's shared-memory region at the same offset modulo 16 as its destination and copied out with
// aligned 16-byte stores (byte stores only in the two boundary chunks shared with the neighbouring mini-tiles).
//   step 0  lanes 0..np: the pair's record, stream offsets and name lengths, prefetched one mini-tile ahead (the names
//           themselves, src/dwgsim.c:923-929, were written by layout_lengths_kernel)
//   step 1  one lane per (pair, end, 8-base group), consecutive lanes = consecutive groups: 8 bases -> ASCII with two
//           PRMTs, 8 qualities from two Philox blocks (src/dwgsim.c:899-918), written into the bwa and bfast records
//           (src/dwgsim.c:920-980)
//   step 2  one lane per record and 16-byte name chunk; "/1", separators
//   step 3  16-byte copy-out

constexpr int kFmtThreads = 512;
constexpr int kFmtWarps = kFmtThreads / 32;

struct PairMeta {                      // per pair of the warp's mini-tile, in shared memory (32 bytes)
    uint32_t so[3];                    // start of the pair's bytes in each stream's staging area
    uint16_t len[2];                   // read lengths
    uint16_t nfull, nbwa;              // name lengths: bfast (full counts) and bwa variant
    uint32_t attempt;
    uint32_t pad[2];
};

struct FormatSmem {
    int guide_off, cdf_off, qbase_off[2], warp_off, warp_stride, meta_off, stage_off[3], total;   // meta/stage: inside a warp's region
};
__host__ __device__ inline FormatSmem format_smem_layout(const SimParams &P)
{
    FormatSmem L;
    const int WP = P.tile_pairs;
    int o = 0;
    L.guide_off = o; o += 1024 * 8;
    L.cdf_off = o; o += ((P.qdelta_n > 0 && P.qdelta_n <= 512 ? P.qdelta_n : 0) * 4 + 15) & ~15;
    for (int e = 0; e < 2; ++e) { L.qbase_off[e] = o; o += (P.cap[e] + 16) & ~15; }
    L.warp_off = o;
    int w = 0;
    L.meta_off = w; w += WP * (int)sizeof(PairMeta);
    for (int k = 0; k < 3; ++k) { L.stage_off[k] = w; w += (WP * P.rec_cap[k] + 32 + 15) & ~15; }
    L.warp_stride = w;
    L.total = o + (P.fmt_warps > 0 ? P.fmt_warps : kFmtWarps) * w;
    return L;
}

// quality noise: inverse CDF through a 1024-bucket guide.  An entry {t, r} holds the rank r at the bucket's lower bound
// and, when exactly one threshold c lies inside the 2^22-wide bucket, t = c - 1 (no threshold: t = 2^32 - 1): one 8-byte
// shared-memory load and one compare.  r < 0 marks buckets with several thresholds (the tails), t = their number: the
// scan starts from the side of the bucket where the probability mass is (top of a lower-tail bucket, bottom of an upper).
__device__ __noinline__ int qdelta_rank_tail(uint2 ent, const uint32_t *cdf, uint32_t u)
{
    const int r = (int)(ent.y & 0x7fffffffu), top = r + (int)ent.x;
    int j;
    if (!(u >> 31)) { j = top; while (j > r && u < cdf[j - 1]) --j; }
    else { j = r; while (j < top && u >= cdf[j]) ++j; }
    return j;
}

// QUAL draw of base i of an 8-base group: upper 16 bits from field i of the group's first Philox block, lower 16 bits
// from the same field of its second block (field f = half f&1 of word f>>1); oracle: draw_qual_delta
__device__ __forceinline__ uint32_t qual_draw32(const uint4 &b0, const uint4 &b1, int i)
{
    const uint32_t h = word_of(b0, (uint32_t)i >> 1), l = word_of(b1, (uint32_t)i >> 1);
    return (i & 1) ? ((h & 0xFFFF0000u) | (l >> 16)) : ((h << 16) | (l & 0xFFFFu));
}

struct FmtPrefetch {                   // step 0 of a mini-tile, held in registers by lane j for pair j (lane np: end offsets)
    uint32_t lens, tail;               // PairRec words 2 (len[0] | len[1] << 16) and 7 (n_err_first, flags, attempt << 16)
    uint32_t off[3];                   // byte offset of the pair's records in each stream (lane np: end of the mini-tile)
    uint32_t nl;                       // name lengths: nfull | nbwa << 16
};

template <bool kSolid, bool kWrap>
__global__ void __launch_bounds__(kFmtThreads, 2)
format_fastq_kernel(const SimParams P, const uint8_t *__restrict__ blob, int64_t first, int64_t gidx_origin, int n,
                    const PairRec *__restrict__ recs, const uint32_t *__restrict__ seqw,
                    const unsigned long long *__restrict__ serial, const uint32_t *__restrict__ offs /* [3][n] */,
                    const unsigned long long *__restrict__ totals /* bytes of the batch per stream */,
                    const char *__restrict__ gnames, const uint16_t *__restrict__ gname_len,
                    char *__restrict__ out0, char *__restrict__ out1, char *__restrict__ out2)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const FormatSmem L = format_smem_layout(P);
    const int WP = P.tile_pairs, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cdf_in_smem = P.qdelta_n > 0 && P.qdelta_n <= 512;
    const uint32_t *cdf = cdf_in_smem ? reinterpret_cast<const uint32_t *>(smem + L.cdf_off) : P.qdelta_cdf;   // tails only
    {
        uint2 *guide = reinterpret_cast<uint2 *>(smem + L.guide_off);
        for (int j = tid; j < 1024; j += (int)blockDim.x) guide[j] = P.qdelta_n > 0 ? reinterpret_cast<const uint2 *>(P.qguide)[j] : make_uint2(0u, 0u);
        if (cdf_in_smem) for (int j = tid; j < P.qdelta_n; j += (int)blockDim.x) reinterpret_cast<uint32_t *>(smem + L.cdf_off)[j] = P.qdelta_cdf[j];
        for (int j = tid; j < P.cap[0]; j += (int)blockDim.x) smem[L.qbase_off[0] + j] = P.qbase[0][j];
        for (int j = tid; j < P.cap[1]; j += (int)blockDim.x) smem[L.qbase_off[1] + j] = P.qbase[1][j];
    }
    __syncthreads();                                                 // the only CTA-wide barrier
    const uint32_t a_base = in_register(smem_addr(smem));
    const uint32_t a_guide = a_base + L.guide_off, a_qb0 = a_base + L.qbase_off[0], a_qb1 = a_base + L.qbase_off[1];
    const uint32_t a_warp = a_base + L.warp_off + warp * L.warp_stride;
    const uint32_t a_meta = a_warp + L.meta_off;
    const uint32_t a_st0 = a_warp + L.stage_off[0], a_st1 = a_warp + L.stage_off[1], a_st2 = a_warp + L.stage_off[2];
    PairMeta *meta = reinterpret_cast<PairMeta *>(smem + L.warp_off + warp * L.warp_stride + L.meta_off);
    (void)a_meta; (void)blob; (void)serial;

    constexpr bool solid = kSolid;
    constexpr int from = kSolid ? 1 : 0;                            // bwa drops the first colour (src/dwgsim.c:949-953)
    const int nvar = (solid && P.out_bwa) ? 2 : 1;                  // SOLiD bwa names carry reduced counts (:945-946)
    const bool on0 = P.out_bwa != 0, on2 = P.out_bfast != 0;
    const int g0 = (P.cap[0] + 7) >> 3, g1 = (P.cap[1] + 7) >> 3, G = g0 + g1, NW = P.nw[0] + P.nw[1];
    const int ntiles = (n + WP - 1) / WP;
    const int n_warps = (int)blockDim.x >> 5;
    const int tstride = gridDim.x * n_warps;

    auto prefetch = [&](int tile) {
        FmtPrefetch f;
        f.lens = f.tail = 0; f.off[0] = f.off[1] = f.off[2] = 0; f.nl = 0;
        if (tile >= ntiles) return f;
        const int p0 = tile * WP, np = min(WP, n - p0), p = p0 + lane;
        if (lane < np) {
            const uint32_t *r = reinterpret_cast<const uint32_t *>(recs + p);
            f.lens = __ldg(r + 2); f.tail = __ldg(r + 7);
            f.nl = __ldg(reinterpret_cast<const uint32_t *>(gname_len) + p);
        }
        if (lane <= np) {
            if (on0) { f.off[0] = p < n ? __ldg(offs + p) : (uint32_t)totals[0]; f.off[1] = p < n ? __ldg(offs + (size_t)n + p) : (uint32_t)totals[1]; }
            if (on2) f.off[2] = p < n ? __ldg(offs + (size_t)2 * n + p) : (uint32_t)totals[2];
        }
        return f;
    };

    int tile = blockIdx.x * n_warps + warp;
    FmtPrefetch cur = prefetch(tile);
    for (; tile < ntiles; tile += tstride) {
        const FmtPrefetch nxt = prefetch(tile + tstride);
        const int p0 = tile * WP, np = min(WP, n - p0);
        // ---- step 0: geometry of the mini-tile ---------------------------------------------------------------
        uint32_t begin[3], shift[3], total[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            begin[k] = __shfl_sync(0xffffffffu, cur.off[k], 0);
            total[k] = __shfl_sync(0xffffffffu, cur.off[k], np) - begin[k];
            char *outk = k == 0 ? out0 : (k == 1 ? out1 : out2);
            shift[k] = (uint32_t)(reinterpret_cast<uintptr_t>(outk + begin[k]) & 15u);
        }
        if (lane < np) {
            PairMeta m;
            m.len[0] = (uint16_t)(cur.lens & 0xFFFFu); m.len[1] = (uint16_t)(cur.lens >> 16); m.attempt = cur.tail >> 16;
#pragma unroll
            for (int k = 0; k < 3; ++k) m.so[k] = cur.off[k] - begin[k] + shift[k];
            m.nfull = (uint16_t)(cur.nl & 0xFFFFu); m.nbwa = (uint16_t)(cur.nl >> 16);
            m.pad[0] = m.pad[1] = 0;
            meta[lane] = m;
        }
        __syncwarp();
        // ---- step 1: bases and qualities, one lane per (pair, end, 8-base group) ----------------------------------
        for (int it = lane; it < np * G; it += 32) {
            const int t = (int)__umulhi((uint32_t)it, P.inv_groups), gi = it - t * G;
            const int e = gi < g0 ? 0 : 1, g = gi - (e ? g0 : 0);
            const PairMeta &m = meta[t];
            const int Le = m.len[e], k0 = g << 3;
            if (k0 >= Le) continue;
            const int cnt = min(8, Le - k0);
            const uint32_t codes = __ldg(seqw + (size_t)(p0 + t) * NW + (e ? P.nw[0] : 0) + g);
            // qualities, src/dwgsim.c:899-918 (char arithmetic there; emulated with an int8 wrap where it can matter)
            uint32_t q_lo = 0, q_hi = 0;
            if (P.fixed_quality) { q_lo = q_hi = 0x01010101u * (uint32_t)P.fixed_quality; }
            else {
                const uint64_t gidx = (uint64_t)(gidx_origin + first + p0 + t);
                const PairKey key{P.seed, (uint32_t)gidx, (uint32_t)(gidx >> 32), m.attempt};
                uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
                if (P.qdelta_n > 0) { b0 = draw_block(key, kStQual, e, (uint32_t)(2 * g)); b1 = draw_block(key, kStQual, e, (uint32_t)(2 * g + 1)); }
                const uint2 qb8 = lds64((e ? a_qb1 : a_qb0) + k0);           // k0 is a multiple of 8
                int qc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) qc[i] = 33 + (int)__byte_perm(i < 4 ? qb8.x : qb8.y, 0u, 0x4440 + (i & 3));
                if (P.qdelta_n > 0) {
                    uint32_t tails = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                           // straight-line code: eight loads, eight compares
                        const uint32_t u = qual_draw32(b0, b1, i);
                        const uint2 ent = lds64(a_guide + ((u >> 22) << 3));
                        qc[i] += P.qdelta_lo + (int)ent.y + (u > ent.x ? 1 : 0);
                        if (i < cnt) tails |= ent.y;                        // (draws past the end of the read are not used)
                    }
                    if ((int)tails < 0) {                                   // some draw fell into a tail bucket (rare): redo those
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t u = qual_draw32(b0, b1, i);
                            const uint2 ent = lds64(a_guide + ((u >> 22) << 3));
                            if (i < cnt && (int)ent.y < 0) qc[i] += qdelta_rank_tail(ent, cdf, u) - (int)ent.y - (u > ent.x ? 1 : 0);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int v = qc[i];
                    if (kWrap) v = (int)(signed char)(v & 0xFF);
                    v = max(33, min(73, v));
                    if (i < 4) q_lo |= (uint32_t)v << (8 * i); else q_hi |= (uint32_t)v << (8 * (i - 4));
                }
            }
            // 8 nibble codes -> 8 characters: the nibbles are PRMT selectors into "ACGTN" / "01234"
            const uint32_t a_lo = __byte_perm(0x54474341u, 0x0000004Eu, codes & 0xFFFFu), a_hi = __byte_perm(0x54474341u, 0x0000004Eu, codes >> 16);
            uint32_t d_lo = a_lo, d_hi = a_hi;
            if (solid) { d_lo = __byte_perm(0x33323130u, 0x00000034u, codes & 0xFFFFu); d_hi = __byte_perm(0x33323130u, 0x00000034u, codes >> 16); }
            // write into the records: predicated byte stores at fixed offsets from four section pointers
            const int len0 = m.len[0];
            const int rec0 = (on2 && len0 > 0) ? m.nfull + 1 + (solid ? 1 : 0) + 2 * len0 + 4 : 0;
            const uint32_t sb = (e ? a_st1 : a_st0) + m.so[e];
            const uint32_t sf = a_st2 + m.so[2] + (e ? rec0 : 0);
            const int me = Le - from;
            const uint32_t ps_b = sb + m.nbwa + 3 - from + k0, pq_b = ps_b + me + 3;           // bwa: sequence, qualities
            const uint32_t ps_f = sf + m.nfull + 1 + (solid ? 1 : 0) + k0, pq_f = ps_f + Le + 3;   // bfast
            const bool skip0 = solid && k0 == 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t ch = (i < 4 ? a_lo : a_hi) >> (8 * (i & 3)), dg = (i < 4 ? d_lo : d_hi) >> (8 * (i & 3));
                const uint32_t qv = (i < 4 ? q_lo : q_hi) >> (8 * (i & 3));
                const bool in = i < cnt;
                if (on0 && in && !(i == 0 && skip0)) {
                    switch (i) {      // the offset is part of the instruction
                        case 0: sts8<0>(ps_b, ch); sts8<0>(pq_b, qv); break; case 1: sts8<1>(ps_b, ch); sts8<1>(pq_b, qv); break;
                        case 2: sts8<2>(ps_b, ch); sts8<2>(pq_b, qv); break; case 3: sts8<3>(ps_b, ch); sts8<3>(pq_b, qv); break;
                        case 4: sts8<4>(ps_b, ch); sts8<4>(pq_b, qv); break; case 5: sts8<5>(ps_b, ch); sts8<5>(pq_b, qv); break;
                        case 6: sts8<6>(ps_b, ch); sts8<6>(pq_b, qv); break; default: sts8<7>(ps_b, ch); sts8<7>(pq_b, qv); break;
                    }
                }
                if (on2 && in) {
                    switch (i) {
                        case 0: sts8<0>(ps_f, dg); sts8<0>(pq_f, qv); break; case 1: sts8<1>(ps_f, dg); sts8<1>(pq_f, qv); break;
                        case 2: sts8<2>(ps_f, dg); sts8<2>(pq_f, qv); break; case 3: sts8<3>(ps_f, dg); sts8<3>(pq_f, qv); break;
                        case 4: sts8<4>(ps_f, dg); sts8<4>(pq_f, qv); break; case 5: sts8<5>(ps_f, dg); sts8<5>(pq_f, qv); break;
                        case 6: sts8<6>(ps_f, dg); sts8<6>(pq_f, qv); break; default: sts8<7>(ps_f, dg); sts8<7>(pq_f, qv); break;
                    }
                }
            }
        }
        // ---- step 2: names (one lane per record and 16-byte chunk), suffixes and separators ----------------------
        {
            const int nchunks = P.name_cap >> 4;
            for (int it = lane; it < np * 4 * nchunks; it += 32) {
                const int rc = (int)__umulhi((uint32_t)it, P.inv_name_chunks), c = it - rc * nchunks, t = rc >> 2, rr = rc & 3, e = rr & 1, bf = rr >> 1;
                const PairMeta &m = meta[t];
                const int Le = m.len[e];
                if (Le <= 0 || (bf ? !on2 : !on0)) continue;
                const int nn = bf ? m.nfull : m.nbwa, x0 = c << 4;
                if (x0 >= nn) continue;
                const int len0 = m.len[0];
                const int rec0 = len0 > 0 ? m.nfull + 1 + (solid ? 1 : 0) + 2 * len0 + 4 : 0;
                const uint32_t d = (bf ? a_st2 + m.so[2] + (e ? rec0 : 0) : (e ? a_st1 : a_st0) + m.so[e]) + x0;
                const char *nm = gnames + ((size_t)(p0 + t) * nvar + (bf ? 0 : nvar - 1)) * P.name_cap;
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(nm + x0));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const int left = nn - x0;
#pragma unroll
                for (int b4 = 0; b4 < 4; ++b4) {
                    const uint32_t x = w[b4];
                    if (4 * b4 + 0 < left) sts8(d + 4 * b4, x);
                    if (4 * b4 + 1 < left) sts8(d + 4 * b4 + 1, x >> 8);
                    if (4 * b4 + 2 < left) sts8(d + 4 * b4 + 2, x >> 16);
                    if (4 * b4 + 3 < left) sts8(d + 4 * b4 + 3, x >> 24);
                }
            }
            for (int it = lane; it < np * 4; it += 32) {
                const int t = it >> 2, rr = it & 3, e = rr & 1, bf = rr >> 1;
                const PairMeta &m = meta[t];
                const int Le = m.len[e];
                if (Le <= 0) continue;
                if (!bf) {
                    if (!on0) continue;
                    const uint32_t sb = (e ? a_st1 : a_st0) + m.so[e] + m.nbwa;
                    const int me = Le - from;
                    sts8(sb, '/'); sts8(sb + 1, solid ? (e == 0 ? '2' : '1') : (e == 0 ? '1' : '2')); sts8(sb + 2, '\n');
                    sts8(sb + 3 + me, '\n'); sts8(sb + 3 + me + 1, '+'); sts8(sb + 3 + me + 2, '\n');
                    sts8(sb + 3 + me + 3 + me, '\n');
                } else {
                    if (!on2) continue;
                    const int len0 = m.len[0];
                    const int rec0 = len0 > 0 ? m.nfull + 1 + (solid ? 1 : 0) + 2 * len0 + 4 : 0;
                    uint32_t sf = a_st2 + m.so[2] + (e ? rec0 : 0) + m.nfull;
                    sts8(sf, '\n');
                    sf += 1;
                    if (solid) { sts8(sf, 'A'); sf += 1; }
                    sts8(sf + Le, '\n'); sts8(sf + Le + 1, '+'); sts8(sf + Le + 2, '\n');
                    sts8(sf + Le + 3 + Le, '\n');
                }
            }
        }
        __syncwarp();
        // ---- step 3: copy-out ------------------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (total[k] == 0) continue;
            const uint32_t a_st = k == 0 ? a_st0 : (k == 1 ? a_st1 : a_st2);
            char *base = (k == 0 ? out0 : (k == 1 ? out1 : out2)) + begin[k] - shift[k];   // 16-byte aligned
            const int sh = (int)shift[k], end = sh + (int)total[k], nchunk = (end + 15) >> 4;
            const int c_first = sh ? 1 : 0, c_full = end >> 4;           // chunks [c_first, c_full) lie wholly inside the range
            uint4 *dst = reinterpret_cast<uint4 *>(base);
            for (int c = c_first + lane; c < c_full; c += 32) dst[c] = lds128(a_st + (c << 4));
            // the two boundary chunks are shared with the neighbouring mini-tiles: lanes 0-15 / 16-31 store their bytes
            {
                const int c = lane < 16 ? 0 : c_full, x = (c << 4) + (lane & 15);
                const bool mine = lane < 16 ? c_first != 0 : (c_full < nchunk && (c_full > 0 || !c_first));
                if (mine && x >= sh && x < end) base[x] = (char)lds8(a_st + x);
            }
        }
        __syncwarp();
        cur = nxt;
    }
}


// ---- kernel B2: format, word-granular -----------------------------------------------------------------------
// Same mini-tile organisation, inputs and output bytes as format_fastq_kernel above, rebuilt around these changes:
//   * qualities: the 8 QUAL draws of a group are the eight 16-bit fields of ONE Philox block; a byte table indexed by the
//     16-bit draw gives the noise rank directly (P.qtab, 2^kQTabBits entries in shared memory), bit 7 marks the draws whose
//     table cell contains a CDF threshold: only those (about 5e-4 of the draws) take their lower 16 bits from the group's
//     second block and search the CDF.  Base + noise + clamp run two qualities per instruction (VIADDMNMX.S16x2.RELU)
//   * everything is assembled in registers and leaves as aligned 32-bit words after a byte funnel shift against the
//     previous lane's last word (SHFL + PRMT); no predicated byte stores in the bulk of the work.  Such a field spills
//     at most three bytes over each of its ends, so the order of the passes makes every spill land on bytes a later pass
//     writes: names first (their spill: the previous record's last bytes and the suffix / first bases), then bases and
//     qualities (spill: the 3-byte "\n+\n" between them, the suffix and up to two name bytes before the bases, the "\n"
//     and up to two name bytes after the qualities), then one lane per record rewrites exactly those bytes
//   * the 16-byte aligned interior of every stream range leaves with one cp.async.bulk.global.shared::cta per stream
//     (UBLKCP) while the warp goes on with its next mini-tile; the staging area is reused after wait_group.read
//   * a warp formats one contiguous range of mini-tiles.  The 16-byte chunk two neighbouring mini-tiles share is handed
//     from one to the next in shared memory (a_carry) and leaves inside the next bulk store; single bytes are stored
//     only where the range meets another warp's (two chunks per warp and stream)
//   * what step 0 needs of the NEXT mini-tile - six words per pair of PairRec / name lengths / stream offsets and the
//     read names - is on its way to shared memory while this one is formatted (cp.async, LDGSTS): no register holds
//     it, no load latency is left in the names pass.  The code words are loaded where they are used (their latency
//     hides behind the Philox rounds), the next mini-tile's lines are prefetched to L2
//   * the geometry that only depends on the read lengths (Format2Smem) is computed on the host and arrives in the
//     constant bank
// Specialisation: colour space.  Configurations whose quality sum can wrap in int8 or whose noise table has 128 or
// more steps (quality_std >= 7.8) stay on format_fastq_kernel.
#ifndef DWG_QTAB_BITS
#define DWG_QTAB_BITS 15
#endif
constexpr int kQTabBits = DWG_QTAB_BITS;
#ifndef DWG_FMT2_WARPS
#define DWG_FMT2_WARPS 24
#endif
constexpr int kFmt2WarpsMax = DWG_FMT2_WARPS;
constexpr int kFmt2ThreadsMax = kFmt2WarpsMax * 32;

struct ReadMeta {                      // per (pair, end) of the warp's mini-tile, in shared memory (32 bytes)
    uint32_t ps_b, pq_b, ps_f, pq_f;   // shared-window byte address of base / quality 0 in the bwa and the bfast record
    uint32_t len;                      // read length
    uint32_t c2;                       // Philox counter word 2 of the QUAL stream: attempt | stream << 16 | end << 24
    uint32_t g_lo, g_hi;               // global pair index (Philox counter words 0, 1)
};
struct NameMeta {                      // per record (pair, end, bwa | bfast), 16 bytes; dst == 0: no such record
    uint32_t dst;                      // shared-window byte address of the '@'
    uint32_t nn;                       // name length
    uint32_t src;                      // byte offset of the name text in gnames
    uint32_t len;                      // read length (bytes of bases in the record: len - from for bwa)
};

struct Format2Smem {                    // computed on the host, passed by value: every field is a constant-bank operand
    int qtab_off, cdf_off, qb_off[2], warp_off, warp_stride, rmeta_off, nmeta_off, stage_off[3], total;
    int pre_off, pre_bytes;            // per warp, two of each: the next mini-tile's FmtPrefetch words (32 bytes per lane 0..WP)
    int names_off, names_bytes;        //   and its read names (WP * nvar rows of name_cap bytes), both filled by cp.async
    int carry_off;                     // per warp: 3 x 16 bytes, the chunk a mini-tile leaves to the next one of its run
    int g0, h0, G, G2;                 // 8-base groups / 16-base lane items of end 0, code words / lane items per pair
};
__host__ __device__ inline Format2Smem format2_smem_layout(const SimParams &P)
{
    Format2Smem L;
    const int WP = P.tile_pairs;
    const bool noise = P.qdelta_n > 0 && !P.fixed_quality;
    int o = 0;
    L.qtab_off = o; o += noise ? (1 << kQTabBits) : 0;
    L.cdf_off = o; o += ((noise ? P.qdelta_n : 0) * 4 + 15) & ~15;
    for (int e = 0; e < 2; ++e) { L.qb_off[e] = o; o += (((P.cap[e] + 7) & ~7) * 2 + 16) & ~15; }
    L.warp_off = o;
    int w = 0;
    L.rmeta_off = w; w += 2 * WP * (int)sizeof(ReadMeta);
    L.nmeta_off = w; w += 4 * WP * (int)sizeof(NameMeta);
    L.pre_off = w; L.pre_bytes = (WP + 1) * 32; w += 2 * L.pre_bytes;
    L.names_off = w; L.names_bytes = WP * ((P.data_type == 1 && P.out_bwa) ? 2 : 1) * P.name_cap; w += 2 * L.names_bytes;
    L.carry_off = w; w += 64;
    for (int k = 0; k < 3; ++k) { L.stage_off[k] = w; w += (WP * P.rec_cap[k] + 32 + 15) & ~15; }
    L.warp_stride = w;
    L.total = o + (P.fmt_warps > 0 ? P.fmt_warps : kFmt2WarpsMax) * w;
    L.g0 = (P.cap[0] + 7) >> 3; L.G = P.nw[0] + P.nw[1];
    L.h0 = (L.g0 + 1) >> 1; L.G2 = L.h0 + ((L.G - L.g0 + 1) >> 1);      // (P.inv_groups = 2^32 / G2 + 1)
    return L;
}

__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// PRMT with the selector used as it is (__byte_perm masks it with 0x7777 first): every selector below has nibbles <= 7
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}

// Philox4x32-10 with the first key word's round keys taken from the kernel parameters (P.qkey[r] = seed + r * 0x9E3779B9:
// constant-bank operands of the XORs) and the second key word's folded into immediates
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const SimParams &P)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ P.qkey[r], n2 = h0 ^ c3 ^ (kPhiloxKey1 + (uint32_t)r * 0xBB67AE85u);
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// one field (bases or qualities of one record): the lane's 16 bytes x[0..3] belong at byte address p; `prev` is the last
// word of the previous 16 bytes of the same read (anything for the first).
// `skip1`: the lane's first byte is not part of the field (first colour of a SOLiD bwa record): when it is the last byte
// of its word that word holds nothing of the field, and writing it would reach four bytes back
__device__ __forceinline__ void store_field16(uint32_t p, uint32_t prev, const uint32_t (&x)[4], bool tail, int cnt, bool skip1)
{
    const uint32_t bs = p & 3u, w = p - bs, sel = 0x7654u - 0x1111u * bs;       // bytes [4 - bs, 8 - bs) of {first, second}
    const int v = (int)bs + cnt;                                                 // end of the lane's bytes in its 20-byte window
    if (!(skip1 && bs == 3u)) sts32(w, prmt(prev, x[0], sel));
    if (!tail || v > 4) sts32(w + 4, prmt(x[0], x[1], sel));
    if (!tail || v > 8) sts32(w + 8, prmt(x[1], x[2], sel));
    if (!tail || v > 12) sts32(w + 12, prmt(x[2], x[3], sel));
    if (tail && v > 16) sts32(w + 16, prmt(x[3], 0u, sel));
}

// the draws of a group whose table cell holds a CDF threshold: full 32-bit draw, rank by binary search
__device__ __noinline__ uint4 qual_ranks_slow(uint4 r, const uint4 b0, const uint4 b1, const uint32_t *cdf, int n)
{
    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t sh = 16u * (i & 1);
        if (!((rr[i >> 1] >> sh) & 0x80u)) continue;
        const uint32_t u = qual_draw32(b0, b1, i);
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (u >= cdf[mid]) lo = mid + 1; else hi = mid; }
        rr[i >> 1] = (rr[i >> 1] & ~(0xFFFFu << sh)) | ((uint32_t)lo << sh);
    }
    return make_uint4(rr[0], rr[1], rr[2], rr[3]);
}

__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Ampere-style asynchronous copies global -> shared (LDGSTS): no register holds the data in flight
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// the low `nb` bytes of c, the others of v (nb <= 0: v, nb >= 4: c)
__device__ __forceinline__ uint32_t merge_low(uint32_t c, uint32_t v, int nb)
{
    const uint32_t m = nb <= 0 ? 0u : (nb >= 4 ? 0xFFFFFFFFu : (1u << (8 * nb)) - 1u);
    return (c & m) | (v & ~m);
}
template <typename T> __device__ __forceinline__ T pick3(int k, T a, T b, T c) { return k == 0 ? a : (k == 1 ? b : c); }

template <bool kSolid, int kQMode /* qualities: 0 no noise, 1 noise table, 2 fixed character (-Q / fixed quality) */,
          int kOut /* bit 0: the two bwa files, bit 1: the bfast file (-o) */>
__global__ void __launch_bounds__(kFmt2ThreadsMax, 1)
format_fastq2_kernel(const SimParams P, const Format2Smem L, int64_t first, int64_t gidx_origin, int n,
                     const PairRec *__restrict__ recs, const uint32_t *__restrict__ seqw,
                     const uint32_t *__restrict__ offs /* [3][n] */,
                     const unsigned long long *__restrict__ totals /* bytes of the batch per stream */,
                     const char *__restrict__ gnames, const uint16_t *__restrict__ gname_len,
                     char *__restrict__ out0, char *__restrict__ out1, char *__restrict__ out2)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int WP = P.tile_pairs, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int qmode = kQMode;
    uint32_t a_qtab, a_qb0, a_qb1, a_rm, a_nm, a_st0, a_st1, a_st2, a_pre, a_names, a_carry;
    const uint32_t *cdf;
    {
        if (qmode == 1) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.qtab);
            uint4 *dst = reinterpret_cast<uint4 *>(smem + L.qtab_off);
            for (int j = tid; j < (1 << kQTabBits) / 16; j += (int)blockDim.x) dst[j] = __ldg(src + j);
            for (int j = tid; j < P.qdelta_n; j += (int)blockDim.x) reinterpret_cast<uint32_t *>(smem + L.cdf_off)[j] = P.qdelta_cdf[j];
        }
        for (int e = 0; e < 2; ++e) {
            int16_t *qb = reinterpret_cast<int16_t *>(smem + L.qb_off[e]);
            const int padded = (P.cap[e] + 7) & ~7;
            for (int j = tid; j < padded; j += (int)blockDim.x) qb[j] = (int16_t)((j < P.cap[e] ? (int)P.qbase[e][j] : 0) + (qmode == 1 ? P.qdelta_lo : 0));
        }
        const uint32_t a_base = smem_addr(smem);
        a_qtab = in_register(a_base + L.qtab_off); a_qb0 = a_base + L.qb_off[0]; a_qb1 = a_base + L.qb_off[1];
        const uint32_t a_warp = a_base + L.warp_off + warp * L.warp_stride;
        a_rm = in_register(a_warp + L.rmeta_off); a_nm = in_register(a_warp + L.nmeta_off);
        a_st0 = a_warp + L.stage_off[0]; a_st1 = a_warp + L.stage_off[1]; a_st2 = a_warp + L.stage_off[2];
        a_pre = a_warp + L.pre_off; a_names = a_warp + L.names_off; a_carry = a_warp + L.carry_off + lane * 16;
        cdf = reinterpret_cast<const uint32_t *>(smem + L.cdf_off);
    }
    __syncthreads();                                                 // the only CTA-wide barrier

    constexpr bool solid = kSolid;
    constexpr int from = kSolid ? 1 : 0;                            // bwa drops the first colour (src/dwgsim.c:949-953)
    constexpr int sfx_f = kSolid ? 2 : 1;                           // bytes between a bfast name and its first base: "\n" ("\nA")
    constexpr int nvar = (kSolid && (kOut & 1)) ? 2 : 1;            // SOLiD bwa names carry reduced counts (:945-946)
    constexpr bool on0 = (kOut & 1) != 0, on2 = (kOut & 2) != 0;
    const int ntiles = (n + WP - 1) / WP;
    constexpr uint32_t full = 0xffffffffu;
    const uint32_t al0 = (uint32_t)reinterpret_cast<uintptr_t>(out0) & 15u, al1 = (uint32_t)reinterpret_cast<uintptr_t>(out1) & 15u,
                   al2 = (uint32_t)reinterpret_cast<uintptr_t>(out2) & 15u;

    // step 0 of a mini-tile, one mini-tile ahead: the FmtPrefetch words of lane j (pair j; lane np: the end offsets) and the
    // read names on their way to shared memory buffer `b` (cp.async: nothing of it lives in registers meanwhile)
    auto issue = [&](int tile, int b) {
        if (tile >= ntiles) return;
        const int p0 = tile * WP, np = min(WP, n - p0), p = p0 + lane;
        const uint32_t slot = a_pre + b * L.pre_bytes + lane * 32;
        if (lane < np) {
            const uint32_t *r = reinterpret_cast<const uint32_t *>(recs + p);
            cp_async4(slot, r + 2); cp_async4(slot + 4, r + 7);
            cp_async4(slot + 8, reinterpret_cast<const uint32_t *>(gname_len) + p);
        }
        if (lane <= np) {
            if (p < n) {
                if (on0) { cp_async4(slot + 12, offs + p); cp_async4(slot + 16, offs + (size_t)n + p); }
                if (on2) cp_async4(slot + 20, offs + (size_t)2 * n + p);
            } else {                                                 // the end of the batch
                if (on0) { cp_async4(slot + 12, totals); cp_async4(slot + 16, totals + 1); }
                if (on2) cp_async4(slot + 20, totals + 2);
            }
        }
        const int nchunk = (np * nvar * P.name_cap) >> 4;
        const char *src = gnames + (size_t)p0 * nvar * P.name_cap;
        for (int c = lane; c < nchunk; c += 32) cp_async16(a_names + b * L.names_bytes + (c << 4), src + ((size_t)c << 4));
    };
    // every warp formats one contiguous range of mini-tiles: the 16-byte chunk two neighbouring mini-tiles share stays in
    // the warp (a_carry), only the two ends of the range meet another warp's bytes
    int tile, run_end;
    {
        const int n_warps = (int)blockDim.x >> 5, gw = blockIdx.x * n_warps + warp, tw = gridDim.x * n_warps;
        tile = (int)((long long)ntiles * gw / tw); run_end = (int)((long long)ntiles * (gw + 1) / tw);
    }
    int buf = 0;
    issue(tile, 0);
    bool bulk_pending = false;                                       // (lanes 0-2) a bulk store of the previous mini-tile may still read the staging area
    int carry_lo = -1;                                               // (lane k < 3) stream k: first valid byte of the chunk the previous mini-tile
                                                                     //   of the run left at a_carry; -1: nothing carried
    for (; tile < run_end; ++tile) {
        const bool last_in_run = tile + 1 == run_end;
        const int tile_next = last_in_run ? ntiles : tile + 1;
        cp_async_wait_all();                                         // this lane's copies for this mini-tile have landed
        if (bulk_pending) { bulk_wait_read(); bulk_pending = false; }   // the staging area is free again
        __syncwarp();                                                // ... and every other lane's
        const uint32_t a_pb = a_pre + buf * L.pre_bytes;              // slot j (32 bytes): lens, tail, name lengths, the three offsets of pair j
        const uint32_t a_nb = a_names + buf * L.names_bytes;          // the read names of the mini-tile
        issue(tile_next, buf ^ 1);
        const int p0 = tile * WP, np = min(WP, n - p0);
        const uint32_t *seqw_tile = seqw + (size_t)p0 * L.G;          // code word of item `it` of this mini-tile: seqw_tile[it]
        {   // the read codes of the next mini-tile towards L2 (measured: -2% kernel time; an L1 prefetch of this one's: nothing)
            const uint32_t lo = (uint32_t)lane << 7;
            const int pn = tile_next * WP;
            if (pn < n) {
                const uint32_t cn = (uint32_t)min(WP, n - pn) * (uint32_t)L.G * 4u;
                if (lo < cn + 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(seqw + (size_t)pn * L.G) + min(lo, cn - 1u)));
            }
        }
        // ---- step 0: geometry of the mini-tile, then per-read and per-record metadata ----------------------------------
        uint32_t bg0 = 0, bg1 = 0, bg2 = 0, tt0 = 0, tt1 = 0, tt2 = 0;      // first byte and bytes of the mini-tile per stream
        if (on0) { bg0 = lds32(a_pb + 12); bg1 = lds32(a_pb + 16); tt0 = lds32(a_pb + np * 32 + 12) - bg0; tt1 = lds32(a_pb + np * 32 + 16) - bg1; }
        if (on2) { bg2 = lds32(a_pb + 20); tt2 = lds32(a_pb + np * 32 + 20) - bg2; }
        const uint32_t sh0 = (al0 + bg0) & 15u, sh1 = (al1 + bg1) & 15u, sh2 = (al2 + bg2) & 15u;   // their misalignment in the output
        const uint32_t my_bg = pick3(lane, bg0, bg1, bg2), my_tot = pick3(lane, tt0, tt1, tt2), my_sh = pick3(lane, sh0, sh1, sh2);   // (step 4)
        int max_nn = 0;
        for (int rd = lane; rd < 2 * np; rd += 32) {                  // ReadMeta + the read's two NameMeta
            const int t = rd >> 1, e = rd & 1;
            const uint4 pv = lds128(a_pb + t * 32);                    // lens, tail, name lengths, offset in stream 0
            const uint2 pu = lds64(a_pb + t * 32 + 16);                // offsets in streams 1 and 2
            const int len0 = (int)(pv.x & 0xFFFFu), Le = e ? (int)(pv.x >> 16) : len0;
            const int nfull = (int)(pv.z & 0xFFFFu), nbwa = (int)(pv.z >> 16);
            const uint32_t rec_b = e ? a_st1 + (pu.x - bg1 + sh1) : a_st0 + (pv.w - bg0 + sh0);
            const int rec0 = len0 > 0 ? nfull + sfx_f + 2 * len0 + 4 : 0;
            const uint32_t rec_f = a_st2 + (pu.y - bg2 + sh2) + (e ? rec0 : 0);
            const uint64_t gidx = (uint64_t)(gidx_origin + first + p0 + t);
            uint4 r0, r1;
            r0.x = rec_b + nbwa + 3 - from; r0.y = r0.x + (Le - from) + 3;
            r0.z = rec_f + nfull + sfx_f; r0.w = r0.z + Le + 3;
            r1.x = (uint32_t)Le; r1.y = (pv.y >> 16) | (kStQual << 16) | ((uint32_t)e << 24);
            r1.z = (uint32_t)gidx; r1.w = (uint32_t)(gidx >> 32);
            sts128(a_rm + rd * 32, r0); sts128(a_rm + rd * 32 + 16, r1);
            const bool has = Le > 0;
            const uint32_t src = (uint32_t)t * (uint32_t)nvar * (uint32_t)P.name_cap;   // relative to the names of the mini-tile
            sts128(a_nm + (rd * 2) * 16, make_uint4(has && on0 ? rec_b : 0u, (uint32_t)nbwa, src + (uint32_t)(nvar - 1) * P.name_cap, (uint32_t)(Le - from)));
            sts128(a_nm + (rd * 2 + 1) * 16, make_uint4(has && on2 ? rec_f : 0u, (uint32_t)nfull, src, (uint32_t)Le));
            if (has) max_nn = max(max_nn, max(on0 ? nbwa : 0, on2 ? nfull : 0));
        }
        max_nn = __reduce_max_sync(full, max_nn);
        __syncwarp();
        // ---- step 1: names, one lane per (record, 16-byte chunk); aligned words, spills allowed (see the kernel comment) ---
        for (int c0 = 0; c0 < max_nn; c0 += 64)
            for (int rb = 0; rb < 4 * np; rb += 8) {
                const int rc = rb + (lane >> 2), x0 = c0 + ((lane & 3) << 4);
                uint4 nm = make_uint4(0, 0, 0, 0);
                if (rc < 4 * np) nm = lds128(a_nm + rc * 16);
                const bool act = nm.x != 0 && x0 < (int)nm.y;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (act) v = lds128(a_nb + nm.z + x0);
                uint32_t pw = __shfl_up_sync(full, v.w, 1);               // last word of the previous chunk of the same name
                if ((lane & 3) == 0) pw = c0 ? lds32(a_nb + nm.z + x0 - 4) : 0u;
                if (act) {
                    const uint32_t d = nm.x + x0, bs = d & 3u, w = d - bs, sel = 0x7654u - 0x1111u * bs;
                    const int left = (int)nm.y - x0 + (int)bs;             // name bytes from w on
                    sts32(w, prmt(pw, v.x, sel));
                    if (left > 4) sts32(w + 4, prmt(v.x, v.y, sel));
                    if (left > 8) sts32(w + 8, prmt(v.y, v.z, sel));
                    if (left > 12) sts32(w + 12, prmt(v.z, v.w, sel));
                    if (left > 16 && (int)nm.y - x0 <= 16) sts32(w + 16, prmt(v.w, 0u, sel));   // (the next chunk's first word otherwise)
                }
            }
        __syncwarp();
        // ---- step 2: bases and qualities, one lane per (pair, end, 16 bases = two 8-base groups), aligned word stores ----
        const int items = np * L.G2, n_iter = (items + 31) >> 5;
        uint32_t carry_a = 0, carry_d = 0, carry_q = 0;
        for (int iter = 0; iter < n_iter; ++iter) {
            const int it = (iter << 5) + lane;
            const int t = (int)__umulhi((uint32_t)it, P.inv_groups), gi = it - t * L.G2;
            const int e = gi < L.h0 ? 0 : 1, j = gi - (e ? L.h0 : 0), rd = 2 * t + e, k0 = j << 4;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (it < items) { r0 = lds128(a_rm + rd * 32); r1 = lds128(a_rm + rd * 32 + 16); }
            const int Le = (int)r1.x;
            const bool active = k0 < Le, active_b = k0 + 8 < Le;       // (Le == 0 beyond the mini-tile)
            uint32_t cw[2] = {0, 0};
            {
                const uint32_t *src = seqw_tile + (t * L.G + (e ? L.g0 : 0) + 2 * j);
                if (active) cw[0] = __ldg(src);
                if (active_b) cw[1] = __ldg(src + 1);
            }
            uint32_t a[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};   // 16 bases (bwa), 16 bases (bfast), 16 qualities
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                if (!(sub ? active_b : active)) continue;
                const uint32_t codes = cw[sub];
                // qualities, src/dwgsim.c:899-918
                if (qmode == 2) q[2 * sub] = q[2 * sub + 1] = 0x01010101u * (uint32_t)P.fixed_quality;
                else {
                    const uint4 qb = lds128((e ? a_qb1 : a_qb0) + ((k0 + 8 * sub) << 1));    // 8 x int16: Phred base (+ lowest noise step)
                    uint4 r = make_uint4(0, 0, 0, 0);
                    if (qmode == 1) {
                        const uint32_t blk = (uint32_t)(4 * j + 2 * sub);                       // the group's first QUAL block
                        const uint4 b0 = philox4x32_10_rk(r1.z, r1.w, r1.y, blk, P);
                        // the table cell of a 16-bit draw: its upper kQTabBits bits
#define DWG_LK_LO(word) lds8(a_qtab + (((word) << 16) >> (32 - kQTabBits)))
#define DWG_LK_HI(word) lds8(a_qtab + ((word) >> (32 - kQTabBits)))
                        r.x = DWG_LK_LO(b0.x) | (DWG_LK_HI(b0.x) << 16);
                        r.y = DWG_LK_LO(b0.y) | (DWG_LK_HI(b0.y) << 16);
                        r.z = DWG_LK_LO(b0.z) | (DWG_LK_HI(b0.z) << 16);
                        r.w = DWG_LK_LO(b0.w) | (DWG_LK_HI(b0.w) << 16);
#undef DWG_LK_LO
#undef DWG_LK_HI
                        if ((r.x | r.y | r.z | r.w) & 0x00800080u)
                            r = qual_ranks_slow(r, b0, philox4x32_10_rk(r1.z, r1.w, r1.y, blk + 1u, P), cdf, P.qdelta_n);
                    }
                    // 33 + clamp(base + noise, 0, 40), two qualities per instruction
                    const uint32_t v0 = __viaddmin_s16x2_relu(r.x, qb.x, 0x00280028u), v1 = __viaddmin_s16x2_relu(r.y, qb.y, 0x00280028u);
                    const uint32_t v2 = __viaddmin_s16x2_relu(r.z, qb.z, 0x00280028u), v3 = __viaddmin_s16x2_relu(r.w, qb.w, 0x00280028u);
                    q[2 * sub] = prmt(v0, v1, 0x6420u) + 0x21212121u;
                    q[2 * sub + 1] = prmt(v2, v3, 0x6420u) + 0x21212121u;
                }
                // 8 nibble codes (0-4) -> 8 characters: the nibbles are PRMT selectors into "ACGTN" / "01234"
                a[2 * sub] = prmt(0x54474341u, 0x0000004Eu, codes); a[2 * sub + 1] = prmt(0x54474341u, 0x0000004Eu, codes >> 16);
                if (solid) { d[2 * sub] = prmt(0x33323130u, 0x00000034u, codes); d[2 * sub + 1] = prmt(0x33323130u, 0x00000034u, codes >> 16); }
            }
            // the last word of the previous 16 bases (previous lane; lane 0: last lane of the previous round)
            uint32_t pa = __shfl_up_sync(full, a[3], 1), pq = __shfl_up_sync(full, q[3], 1), pd = pa;
            if (solid) pd = __shfl_up_sync(full, d[3], 1);
            if (lane == 0) { pa = carry_a; pq = carry_q; pd = carry_d; }
            carry_a = __shfl_sync(full, a[3], 31); carry_q = __shfl_sync(full, q[3], 31);
            carry_d = solid ? __shfl_sync(full, d[3], 31) : carry_a;
            if (active) {
                const int cnt = min(16, Le - k0);
                const bool tail = k0 + 16 >= Le;
                if (on0) {
                    store_field16(r0.x + k0, pa, a, tail, cnt, solid && j == 0);
                    store_field16(r0.y + k0, pq, q, tail, cnt, solid && j == 0);
                }
                if (on2) {
                    store_field16(r0.z + k0, pd, solid ? d : a, tail, cnt, false);
                    store_field16(r0.w + k0, pq, q, tail, cnt, false);
                }
            }
        }
        __syncwarp();
        // ---- step 3: one lane per record rewrites the bytes the spills may have touched: the first and the last two name
        //      bytes, the suffix, "\n+\n" and the closing "\n" -----------------------------------------------------------
        for (int rc = lane; rc < 4 * np; rc += 32) {
            const uint4 nm = lds128(a_nm + rc * 16);
            if (nm.x == 0) continue;
            const int bf = rc & 1, e = (rc >> 1) & 1, nn = (int)nm.y, Lr = (int)nm.w;
            const uint32_t head = lds32(a_nb + nm.z);
            sts8(nm.x, head); sts8(nm.x + 1, head >> 8);
            uint32_t s = nm.x + nn;
            sts8(s - 2, lds8(a_nb + nm.z + nn - 2)); sts8(s - 1, lds8(a_nb + nm.z + nn - 1));
            if (!bf) { sts8(s, '/'); sts8(s + 1, solid ? (e == 0 ? '2' : '1') : (e == 0 ? '1' : '2')); sts8(s + 2, '\n'); s += 3; }
            else { sts8(s, '\n'); s += 1; if (solid) { sts8(s, 'A'); s += 1; } }
            sts8(s + Lr, '\n'); sts8(s + Lr + 1, '+'); sts8(s + Lr + 2, '\n');
            sts8(s + Lr + 3 + Lr, '\n');
        }
        fence_async_smem();                                          // the bulk engine reads what the lanes wrote
        __syncwarp();
        // ---- step 4: copy-out, lane k < 3 for stream k.  Staging byte x of the stream belongs at base[x]; the bytes that
        //      are this warp's to store are [lo, end): lo = sh, or the start of what the previous mini-tile of the run left
        //      behind in `carry` (the chunk the two share).  Whole 16-byte chunks leave as one bulk store; the partial chunk
        //      at the end is carried to the next mini-tile of the run; only the ends of a run store single bytes -----------
        if (lane < 3) {
            const uint32_t a_st = pick3(lane, a_st0, a_st1, a_st2);
            const int sh = (int)my_sh, tot = (int)my_tot;
            if (tot != 0 || carry_lo >= 0) {
                char *base = pick3(lane, out0, out1, out2) + my_bg - sh;                         // 16-byte aligned
                const int end = sh + tot;
                int lo = sh;
                if (carry_lo >= 0) {                                 // chunk 0: the carried bytes below sh, this mini-tile's from sh on
                    const uint4 carry = lds128(a_carry);
                    uint4 c = lds128(a_st);
                    c.x = merge_low(carry.x, c.x, sh); c.y = merge_low(carry.y, c.y, sh - 4);
                    c.z = merge_low(carry.z, c.z, sh - 8); c.w = merge_low(carry.w, c.w, sh - 12);
                    sts128(a_st, c);
                    fence_async_smem();
                    lo = carry_lo;
                }
                const int c_lo = (lo + 15) >> 4, c_full = end >> 4;  // whole chunks: [c_lo, c_full)
                if (c_full > c_lo) {
                    bulk_store(base + (c_lo << 4), a_st + (c_lo << 4), (uint32_t)(c_full - c_lo) << 4);
                    bulk_commit();
                    bulk_pending = true;
                }
                const bool reach = c_full >= c_lo;                   // the staged bytes reach the chunk boundary c_lo
                if (reach && (lo & 15)) {                            // head of a run: bytes [lo, 16)
#pragma unroll
                    for (int x = 1; x < 16; ++x) if (x >= lo) base[x] = (char)lds8(a_st + x);
                }
                const int t0 = reach ? (c_full << 4) : lo;           // the bytes after the last whole chunk: [t0, end)
                carry_lo = -1;
                if (end > t0) {
                    if (!last_in_run) { sts128(a_carry, lds128(a_st + (c_full << 4))); carry_lo = reach ? 0 : lo; }
                    else {
#pragma unroll
                        for (int xx = 0; xx < 16; ++xx) { const int x = (c_full << 4) + xx; if (x >= t0 && x < end) base[x] = (char)lds8(a_st + x); }
                    }
                }
            }
        }
        buf ^= 1;
    }
    if (bulk_pending) bulk_wait_read();
}


}  // namespace dwg
