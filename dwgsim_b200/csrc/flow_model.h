// flow_model.h -- the Ion Torrent flow-space error model (generate_errors_flows, src/dwgsim.c:246-417) of ONE read as two
// streaming passes over nibble-packed rows.  Host and device share this source: tests/flow_model_check.cpp compiles it
// with g++ and compares it with the oracle's restatement of the reference on random reads; simulate_pairs_tp_kernel<true>
// runs it one thread per read.
//
// The reference edits the read in place (every insertion / deletion shifts the tail) and asks its error coin once per
// homopolymer start (pass 1) and once per empty flow (pass 2).  Here
//   * the coin is a Bernoulli process generated from its geometric gaps (FlowCoin: `left` failures are known to come before
//     the next success), so a base whose trials all fail costs no draw and no loop over its flows: the distance to the
//     base's next flow comes from a table (nd) and is subtracted from `left`;
//   * both passes read a source row front to back and append to a destination row, so an insertion emits extra symbols
//     and a deletion skips source symbols -- nothing is shifted.  What the reference's in-place loop does to the symbols
//     after an edit (they are passed without a trial, or visited later with the flow pointer elsewhere) is reproduced
//     symbol by symbol; see the comments at the events.
// Rows hold 8 symbols per 32-bit word (codes 0-3), `cap` symbols of room each.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DWG_HD __host__ __device__ __forceinline__
#else
#define DWG_HD inline
#endif

namespace dwg {

constexpr int kFlowGapTableN = 4096;   // == kFlowGapN (layout.h) == ORC_FLOW_GAP_N (oracle)

// a row of symbols, 8 per 32-bit word; word k lives at p[k * stride] (stride 1: a plain array; stride 32: the rows of a
// warp's 32 threads interleaved word by word, so that the lanes' accesses to "their word k" coalesce)
struct FlowRow {
    uint32_t *p;
    int stride;
    DWG_HD uint32_t &operator[](int k) const { return p[k * stride]; }
};
DWG_HD uint32_t fm_get(const FlowRow &r, int k) { return (r[k >> 3] >> ((k & 7) << 2)) & 15u; }
DWG_HD void fm_set(const FlowRow &r, int k, uint32_t v)
{
    const int sh = (k & 7) << 2;
    r[k >> 3] = (r[k >> 3] & ~(15u << sh)) | (v << sh);
}
DWG_HD uint32_t fm_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
DWG_HD uint32_t fm_brev(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
DWG_HD uint32_t fm_rev_nibbles(uint32_t x)
{
    x = fm_brev(x);
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    return ((x & 0x33333333u) << 2) | ((x >> 2) & 0x33333333u);
}
// dst[0..len) = src[len-1..0]
DWG_HD void fm_reverse_into(const FlowRow &dst, const FlowRow &src, int len)
{
    const int nw = (len + 7) >> 3;
    for (int dw = 0; dw < nw; ++dw) {
        const int a = len - 8 - 8 * dw;                        // source symbol of the word's LAST nibble
        uint32_t x;
        if (a >= 0) {
            const int sh = (a & 7) << 2;
            const uint32_t lo = src[a >> 3], hi = sh ? src[(a >> 3) + 1] : 0u;
            x = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        } else x = src[0] << ((-a) << 2);
        x = fm_rev_nibbles(x);
        if (dw == nw - 1 && (len & 7)) x &= ~(~0u << ((len & 7) << 2));
        dst[dw] = x;
    }
}

// #{j < n : u >= cdf[j]} for a non-decreasing table
DWG_HD int fm_rank(const uint32_t *cdf, int n, uint32_t u)
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (u >= cdf[mid]) lo = mid + 1; else hi = mid; }
    return lo;
}

// The per-flow error coin and the uniforms of its events over two sources of sequential 32-bit draws (Draw::gap_word(),
// Draw::unif_word(): oracle flow_coin / flow_unif).  The gap draws are a plain sequence, so a caller may hand over the
// first gaps already ranked (q, qn): the device draws them for all the lanes of a warp at once.
template <class Draw>
struct FlowCoin {
    Draw draw;
    const uint32_t *gap;               // P(gap <= g) as 32-bit thresholds, kFlowGapTableN entries
    const uint16_t *q;                 // gaps drawn ahead (ranks), consumed before any further draw
    int qn, qi;
    int left;                          // failures still to come before `succ` (or before the next draw)
    bool succ, need;
    DWG_HD FlowCoin(const Draw &d, const uint32_t *g, const uint16_t *ahead = nullptr, int n_ahead = 0)
        : draw(d), gap(g), q(ahead), qn(n_ahead), qi(0), left(0), succ(false), need(true) {}
    DWG_HD bool coin()
    {
        for (;;) {
            if (need) {
                const int g = qi < qn ? (int)q[qi++] : fm_rank(gap, kFlowGapTableN, draw.gap_word());
                left = g; succ = g < kFlowGapTableN; need = false;
            }
            if (left > 0) { --left; return false; }
            need = true;
            if (succ) return true;
        }
    }
    DWG_HD bool fails(int d)           // true (and d trials consumed) when the next d trials are known to fail
    {
        if (need || left < d) return false;
        left -= d;
        return true;
    }
    DWG_HD uint32_t unif() { return draw.unif_word(); }
};

// appends symbols to a row, eight at a time
struct FlowEmit {
    FlowRow dst;
    uint32_t acc;
    int sh, w, n;                      // bit position in acc, next word, symbols appended so far
    DWG_HD void begin(const FlowRow &d) { dst = d; acc = 0; sh = 0; w = 0; n = 0; }
    DWG_HD void put(uint32_t c)
    {
        acc |= c << sh;
        sh += 4; ++n;
        // (written so that it compiles to one predicated store and selects: the lanes of a warp fill their words at different
        // symbols once their reads have had different edits, and a branch would run once per group of lanes)
        const bool full = sh == 32;
        if (full) dst[w] = acc;
        w += full ? 1 : 0; acc = full ? 0u : acc; sh = full ? 0 : sh;
    }
    DWG_HD void flush() { if (sh) dst[w] = acc; }               // (the word keeps zeros above the last symbol)
};
// reads a row front to back
struct FlowSrc {
    FlowRow r;
    uint32_t cur;
    int pos;
    DWG_HD void seek(const FlowRow &row, int k) { r = row; pos = k; cur = (k & 7) ? row[k >> 3] >> ((k & 7) << 2) : 0u; }
    DWG_HD uint32_t next()
    {
        if ((pos & 7) == 0) cur = r[pos >> 3];
        const uint32_t c = cur & 15u;
        cur >>= 4; ++pos;
        return c;
    }
};

// nd[f * 4 + b]: steps from flow f to the first flow f' >= f (cyclically) whose base is b; 0 when fo[f] == b
DWG_HD int fm_build_nd_entry(const int8_t *fo, int fl, int f, int b)
{
    for (int d = 0; d < fl; ++d) if (fo[(f + d) % fl] == b) return d;
    return fl;                                                  // base not in the flow order
}

// mask: one bit per flow, set when a deletion left the flow's homopolymer short (pass 2 then adds nothing there)
DWG_HD bool fm_mask_get(const uint32_t *mask, int f) { return (mask[f >> 5] >> (f & 31)) & 1u; }

// Returns the new length (result in A), -1 when the first base has no flow; *overflow: 1 = the read would grow to `cap`
// symbols, 2 = the reference's assert(0 < j) at src/dwgsim.c:348.  A and B: rows with room for `cap` symbols.
template <class Draw>
DWG_HD int flow_model_rows(const FlowRow &A, const FlowRow &B, int len, int cap, int strand, const int8_t *fo, int fl, const uint16_t *nd,
                           uint32_t *mask, FlowCoin<Draw> &rng, int *n_err_out, int *overflow)
{
    {   // N -> A (src/dwgsim.c:253-257); per-read mask
        const int nw = (len + 7) >> 3;
        for (int w = 0; w < nw; ++w) { const uint32_t x = A[w]; A[w] = x & ~(((x & 0x44444444u) >> 2) * 15u); }
        for (int w = 0; w < ((fl + 31) >> 5); ++w) mask[w] = 0;
    }
    FlowRow S = A, D = B;                                       // pass 1: S -> D, pass 2: D -> S
    if (strand) { fm_reverse_into(B, A, len); S = B; D = A; }
    int flow_i;
    {
        const int c = len > 0 ? (int)fm_get(S, 0) : 0;
        for (flow_i = 0; flow_i < fl; ++flow_i) if (c == fo[flow_i]) break;
        if (flow_i == fl) return -1;
    }
    int mask_cnt = 0;                                           // set bits of mask
    // ---- pass 1 (src/dwgsim.c:281-364): one trial per homopolymer start -----------------------------------------
    {
        const int src_len = len;
        uint32_t prev_c = 4;
        FlowEmit E;
        E.begin(D);
        FlowSrc in;
        in.seek(S, 0);
        while (E.n < len) {                                     // invariant: len - E.n == src_len - in.pos
            const uint32_t c = in.next();
            const bool start = c != prev_c;
            // the flow pointer moves to the base's flow: d = 0 inside a homopolymer (fo[flow_i] == prev_c there)
            const int d = nd[flow_i * 4 + (int)c];
            if (!start || (mask_cnt == 0 && rng.fails(1))) {    // the common cases: no trial, or a trial known to fail
                flow_i += d; if (flow_i >= fl) flow_i -= fl;
                prev_c = c;
                E.put(c);
                continue;
            }
            // first base of a homopolymer, the general way: the flows the pointer passes (and the one it lands on) lose
            // their mask bit, then the trial
            int si = in.pos - 1;                                // source index of c
            if (mask_cnt) {
                for (int t = 0; t <= d; ++t) {
                    int f = flow_i + t; if (f >= fl) f -= fl;
                    if (fm_mask_get(mask, f)) { mask[f >> 5] &= ~(1u << (f & 31)); --mask_cnt; }
                }
            }
            flow_i += d; if (flow_i >= fl) flow_i -= fl;
            const uint32_t before = prev_c;                     // base of the previous homopolymer (4: none yet)
            prev_c = c;
            int n_err = 0;
            while (rng.coin()) ++n_err;
            if (n_err == 0) { E.put(c); continue; }
            if (!(rng.unif() >> 31)) {                          // U < 0.5: over-call, n_err more copies of the base
                if (len + n_err >= cap) { *overflow = 1; return len; }
                for (int k = 0; k < n_err; ++k) E.put(c);       // (the reference then walks over them: same base, no trial)
                len += n_err;
                E.put(c);
            } else {                                            // under-call, bounded by the homopolymer
                int j = si;
                while (j < src_len && fm_get(S, j) == c) ++j;
                const int hp_l = j - si;
                const uint32_t next_c = j < src_len ? fm_get(S, j) : c;     // (the reference's scan leaves c here at the end of the read)
                if (hp_l < n_err) n_err = hp_l;
                si += n_err; len -= n_err;
                if (!fm_mask_get(mask, flow_i)) { mask[flow_i >> 5] |= 1u << (flow_i & 31); ++mask_cnt; }
                if (n_err < hp_l) { E.put(c); ++si; }           // the rest of the homopolymer follows
                else if (E.n == 0 || before == next_c) {
                    // the whole homopolymer vanished between equal neighbours (or at the start): "dot-fill" with the base of
                    // a flow between this one and the next base's (src/dwgsim.c:342-358); the reference passes it unexamined
                    const int jj = nd[flow_i * 4 + (int)next_c];
                    if (jj <= 0) { *overflow = 2; return len; }
                    const int k = (int)fm_mulhi(rng.unif(), (uint32_t)jj);
                    if (len + 1 >= cap) { *overflow = 1; return len; }
                    int f = flow_i + k; if (f >= fl) f -= fl;
                    E.put((uint32_t)fo[f]);
                    len += 1;
                } else if (si < src_len) { E.put(fm_get(S, si)); ++si; }    // the next symbol takes the position: passed without a
                                                                            // trial and without moving the flow pointer
                in.seek(S, si);
            }
            *n_err_out += n_err;
        }
        E.flush();
    }
    // ---- pass 2 (src/dwgsim.c:366-406): one trial per empty flow; the flow pointer continues from pass 1 ------------
    {
        const FlowRow T = S; S = D; D = T;                      // read what pass 1 wrote
        // pend: symbols on a stack in the free tail of the SOURCE row (top at S[cap - pend]); it never reaches the unread
        // symbols because pend <= len - len_after_pass_1 + 1 and len < cap
        int pend = 0;
        FlowEmit E;
        E.begin(D);
        FlowSrc in;
        in.seek(S, 0);
        while (E.n < len) {
            uint32_t c;
            if (pend) { c = fm_get(S, cap - pend); --pend; } else c = in.next();
            const int d = nd[flow_i * 4 + (int)c];              // empty flows before the base's flow: d trials
            if (rng.fails(d)) { flow_i += d; if (flow_i >= fl) flow_i -= fl; E.put(c); continue; }
            // some flow on the way may fire: flow by flow.  An insertion goes in FRONT of the current symbol (and of what
            // earlier flows of this symbol inserted); the reference's loop then visits the inserted symbols like any other
            bool inserted = false;
            while ((int)c != fo[flow_i]) {
                int n_err = 0;
                while (rng.coin()) ++n_err;
                if (n_err > 0 && !fm_mask_get(mask, flow_i)) {
                    if (len + n_err >= cap) { *overflow = 1; return len; }
                    if (!inserted) { ++pend; fm_set(S, cap - pend, c); inserted = true; }
                    for (int k = 0; k < n_err; ++k) { ++pend; fm_set(S, cap - pend, (uint32_t)fo[flow_i]); }
                    len += n_err;
                    *n_err_out += n_err;
                }
                flow_i = flow_i + 1 == fl ? 0 : flow_i + 1;
            }
            if (inserted) { c = fm_get(S, cap - pend); --pend; }
            E.put(c);
        }
        E.flush();
    }
    // the result is in D: A for a forward read (A -> B -> A), B for a reversed one (A -> B reversed -> A -> B)
    if (strand) fm_reverse_into(A, D, len);
    return len;
}

}  // namespace dwg
