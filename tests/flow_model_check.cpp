// tests/flow_model_check.cpp -- CPU check of dwgsim_b200/csrc/flow_model.h (the streaming form of the Ion Torrent flow
// model the device runs) against the oracle's restatement of generate_errors_flows, on random reads, flow orders and
// error rates, both driven by the same Philox FLOW draws.  Built and run by tests/test_flow_model.py (g++, liboracle.so).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../dwgsim_b200/csrc/flow_model.h"
extern "C" {
#include "../oracle/dwgsim_oracle.h"
uint32_t orc_philox_draw(int32_t seed, uint64_t gidx, uint32_t attempt, uint32_t stream, uint32_t end, uint32_t idx);
int32_t orc_generate_errors_flows_impl(const orc_opt_t *opt, uint8_t *seq, int32_t cap, uint8_t *mask, int32_t len, int32_t strand,
                                       int32_t *n_err_out, int (*coin)(void *), double (*unif)(void *), void *rng, int *overflow);
}

struct Draw {                                   // sequential words of the FLOW (gaps) and FLOWU (uniforms) streams of one (pair, end)
    int32_t seed; uint64_t gidx; uint32_t end, next, unext;
    uint32_t gap_word() { return orc_philox_draw(seed, gidx, 0, 5 /* ST_FLOW */, end, next++); }
    uint32_t unif_word() { return orc_philox_draw(seed, gidx, 0, 3 /* ST_FLOWU */, end, unext++); }
};
struct OrcRng {                                 // the oracle's flow_coin / flow_unif (oracle/dwgsim_oracle.c), restated
    Draw d; const uint32_t *gap; int left, succ, need;
};
static int orc_coin(void *p)
{
    OrcRng *r = (OrcRng *)p;
    for (;;) {
        if (r->need) { const int g = dwg::fm_rank(r->gap, ORC_FLOW_GAP_N, r->d.gap_word()); r->left = g; r->succ = g < ORC_FLOW_GAP_N; r->need = 0; }
        if (r->left > 0) { r->left--; return 0; }
        r->need = 1;
        if (r->succ) return 1;
    }
}
static double orc_unif(void *p) { return ldexp((double)((OrcRng *)p)->d.unif_word(), -32); }

static uint64_t rs = 88172645463325252ull;
static uint32_t xr() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char **argv)
{
    const long n_cases = argc > 1 ? atol(argv[1]) : 200000;
    long n_events = 0, n_ovf = 0, n_dot = 0;
    for (long t = 0; t < n_cases; ++t) {
        orc_opt_t o;
        orc_opt_init(&o);
        // flow order: a random permutation-rich string of length 4..40 that contains every base
        const int fl = 4 + (int)(xr() % 37);
        for (int i = 0; i < fl; ++i) o.flow_order[i] = (int8_t)(i < 4 ? i : xr() & 3);
        for (int i = fl - 1; i > 0; --i) { int j = (int)(xr() % (i + 1)); int8_t x = o.flow_order[i]; o.flow_order[i] = o.flow_order[j]; o.flow_order[j] = x; }
        o.flow_order_len = fl;
        const double rates[] = {0.0, 0.001, 0.01, 0.03, 0.1, 0.3, 0.5};
        const double e = rates[xr() % 7];
        std::vector<uint32_t> gap(ORC_FLOW_GAP_N);
        for (int j = 0; j < ORC_FLOW_GAP_N; ++j) {
            double v = e > 0 ? ceil((1.0 - pow(1.0 - e, (double)(j + 1))) * 4294967296.0) : 0.0;
            gap[j] = v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
        }
        const int len = 1 + (int)(xr() % (xr() % 4 == 0 ? 12 : 200));
        const int cap = 2 * len + 64;
        const int strand = (int)(xr() & 1);
        std::vector<uint8_t> seq((size_t)cap + 8, 0), mask((size_t)fl, 0);
        const int hp_bias = (int)(xr() % 3);                       // homopolymer-rich reads exercise the deletion paths
        for (int i = 0; i < len; ++i) {
            seq[i] = (uint8_t)((i > 0 && hp_bias && (xr() % 3) < (uint32_t)hp_bias) ? seq[i - 1] : (xr() % 41 == 0 ? 4 : xr() & 3));
        }
        const int nw = (cap + 7) / 8 + 1;
        const int stride = (t & 1) ? 32 : 1;                       // plain rows and the interleaved layout of the device
        std::vector<uint32_t> Av((size_t)nw * stride, 0), Bv((size_t)nw * stride, 0xDEADBEEFu), m((size_t)((fl + 31) / 32), 0xFFFFFFFFu);
        dwg::FlowRow A{Av.data() + (stride > 1 ? 5 : 0), stride}, B{Bv.data() + (stride > 1 ? 17 : 0), stride};
        for (int i = 0; i < len; ++i) A[i >> 3] |= (uint32_t)seq[i] << ((i & 7) * 4);
        std::vector<uint16_t> nd((size_t)fl * 4);
        for (int f = 0; f < fl; ++f) for (int b = 0; b < 4; ++b) nd[(size_t)f * 4 + b] = (uint16_t)dwg::fm_build_nd_entry(o.flow_order, fl, f, b);
        const int32_t seed = (int32_t)xr();
        const uint64_t gidx = ((uint64_t)xr() << 20) ^ xr();
        // oracle
        OrcRng r{{seed, gidx, 1u, 0u, 0u}, gap.data(), 0, 0, 1};
        int32_t nerr_o = 0; int ovf_o = 0;
        const int len_o = orc_generate_errors_flows_impl(&o, seq.data(), cap, mask.data(), len, strand, &nerr_o, orc_coin, orc_unif, &r, &ovf_o);
        // streaming form
        // every third case hands the first gaps over already ranked, like the device does
        const int n_ahead = (t % 3 == 0) ? (int)(xr() % 9) : 0;
        uint16_t ahead[8];
        Draw d0{seed, gidx, 1u, 0u, 0u};
        for (int i = 0; i < n_ahead; ++i) ahead[i] = (uint16_t)dwg::fm_rank(gap.data(), ORC_FLOW_GAP_N, d0.gap_word());
        dwg::FlowCoin<Draw> rng(d0, gap.data(), ahead, n_ahead);
        int nerr_s = 0, ovf_s = 0;
        const int len_s = dwg::flow_model_rows(A, B, len, cap, strand, o.flow_order, fl, nd.data(), m.data(), rng, &nerr_s, &ovf_s);
        bool ok = (ovf_o != 0) == (ovf_s != 0);
        if (ok && !ovf_o) {
            ok = len_o == len_s && nerr_o == nerr_s;
            for (int i = 0; ok && i < len_o; ++i) ok = (uint32_t)(seq[i] >= 4 ? 0 : seq[i]) == ((A[i >> 3] >> ((i & 7) * 4)) & 15u);
            ok = ok && r.d.unext == rng.draw.unext && (n_ahead ? r.d.next <= rng.draw.next : r.d.next == rng.draw.next);   // same draws consumed
        }
        if (!ok) {
            fprintf(stderr, "MISMATCH case %ld: fl %d e %g len %d strand %d | oracle len %d nerr %d ovf %d draws %u | stream len %d nerr %d ovf %d draws %u\n",
                    t, fl, e, len, strand, len_o, nerr_o, ovf_o, r.d.next, len_s, nerr_s, ovf_s, rng.draw.next);
            return 1;
        }
        n_events += nerr_o; n_ovf += ovf_o != 0;
        (void)n_dot;
    }
    printf("flow_model_check: %ld cases ok (%ld errors applied, %ld overflows)\n", n_cases, n_events, n_ovf);
    return 0;
}
