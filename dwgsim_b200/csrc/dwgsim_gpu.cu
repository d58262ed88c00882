// dwgsim_gpu.cu -- C ABI of libdwgsim_b200.so (include/dwgsim_gpu.h): host packer, derived tables,
// batch pipeline (compute stream + copy stream + pinned ring) around the kernels in kernels.cuh.
// There is no CPU fallback: every entry point that needs a device fails with DWGSIM_GPU_ENODEV / ECUDA.
#include <cuda_runtime.h>
#include <errno.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dwgsim_gpu.h"
#include "kernels.cuh"
#include "gz_host.h"
#include "gz_device.cuh"

using namespace dwg;

namespace {

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
inline double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- symbol table (reference: nst_nt4_table, src/dwgsim.c:56-73); '-' (code 5 there) is folded into N ----
struct Nt4 {
    uint8_t t[256];
    Nt4()
    {
        memset(t, 4, sizeof t);
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
        t['-'] = 5;
    }
};
const Nt4 kNt4;

struct HostContig {
    std::string name;
    int32_t contig_i = 0, len = 0;
    int64_t n_pairs = 0;
    std::vector<uint32_t> ref2, nmask;
    std::vector<Event> ev[2];
    std::vector<uint32_t> blk[2];
    std::vector<uint8_t> pool[2];
    std::vector<Region> regions;               // -x
    int32_t sample_len = 0;
};

struct DeviceTables {
    uint32_t *isize_cdf = nullptr, *qdelta_cdf = nullptr, *err_gap[2] = {nullptr, nullptr}, *err_acc[2] = {nullptr, nullptr};
    uint16_t *isize_guide = nullptr, *gap_guide[2] = {nullptr, nullptr};
    uint32_t *qguide = nullptr;
    uint8_t *qtab = nullptr;
    uint8_t *qbase[2] = {nullptr, nullptr};
    int8_t *flow_order = nullptr;
    uint32_t *flow_gap[2] = {nullptr, nullptr};
    char *prefix = nullptr;
};

struct Workspace {
    int64_t cap_pairs = 0;
    PairRec *recs = nullptr;
    uint32_t *seqs = nullptr;                 // nibble-packed read codes, word-major [nw0 + nw1][n]
    unsigned long long *serial = nullptr;
    uint32_t *lens = nullptr;                 // [3][cap]
    char *names = nullptr;                    // [cap][nvar][name_cap]
    uint16_t *name_len = nullptr;             // [cap][2]
    unsigned long long *blk_rand = nullptr;   // [nblk]
    unsigned long long *blk_len = nullptr;    // [3][nblk]
    unsigned long long *totals = nullptr;     // [8]: 0 n_random, 1..3 stream bytes
    unsigned long long *status = nullptr;     // [4]: error bits, failed attempts, sizes of the job lists F1, R1
    uint2 *jobs = nullptr;                    // [2][cap]: retry and random job lists of the simulate passes
    uint32_t *flow_scratch = nullptr;         // Ion Torrent: one scratch row per resident thread of the simulate kernel (flow_model.h)
    char *out[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    uint64_t out_cap[3] = {0, 0, 0};
    int32_t name_cap = 0;                     // the name / record bounds the buffers were sized with (update_caps raises them
    uint64_t rec_cap[3] = {0, 0, 0};          //   when a later contig has a longer name)
    unsigned long long *h_totals = nullptr;   // pinned, mapped [16]: batch totals the kernels publish straight to the host
    unsigned long long *h_totals_dev = nullptr;   // its device alias (the copy engines stay free for the FASTQ streams)
    // device gzip writer
    uint8_t *gz_slots[3] = {nullptr, nullptr, nullptr};
    char *gz_out[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    uint64_t gz_cap[3] = {0, 0, 0}, gz_members[3] = {0, 0, 0};
    unsigned long long *gz_sizes = nullptr, *gz_offs = nullptr, *gz_totals = nullptr, *gz_hist = nullptr;   // [3][members], .., [3], [256]
    uint64_t gz_members_max = 0;
};

}  // namespace
struct dwgsim_gpu;
namespace {
// Several devices behind one handle (dwgsim_gpu_create_group): the ranks of a run are threads of this process.  They meet
// twice per round: to add up their random-pair counts (rand_ii is a running count over all pairs, src/dwgsim.c:1096) and
// to hand their batches to the sink in batch order.
struct GroupSync {
    std::mutex mu;
    std::condition_variable cv;
    int world = 1;
    std::vector<int64_t> vals;         // this round's random-pair count of every rank
    int arrived = 0;
    int64_t generation = 0;
    std::vector<int64_t> before;       // results of the last completed exchange
    int64_t total = 0;
    int64_t turn = 0;                  // next batch index the sink may see
    bool failed = false;               // some rank gave up: nobody waits any more
    dwgsim_gpu *leader = nullptr;
    void reset(int w) { world = w; vals.assign((size_t)w, 0); before.assign((size_t)w, 0); arrived = 0; total = 0; turn = 0; failed = false; }
    void fail() { std::lock_guard<std::mutex> g(mu); failed = true; cv.notify_all(); }
    // all-gather + exclusive prefix of one value per rank; false when the run was abandoned
    bool exchange(int rank, int64_t mine, int64_t *before_me, int64_t *round_total)
    {
        std::unique_lock<std::mutex> lk(mu);
        if (failed) return false;
        vals[(size_t)rank] = mine;
        const int64_t gen = generation;
        if (++arrived == world) {
            int64_t acc = 0;
            for (int r = 0; r < world; ++r) { before[(size_t)r] = acc; acc += vals[(size_t)r]; }
            total = acc; arrived = 0; ++generation;
            cv.notify_all();
        } else cv.wait(lk, [&]() { return generation != gen || failed; });
        if (failed && generation == gen) return false;
        *before_me = before[(size_t)rank]; *round_total = total;
        return true;
    }
    bool wait_turn(int64_t batch)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return turn == batch || failed; });
        return !failed;
    }
    void pass_turn() { std::lock_guard<std::mutex> g(mu); ++turn; cv.notify_all(); }
};

}  // namespace

struct dwgsim_gpu {
    dwgsim_gpu_params_t p{};
    // device group (this handle is the leader, rank 0; peers are ranks 1..)
    std::vector<dwgsim_gpu *> peers;
    GroupSync *group = nullptr;               // shared by the leader and its peers while a run is in flight
    std::string prefix_s;                     // "pfx_" or ""
    std::vector<int8_t> flow_order;
    int device = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_t[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string last_error;
    // derived tables (host + device)
    std::vector<uint32_t> isize_cdf, qdelta_cdf, err_gap[2], err_acc[2];
    std::vector<uint16_t> isize_guide, gap_guide[2];
    std::vector<uint32_t> qguide;
    std::vector<uint8_t> qtab;
    std::vector<uint32_t> flow_gap[2];
    bool ion_warp_kernel = false;
    // device gzip writer: mode, per-stream tables in HBM
    int gz_mode = 0;
    double ms_gz = 0;
    bool gz_ready = false;
    uint64_t gz_hist_host[3][256] = {};       // byte histograms the Huffman codes were fitted to (group leader: shared with the peers)
    bool gz_hist_valid = false;
    uint32_t *gz_code[3] = {nullptr, nullptr, nullptr};
    uint8_t *gz_prefix[3] = {nullptr, nullptr, nullptr};
    uint32_t gz_prefix_bits[3] = {0, 0, 0};
    uint32_t *gz_crc = nullptr;             // Ion Torrent reads too long for the thread-per-pair rows: warp-per-pair kernel
    std::vector<uint8_t> qbase[2];
    uint64_t thr_genomic = 0, thr_hap0 = 0;
    int32_t isize_lo = 0, qdelta_lo = 0;
    uint32_t flow_thr[2] = {0, 0};
    DeviceTables dt;
    SimParams sp{};
    // queue / genome
    std::vector<HostContig> queue;
    uint8_t *blob = nullptr;
    uint64_t blob_bytes = 0;
    bool blob_owned = true;
    uint8_t *blob_spare = nullptr;            // the allocation of the previous run's genome, reused when large enough
    uint64_t blob_spare_cap = 0;
    int64_t blob_pairs = 0;
    int max_name_len = 4;
    int host_threads = 0;                     // packer threads (0: min(hardware threads, 32))
    double ms_pack = 0;
    int64_t h2d_bytes = 0;
    // global counters carried across runs (ctr / rand_ii of src/dwgsim.c:423)
    int64_t gidx_origin = 0, rand_serial = 0;
    // batching
    int64_t batch_pairs = 1 << 17;
    int ring = 3;
    int shard_rank = 0, shard_world = 1;
    dwgsim_gpu_exchange_fn exchange = nullptr;
    void *exchange_user = nullptr;
    int64_t pending_first = -1; int pending_n = 0, pending_launches = 0;
    // batches queued without a host sync (dwgsim_gpu_resident_enqueue / _finish_async)
    unsigned long long *queue_dev = nullptr;   // [0] running count of random pairs, [1] error bits since the last wait,
                                               // [2] count before this rank's batch (dwgsim_gpu_resident_finish_gathered)
    bool queue_active = false;                 // the batch being launched belongs to the queue; advance [0] with it?
    bool queue_advance = false;
    int queued_launches = 0, queued_n = 0;
    Workspace ws;
    char *pinned[8][3] = {};
    uint64_t pinned_cap[3] = {0, 0, 0};
    int pinned_slots = 0;
    // last resident batch
    int last_slot = 0;
    uint64_t last_bytes[3] = {0, 0, 0};
};

namespace {

// the format kernel is specialised on colour space and on whether the quality sum can wrap in int8
typedef void (*format_kernel_t)(const SimParams, const uint8_t *, int64_t, int64_t, int, const PairRec *, const uint32_t *,
                                const unsigned long long *, const uint32_t *, const unsigned long long *, const char *,
                                const uint16_t *, char *, char *, char *);
format_kernel_t format_kernel_of(const SimParams &sp)
{
    if (sp.data_type == 1) return sp.q_wrap ? format_fastq_kernel<true, true> : format_fastq_kernel<true, false>;
    return sp.q_wrap ? format_fastq_kernel<false, true> : format_fastq_kernel<false, false>;
}

// dynamic shared memory of simulate_pairs_tp_kernel: staging tile + sampling tables (see the kernel prologue)
// the instance of the word-granular format kernel for a configuration: colour space x quality mode
using format2_kernel_t = void (*)(const SimParams, const Format2Smem, int64_t, int64_t, int, const PairRec *, const uint32_t *, const uint32_t *,
                                  const unsigned long long *, const char *, const uint16_t *, char *, char *, char *);
template <bool kSolid, int kQMode>
format2_kernel_t format2_kernel_out(int out)
{
    return out == 1 ? format_fastq2_kernel<kSolid, kQMode, 1> : (out == 2 ? format_fastq2_kernel<kSolid, kQMode, 2> : format_fastq2_kernel<kSolid, kQMode, 3>);
}
format2_kernel_t format2_kernel_of(const SimParams &sp)
{
    const int qmode = sp.fixed_quality ? 2 : (sp.qdelta_n > 0 ? 1 : 0);       // 0: no noise, 1: noise table, 2: fixed character
    const int out = (sp.out_bwa ? 1 : 0) | (sp.out_bfast ? 2 : 0);
    if (sp.data_type == 1)
        return qmode == 2 ? format2_kernel_out<true, 2>(out) : (qmode == 1 ? format2_kernel_out<true, 1>(out) : format2_kernel_out<true, 0>(out));
    return qmode == 2 ? format2_kernel_out<false, 2>(out) : (qmode == 1 ? format2_kernel_out<false, 1>(out) : format2_kernel_out<false, 0>(out));
}

// the instance of the thread-per-pair simulate kernel for a configuration (SimParams.tp_tables is 3 or 0)
using tp_kernel_t = void (*)(const SimParams, const uint8_t *, int64_t, int64_t, int, int, JobLists, PairRec *, uint32_t *, unsigned long long *, uint32_t *);
tp_kernel_t tp_kernel_of(const SimParams &sp)
{
    if (sp.data_type == 2) return simulate_pairs_tp_kernel<true, 3>;
    return sp.tp_tables >= 3 ? simulate_pairs_tp_kernel<false, 3> : simulate_pairs_tp_kernel<false, 0>;
}

size_t tp_smem_bytes(const SimParams &sp)
{
    size_t words = (((size_t)kTpThreads * sp.row_stride + 3) & ~(size_t)3) +
                   (sp.isize_n <= kIsizeSmemMax && sp.tp_tables >= 1 ? ((sp.isize_n + 1) & ~1) : 0);
    words += 2 * (size_t)sp.win_slots * kTpThreads;                                                 // the reference window
    if (sp.data_type != 2 && sp.tp_tables >= 2) for (int e = 0; e < 2; ++e) words += 2 * (size_t)((sp.len[e] + 1) & ~1);   // (not used by the flow model)
    const size_t flow = sp.data_type != 2 ? 0 :                                                      // Ion Torrent only:
                        (size_t)((sp.flow_order_len + 15) & ~15) + (size_t)kTpThreads * ((sp.flow_order_len + 31) >> 5) * 4 +
                        (size_t)sp.flow_order_len * 8 + 32 + (size_t)kTpThreads * kFlowGapsAhead * 2;  // flow order, masks, nd table, gaps drawn ahead
    return words * 4 + (sp.tp_tables >= 3 ? 3 * 1026 * 2 : 0) + flow + 32;
}

#define CUDA_TRY(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            (h)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                       \
            return _e == cudaErrorMemoryAllocation ? DWGSIM_GPU_ENOMEM : DWGSIM_GPU_ECUDA;              \
        }                                                                                              \
    } while (0)

// ---- derived tables: every probability becomes a 32-bit threshold (DESIGN.md "RNG addressing") ----------
double phi(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
uint32_t thr32(double p)
{
    if (!(p > 0.0)) return 0;
    double v = ceil(p * 4294967296.0);
    return v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
}
uint64_t thr64(double p)
{
    if (!(p > 0.0)) return 0;
    double v = ceil(p * 4294967296.0);
    return v >= 4294967296.0 ? 4294967296ull : (uint64_t)v;
}

void derive_tables(dwgsim_gpu *h)
{
    const dwgsim_gpu_params_t &p = h->p;
    // genomic iff rand_read < U (src/dwgsim.c:649)
    double v = floor(p.rand_read * 4294967296.0);
    h->thr_genomic = v >= 4294967296.0 ? 4294967297ull : (uint64_t)v + 1ull;
    h->thr_hap0 = thr64(p.mut_freq);                                        // src/dwgsim.c:716
    // insert size d = (int)(N(0,1)*std + dist + 0.5) (src/dwgsim.c:657-659): inverse-CDF table
    h->isize_cdf.clear();
    if (p.std_dev > 0.0) {
        int span = (int)ceil(8.0 * p.std_dev) + 1;
        h->isize_lo = p.dist - span;
        for (int j = 0; j < 2 * span; ++j)
            h->isize_cdf.push_back(thr32(phi(((double)(h->isize_lo + j) + 0.5 - (double)p.dist) / p.std_dev)));
    } else h->isize_lo = p.dist;
    // quality noise (int)(N(0,1)*qstd + 0.5), truncation toward zero (src/dwgsim.c:911-913)
    h->qdelta_cdf.clear();
    h->qdelta_lo = 0;
    if (p.quality_std > 0.0 && p.fixed_quality == 0) {
        int span = (int)ceil(8.0 * p.quality_std) + 1;
        if (span > 32768) span = 32768;
        h->qdelta_lo = -span;
        for (int j = 0; j < 2 * span; ++j) {
            int k = h->qdelta_lo + j;
            double x = k < 0 ? (double)k - 0.5 : (double)k + 0.5;
            h->qdelta_cdf.push_back(thr32(phi(x / p.quality_std)));
        }
    }
    // acceleration only (not part of the sampling rule): one-load guide, 1024 buckets of 2^22.  Entry g = {t, r}: r = rank
    // at the bucket's lower bound; when exactly one threshold c lies inside the bucket t = c - 1 (rank = r + (u > t)),
    // with none t = 2^32 - 1; with several (the tails) r carries bit 31, t = their number and the kernel scans the CDF
    h->qguide.assign(2 * 1024, 0);
    for (uint32_t g = 0; g < 1024; ++g) {
        const std::vector<uint32_t> &cdf = h->qdelta_cdf;
        const uint64_t lo = (uint64_t)g << 22, hi = lo + (1ull << 22);
        const size_t j = (size_t)(std::upper_bound(cdf.begin(), cdf.end(), (uint32_t)lo) - cdf.begin());
        size_t inside = 0;
        while (j + inside < cdf.size() && (uint64_t)cdf[j + inside] < hi) ++inside;
        h->qguide[2 * g] = inside == 1 ? cdf[j] - 1u : (inside > 1 ? (uint32_t)inside : 0xFFFFFFFFu);
        h->qguide[2 * g + 1] = (uint32_t)j | (inside > 1 ? 0x80000000u : 0u);
    }
    // format_fastq2_kernel: rank by the upper kQTabBits of a draw; cells that hold a threshold are marked (bit 7) and
    // decided from the whole 32-bit draw
    h->qtab.assign((size_t)1 << kQTabBits, 0);
    if (!h->qdelta_cdf.empty() && h->qdelta_cdf.size() < 128) {
        const std::vector<uint32_t> &cdf = h->qdelta_cdf;
        for (uint32_t c = 0; c < (1u << kQTabBits); ++c) {
            const uint32_t lo = c << (32 - kQTabBits), hi = lo + ((1u << (32 - kQTabBits)) - 1u);
            const size_t r_lo = (size_t)(std::upper_bound(cdf.begin(), cdf.end(), lo) - cdf.begin());
            const size_t r_hi = (size_t)(std::upper_bound(cdf.begin(), cdf.end(), hi) - cdf.begin());
            h->qtab[c] = r_lo == r_hi ? (uint8_t)r_lo : (uint8_t)0x80;
        }
    }
    auto make_guide = [](const uint32_t *cdf, size_t n, std::vector<uint16_t> &g) {     // g[b] = rank of (b << 22), b = 0..1024
        g.assign(1025, 0);
        for (uint32_t b = 0; b < 1024; ++b) g[b] = (uint16_t)(std::upper_bound(cdf, cdf + n, b << 22) - cdf);
        g[1024] = (uint16_t)n;
    };
    make_guide(h->isize_cdf.data(), h->isize_cdf.size(), h->isize_guide);
    for (int e = 0; e < 2; ++e) {
        int n = p.length[e];
        if (p.data_type == 2) n = 2 * n + 64;
        h->err_gap[e].assign((size_t)n + 1, 0);
        h->err_acc[e].assign((size_t)n + 1, 0);
        h->qbase[e].assign((size_t)n + 1, 0);
        // per-cycle Bernoulli(start + by*i) errors (src/dwgsim.c:237) by thinning: candidates arrive with geometric
        // gaps at the largest per-cycle rate, a candidate at cycle i is kept with probability p_i / pmax
        double pmax = 0.0;
        for (int j = 0; j < p.length[e]; ++j) pmax = std::max(pmax, p.e_start[e] + p.e_by[e] * j);
        if (pmax > 1.0) pmax = 1.0;
        for (int j = 0; j < n; ++j) {
            double pr = p.e_start[e] + p.e_by[e] * j;                       // src/dwgsim.c:237, :906-910
            h->err_gap[e][j] = thr32(1.0 - pow(1.0 - pmax, (double)(j + 1)));
            h->err_acc[e][j] = (pmax > 0.0 && pr > 0.0) ? thr32(pr / pmax) : 0;
            h->qbase[e][j] = pr > 0 ? (uint8_t)(int)(-10.0 * log(pr) / log(10.0) + 0.499) : 40;
        }
        make_guide(h->err_gap[e].data(), (size_t)p.length[e], h->gap_guide[e]);
        h->flow_thr[e] = thr32(p.e_start[e]);
        // Ion Torrent: the per-flow error coin (src/dwgsim.c:290,372) as the geometric gaps of its Bernoulli process
        h->flow_gap[e].assign(p.data_type == 2 ? (size_t)kFlowGapN : 1, 0);
        if (p.data_type == 2 && p.e_start[e] > 0.0)
            for (int j = 0; j < kFlowGapN; ++j) h->flow_gap[e][j] = thr32(1.0 - pow(1.0 - std::min(p.e_start[e], 1.0), (double)(j + 1)));
    }
}

template <typename T>
int upload(dwgsim_gpu *h, T **dst, const T *src, size_t n)
{
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CUDA_TRY(h, cudaMalloc((void **)dst, bytes));
    if (n) CUDA_TRY(h, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return DWGSIM_GPU_OK;
}

int upload_tables(dwgsim_gpu *h)
{
    int rc;
    if ((rc = upload(h, &h->dt.isize_cdf, h->isize_cdf.data(), h->isize_cdf.size()))) return rc;
    if ((rc = upload(h, &h->dt.qdelta_cdf, h->qdelta_cdf.data(), h->qdelta_cdf.size()))) return rc;
    if ((rc = upload(h, &h->dt.qguide, h->qguide.data(), h->qguide.size()))) return rc;
    if ((rc = upload(h, &h->dt.qtab, h->qtab.data(), h->qtab.size()))) return rc;
    if ((rc = upload(h, &h->dt.isize_guide, h->isize_guide.data(), h->isize_guide.size()))) return rc;
    for (int e = 0; e < 2; ++e) if ((rc = upload(h, &h->dt.gap_guide[e], h->gap_guide[e].data(), h->gap_guide[e].size()))) return rc;
    for (int e = 0; e < 2; ++e) {
        if ((rc = upload(h, &h->dt.err_gap[e], h->err_gap[e].data(), h->err_gap[e].size()))) return rc;
        if ((rc = upload(h, &h->dt.err_acc[e], h->err_acc[e].data(), h->err_acc[e].size()))) return rc;
        if ((rc = upload(h, &h->dt.qbase[e], h->qbase[e].data(), h->qbase[e].size()))) return rc;
    }
    if ((rc = upload(h, &h->dt.flow_order, h->flow_order.data(), h->flow_order.size()))) return rc;
    for (int e = 0; e < 2; ++e) if ((rc = upload(h, &h->dt.flow_gap[e], h->flow_gap[e].data(), h->flow_gap[e].size()))) return rc;
    if ((rc = upload(h, &h->dt.prefix, h->prefix_s.data(), h->prefix_s.size()))) return rc;
    const dwgsim_gpu_params_t &p = h->p;
    SimParams &s = h->sp;
    memset(&s, 0, sizeof s);
    for (int e = 0; e < 2; ++e) {
        s.len[e] = p.length[e];
        s.cap[e] = p.data_type == 2 ? 2 * p.length[e] + 64 : p.length[e];
        s.err_gap[e] = h->dt.err_gap[e]; s.err_acc[e] = h->dt.err_acc[e];
        s.nw[e] = (s.cap[e] + 7) / 8;
        s.qbase[e] = h->dt.qbase[e];
        s.flow_thr[e] = h->flow_thr[e];
    }
    s.is_inner = p.is_inner; s.max_n = p.max_n; s.data_type = p.data_type; s.strandedness = p.strandedness;
    s.read_one_strand = p.read_one_strand; s.amplicons = p.amplicons;
    s.seed = (uint32_t)p.seed;
    s.thr_genomic = h->thr_genomic; s.thr_hap0 = h->thr_hap0;
    s.isize_lo = h->isize_lo; s.isize_n = (int32_t)h->isize_cdf.size();
    s.qdelta_lo = h->qdelta_lo; s.qdelta_n = (int32_t)h->qdelta_cdf.size();
    // the reference adds the noise in `char` (src/dwgsim.c:911-913): only a wide noise table can leave the int8 range
    s.q_wrap = (73 + (int)h->qdelta_cdf.size() / 2 > 127 || 33 + h->qdelta_lo < -128) ? 1 : 0;
    s.fixed_quality = p.fixed_quality;
    s.out_bwa = p.reads_output_type != 2; s.out_bfast = p.reads_output_type != 1;
    s.prefix_len = (int32_t)h->prefix_s.size();
    s.flow_order_len = p.flow_order_len;
    s.tile_pairs = 4;
    s.isize_cdf = h->dt.isize_cdf; s.qdelta_cdf = h->dt.qdelta_cdf; s.qguide = h->dt.qguide; s.qtab = h->dt.qtab;
    for (int r = 0; r < 10; ++r) s.qkey[r] = s.seed + (uint32_t)r * 0x9E3779B9u;
    s.fmt_v2 = (!s.q_wrap && h->qdelta_cdf.size() < 128) ? 1 : 0;
    if (const char *e = getenv("DWGSIM_FORMAT")) if (atoi(e) == 1) s.fmt_v2 = 0;                  // 1: the byte-granular kernel everywhere
    s.isize_guide = h->dt.isize_guide; s.gap_guide[0] = h->dt.gap_guide[0]; s.gap_guide[1] = h->dt.gap_guide[1];
    s.inv_nw = (uint32_t)(4294967296.0 / std::max(s.nw[0] + s.nw[1], 1)) + 1u;
    {   // staged rows: unpadded when lanes then collide two ways at most (stride = 2 mod 4 words); Ion Torrent rows are edited
        // in place word by word and keep the conflict-free odd stride
        const int nw = s.nw[0] + s.nw[1];
        s.row_stride = (p.data_type != 2 && (nw & 3) == 2) ? nw : (nw | 1);
        if (const char *e = getenv("DWGSIM_ROW_PAD")) if (atoi(e) && p.data_type != 2) s.row_stride = nw | 1;
        if (p.data_type == 2) s.row_stride = 0;                 // Ion Torrent rows live in HBM / L2 (Workspace::flow_scratch)
    }
    s.inv_groups = (uint32_t)(4294967296.0 / std::max((s.cap[0] + 7) / 8 + (s.cap[1] + 7) / 8, 1)) + 1u;
    if (s.fmt_v2)                                              // format_fastq2_kernel: 16 bases per lane
        s.inv_groups = (uint32_t)(4294967296.0 / std::max(((s.cap[0] + 7) / 8 + 1) / 2 + ((s.cap[1] + 7) / 8 + 1) / 2, 1)) + 1u;
    s.flow_order = h->dt.flow_order; s.prefix = h->dt.prefix;
    s.flow_gap[0] = h->dt.flow_gap[0]; s.flow_gap[1] = h->dt.flow_gap[1];
    return DWGSIM_GPU_OK;
}

// ---- dense (seq_t + 2 x mutseq_t) -> packed sections ---------------------------------------------------
// long insertion record of the reference: [tag 1|2|4][length][2-bit bases, first inserted base stored last]
// (src/mut.c:200-246, :345-365)
const uint8_t *long_ins_payload(const uint8_t *rec, uint32_t *n)
{
    if (rec[0] == 1) { *n = rec[1]; return rec + 2; }
    if (rec[0] == 2) { uint16_t v; memcpy(&v, rec + 1, 2); *n = v; return rec + 3; }
    uint32_t v; memcpy(&v, rec + 1, 4); *n = v; return rec + 5;
}

// one worker's share of a contig: positions [p0, p1), p0 a multiple of 128 so no output word is shared
struct PackPart {
    std::vector<Event> ev[2];
    std::vector<uint8_t> pool[2];      // 2-bit packed, part-local base offsets
    uint64_t pool_bases[2] = {0, 0};
    int rc = DWGSIM_GPU_OK;
    const char *err = nullptr;
};

void pack_range(HostContig &c, const uint8_t *seq, const uint64_t *const hap[2], uint8_t *const *const ins[2],
                const int32_t ins_n[2], int p0, int p1, PackPart &out)
{
    for (int w0 = p0; w0 < p1; w0 += 32) {
        const int wend = std::min(w0 + 32, p1);
        uint32_t lo = 0, hi = 0, nm = 0;
        for (int p = w0; p < wend; ++p) {
            const uint8_t raw = kNt4.t[seq[p]];
            const int b = p - w0;
            if (raw < 4) { if (b < 16) lo |= (uint32_t)raw << (b << 1); else hi |= (uint32_t)raw << ((b - 16) << 1); }
            else nm |= 1u << b;
            for (int hh = 0; hh < 2; ++hh) {
                const uint64_t m = hap[hh][p];
                if (m == (uint64_t)raw) continue;
                Event e;
                e.pos = (uint32_t)p;
                uint32_t type = (uint32_t)(m >> 4) & 3u, base = (uint32_t)(m & 0xf);
                if (base > 4) base = 4;
                uint32_t n = 0;
                e.payload = 0;
                if (type == kEvInsert) {
                    n = (uint32_t)(m >> 59) & 0x1fu;
                    if (n) e.payload = (m >> 6) & ((1ull << 52) - 1);
                    else {
                        const uint64_t idx = (m >> 6) & ((1ull << 52) - 1);
                        if ((int64_t)idx >= ins_n[hh]) { out.rc = DWGSIM_GPU_EINVAL; out.err = "long insertion index out of range"; return; }
                        const uint8_t *pl = long_ins_payload(ins[hh][idx], &n);
                        if (n >= (1u << 27)) { out.rc = DWGSIM_GPU_EUNSUPPORTED; out.err = "insertion longer than 2^27-1 bases"; return; }
                        auto src = [&](uint32_t j) { uint32_t at = n - 1 - j; return (uint32_t)(pl[at >> 2] >> ((at & 3) << 1)) & 3u; };
                        if (n <= kInlineInsMax) {
                            for (uint32_t j = 0; j < n; ++j) e.payload |= (uint64_t)src(j) << (2 * j);
                        } else {
                            // pool entries start on a byte boundary so parts can be concatenated
                            out.pool_bases[hh] = (out.pool_bases[hh] + 3) & ~3ull;
                            e.payload = out.pool_bases[hh];
                            out.pool[hh].resize((size_t)((out.pool_bases[hh] + n + 3) >> 2), 0);
                            for (uint32_t j = 0; j < n; ++j) {
                                uint64_t at = out.pool_bases[hh] + j;
                                out.pool[hh][at >> 2] |= (uint8_t)(src(j) << ((at & 3) << 1));
                            }
                            out.pool_bases[hh] += n;
                        }
                    }
                }
                e.meta = type | (base << 2) | (n << 5);
                out.ev[hh].push_back(e);
            }
        }
        c.ref2[w0 >> 4] = lo;
        if (w0 + 16 < p1 || hi) c.ref2[(w0 >> 4) + 1] = hi;
        c.nmask[w0 >> 5] = nm;
    }
}

int pack_contig(int host_threads, std::string &err, HostContig &c, const uint8_t *seq, const uint64_t *hap[2], uint8_t *const *ins[2],
                const int32_t ins_n[2])
{
    const int len = c.len;
    c.ref2.assign(((size_t)len + 15) / 16 + 2, 0);
    c.nmask.assign(((size_t)len + 31) / 32 + 1, 0);
    const int nblk = (len >> kBlkShift) + 2;
    // split into 128-base aligned ranges, one per worker thread
    unsigned nt = host_threads > 0 ? (unsigned)host_threads : std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (const char *e = getenv("DWGSIM_HOST_THREADS")) if (atoi(e) > 0 && host_threads <= 0) nt = (unsigned)atoi(e);
    if (len < (1 << 20)) nt = 1;
    const int per = (int)((((int64_t)len + nt - 1) / nt + 127) & ~127ll);
    std::vector<PackPart> parts(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        const int p0 = (int)std::min<int64_t>((int64_t)t * per, len), p1 = (int)std::min<int64_t>((int64_t)(t + 1) * per, len);
        if (p0 >= p1) continue;
        if (nt == 1) pack_range(c, seq, hap, ins, ins_n, p0, p1, parts[t]);
        else th.emplace_back([&, t, p0, p1]() { pack_range(c, seq, hap, ins, ins_n, p0, p1, parts[t]); });
    }
    for (auto &x : th) x.join();
    for (int hh = 0; hh < 2; ++hh) {
        size_t total = 0;
        for (auto &pt : parts) total += pt.ev[hh].size();
        c.ev[hh].clear(); c.ev[hh].reserve(total);
        c.pool[hh].clear();
        for (auto &pt : parts) {
            if (pt.rc) { err = pt.err ? pt.err : "pack failed"; return pt.rc; }
            const uint64_t base_bases = (uint64_t)c.pool[hh].size() * 4;
            for (Event e : pt.ev[hh]) {
                if ((e.meta & 3u) == kEvInsert && (e.meta >> 5) > kInlineInsMax) e.payload += base_bases;
                c.ev[hh].push_back(e);
            }
            c.pool[hh].insert(c.pool[hh].end(), pt.pool[hh].begin(), pt.pool[hh].end());
        }
        c.blk[hh].assign((size_t)nblk, 0);
        size_t e = 0;
        for (int b = 0; b < nblk; ++b) {
            const uint64_t start = (uint64_t)b << kBlkShift;
            while (e < c.ev[hh].size() && c.ev[hh][e].pos < start) ++e;
            c.blk[hh][b] = (uint32_t)e;
        }
        c.pool[hh].resize(c.pool[hh].size() + 4, 0);
    }
    return DWGSIM_GPU_OK;
}

// upper bounds of the record sizes (src/dwgsim.c:923-978): name = '@' prefix contig 13 separators,
// 2 x 10 digits, 4 flags, 6 counts of <= 5 digits, 16 hex digits
uint64_t name_cap_of(const dwgsim_gpu *h)
{
    return 1 + h->prefix_s.size() + (uint64_t)std::max(h->max_name_len, 4) + 13 + 20 + 4 + 30 + 16;
}
void record_caps(const dwgsim_gpu *h, uint64_t cap[3], uint64_t name_slack = 0)
{
    const uint64_t name = name_cap_of(h) + name_slack;
    const SimParams &s = h->sp;
    for (int e = 0; e < 2; ++e) cap[e] = s.out_bwa && s.len[e] > 0 ? name + 3 + 2ull * s.cap[e] + 4 : 0;
    cap[2] = 0;
    if (s.out_bfast)
        for (int e = 0; e < 2; ++e) if (s.len[e] > 0) cap[2] += name + 1 + 2ull * s.cap[e] + 5;
}
// record geometry that depends on the longest contig name: kernel parameters + shared memory of the format kernel
int update_caps(dwgsim_gpu *h)
{
    uint64_t cap[3];
    record_caps(h, cap);
    h->sp.name_cap = (int32_t)((name_cap_of(h) + 15) & ~15ull);
    h->sp.inv_name_chunks = (uint32_t)(4294967296.0 / std::max(h->sp.name_cap >> 4, 1)) + 1u;
    for (int k = 0; k < 3; ++k) h->sp.rec_cap[k] = (int32_t)cap[k];
    // mini-tile of the format kernel: pairs per warp, about 150 eight-base groups (2x150: 4 pairs = 152 groups = 4.75 rounds
    // of the warp; 2x50: 11 pairs), fewer while two 16-warp CTAs do not fit an SM (<= 113 KB each); at most 31 (one lane
    // per pair + one in step 0); DWGSIM_TILE_PAIRS overrides it for experiments
    {
        const int groups = std::max((h->sp.cap[0] + 7) / 8 + (h->sp.cap[1] + 7) / 8, 1);
        h->sp.tile_pairs = std::max(1, std::min(31, (152 + groups - 1) / groups));       // ~150 groups of 8 bases per warp
    }
    if (const char *e = getenv("DWGSIM_TILE_PAIRS")) h->sp.tile_pairs = std::max(1, std::min(31, atoi(e)));
    if (h->sp.fmt_v2) {
        // one CTA per SM (the noise table takes 64 KB of its shared memory): as many warps as fit with the mini-tile wanted
        h->sp.fmt_warps = kFmt2WarpsMax;
        if (const char *e = getenv("DWGSIM_FMT_WARPS")) h->sp.fmt_warps = std::max(1, std::min(kFmt2WarpsMax, atoi(e)));
        const int warps_floor = getenv("DWGSIM_TILE_PAIRS") ? 1 : std::min(h->sp.fmt_warps, 20);
        while (h->sp.fmt_warps > warps_floor && format2_smem_layout(h->sp).total > 227 * 1024) --h->sp.fmt_warps;
        while (h->sp.tile_pairs > 1 && format2_smem_layout(h->sp).total > 227 * 1024) --h->sp.tile_pairs;
        while (h->sp.fmt_warps > 1 && format2_smem_layout(h->sp).total > 227 * 1024) --h->sp.fmt_warps;
        const Format2Smem L2 = format2_smem_layout(h->sp);
        if (L2.total > 227 * 1024) { h->last_error = "reads / names too long for the format kernel's shared memory"; return DWGSIM_GPU_EUNSUPPORTED; }
        CUDA_TRY(h, cudaFuncSetAttribute(format2_kernel_of(h->sp), cudaFuncAttributeMaxDynamicSharedMemorySize, L2.total));
        return DWGSIM_GPU_OK;
    }
    h->sp.fmt_warps = kFmtWarps;
    while (h->sp.tile_pairs > 1 && format_smem_layout(h->sp).total > 113 * 1024) --h->sp.tile_pairs;
    while (h->sp.fmt_warps > 1 && format_smem_layout(h->sp).total > 227 * 1024) h->sp.fmt_warps >>= 1;   // long reads: fewer warps per CTA
    const FormatSmem L = format_smem_layout(h->sp);
    if (L.total > 227 * 1024) { h->last_error = "reads / names too long for the format kernel's shared memory"; return DWGSIM_GPU_EUNSUPPORTED; }
    CUDA_TRY(h, cudaFuncSetAttribute(format_kernel_of(h->sp), cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    return DWGSIM_GPU_OK;
}

void free_blob(dwgsim_gpu *h)
{
    if (h->blob && h->blob_owned) {                            // per-contig runs allocate the same few MB again and again
        if (h->blob_bytes > h->blob_spare_cap && h->blob_bytes <= (256ull << 20)) {
            cudaFree(h->blob_spare);
            h->blob_spare = h->blob; h->blob_spare_cap = h->blob_bytes;
        } else cudaFree(h->blob);
    }
    h->blob = nullptr; h->blob_bytes = 0; h->blob_pairs = 0; h->blob_owned = true;
    h->sp.regions = 0;
}

// lay the queued contigs out in one blob and upload it
int finalize_genome(dwgsim_gpu *h)
{
    if (h->blob) return DWGSIM_GPU_OK;
    if (h->queue.empty()) { h->last_error = "no contigs queued"; return DWGSIM_GPU_ESTATE; }
    { int rc0 = update_caps(h); if (rc0) return rc0; }
    const size_t nc = h->queue.size();
    uint64_t off = align_up(sizeof(BlobHeader), 256);
    BlobHeader hd{};
    hd.magic = kBlobMagic; hd.version = kBlobVersion; hd.n_contigs = (uint32_t)nc;
    hd.contigs_off = off;
    off = align_up(off + nc * sizeof(ContigDesc), 256);
    hd.names_off = off;
    std::vector<ContigDesc> cds(nc);
    std::string names;
    int64_t pair_base = 0, total_len = 0;
    for (size_t i = 0; i < nc; ++i) {
        HostContig &c = h->queue[i];
        ContigDesc &d = cds[i];
        memset(&d, 0, sizeof d);
        d.len = c.len; d.contig_i = c.contig_i; d.pair_base = pair_base; d.n_pairs = c.n_pairs;
        d.name_off = (uint32_t)names.size(); d.name_len = (uint32_t)c.name.size();
        names += c.name;
        pair_base += c.n_pairs; total_len += c.len;
    }
    off = align_up(off + names.size() + 1, 256);
    for (size_t i = 0; i < nc; ++i) {
        HostContig &c = h->queue[i];
        ContigDesc &d = cds[i];
        d.ref2_off = off; off = align_up(off + c.ref2.size() * 4, 256);
        d.nmask_off = off; off = align_up(off + c.nmask.size() * 4, 256);
        for (int hh = 0; hh < 2; ++hh) {
            d.n_ev[hh] = (uint32_t)c.ev[hh].size();
            d.ev_off[hh] = off; off = align_up(off + std::max<size_t>(c.ev[hh].size(), 1) * sizeof(Event), 256);
            d.blk_off[hh] = off; off = align_up(off + c.blk[hh].size() * 4, 256);
            d.pool_off[hh] = off; off = align_up(off + c.pool[hh].size(), 256);
        }
        d.n_reg = (uint32_t)c.regions.size(); d.sample_len = c.sample_len;
        if (c.sample_len > 0) hd.flags |= 1u;
        d.reg_off = off; off = align_up(off + c.regions.size() * sizeof(Region), 256);
    }
    hd.n_bytes = off; hd.total_pairs = pair_base; hd.total_len = total_len;
    if (h->blob_spare && h->blob_spare_cap >= off) { h->blob = h->blob_spare; h->blob_spare = nullptr; h->blob_spare_cap = 0; }
    else CUDA_TRY(h, cudaMalloc((void **)&h->blob, off));
    h->blob_owned = true; h->blob_bytes = off; h->blob_pairs = pair_base;
    h->sp.regions = (int32_t)(hd.flags & 1u);
    auto put = [&](uint64_t at, const void *src, size_t n) -> cudaError_t {
        h->h2d_bytes += (int64_t)n;
        return n ? cudaMemcpyAsync(h->blob + at, src, n, cudaMemcpyHostToDevice, h->s_compute) : cudaSuccess;
    };
    CUDA_TRY(h, put(0, &hd, sizeof hd));
    CUDA_TRY(h, put(hd.contigs_off, cds.data(), nc * sizeof(ContigDesc)));
    CUDA_TRY(h, put(hd.names_off, names.data(), names.size()));
    for (size_t i = 0; i < nc; ++i) {
        HostContig &c = h->queue[i];
        ContigDesc &d = cds[i];
        CUDA_TRY(h, put(d.ref2_off, c.ref2.data(), c.ref2.size() * 4));
        CUDA_TRY(h, put(d.nmask_off, c.nmask.data(), c.nmask.size() * 4));
        for (int hh = 0; hh < 2; ++hh) {
            CUDA_TRY(h, put(d.ev_off[hh], c.ev[hh].data(), c.ev[hh].size() * sizeof(Event)));
            CUDA_TRY(h, put(d.blk_off[hh], c.blk[hh].data(), c.blk[hh].size() * 4));
            CUDA_TRY(h, put(d.pool_off[hh], c.pool[hh].data(), c.pool[hh].size()));
        }
        CUDA_TRY(h, put(d.reg_off, c.regions.data(), c.regions.size() * sizeof(Region)));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->s_compute));
    h->queue.clear();
    h->queue.shrink_to_fit();
    return DWGSIM_GPU_OK;
}

void free_workspace(dwgsim_gpu *h)
{
    Workspace &w = h->ws;
    cudaFree(w.recs); cudaFree(w.seqs); cudaFree(w.serial); cudaFree(w.lens); cudaFree(w.names); cudaFree(w.name_len);
    cudaFree(w.blk_rand); cudaFree(w.blk_len); cudaFree(w.totals); cudaFree(w.status); cudaFree(w.jobs); cudaFree(w.flow_scratch);
    for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) cudaFree(w.out[s][k]);
    for (int k = 0; k < 3; ++k) { cudaFree(w.gz_slots[k]); cudaFree(w.gz_out[0][k]); cudaFree(w.gz_out[1][k]); }
    cudaFree(w.gz_sizes); cudaFree(w.gz_offs); cudaFree(w.gz_totals); cudaFree(w.gz_hist);
    if (w.h_totals) cudaFreeHost(w.h_totals);
    w = Workspace();
    for (int s = 0; s < h->pinned_slots; ++s) for (int k = 0; k < 3; ++k) if (h->pinned[s][k]) { cudaFreeHost(h->pinned[s][k]); h->pinned[s][k] = nullptr; }
    h->pinned_slots = 0;
}

int ensure_workspace(dwgsim_gpu *h, int64_t n, bool want_pinned)
{
    Workspace &w = h->ws;
    uint64_t cap_now[3];
    record_caps(h, cap_now);
    // the names buffer, the output streams and the pinned ring are sized from the longest contig name seen so far: a later,
    // longer name (add_contig / genome_import after a run) needs them again.  They are allocated with room for names 32
    // characters longer, so "chr9" -> "chr10" does not cost a reallocation (seconds, with the pinned ring)
    const bool grown = w.cap_pairs > 0 && (h->sp.name_cap > w.name_cap || cap_now[0] > w.rec_cap[0] || cap_now[1] > w.rec_cap[1] ||
                                           cap_now[2] > w.rec_cap[2]);
    if (w.cap_pairs < n || grown) {
        n = std::max<int64_t>(n, w.cap_pairs);
        free_workspace(h);
        const uint64_t kNameSlack = 32;
        w.name_cap = h->sp.name_cap + (int32_t)kNameSlack;
        record_caps(h, cap_now, kNameSlack);
        for (int k = 0; k < 3; ++k) w.rec_cap[k] = cap_now[k];
        const int64_t nblk = (n + kScanTile - 1) / kScanTile;
        CUDA_TRY(h, cudaMalloc((void **)&w.recs, (size_t)n * sizeof(PairRec)));
        CUDA_TRY(h, cudaMalloc((void **)&w.seqs, (size_t)n * 4 * (size_t)(h->sp.nw[0] + h->sp.nw[1] + 1)));
        CUDA_TRY(h, cudaMalloc((void **)&w.serial, (size_t)n * 8));
        CUDA_TRY(h, cudaMalloc((void **)&w.lens, (size_t)n * 12));
        CUDA_TRY(h, cudaMalloc((void **)&w.names, (size_t)n * 2 * (size_t)w.name_cap));
        CUDA_TRY(h, cudaMalloc((void **)&w.name_len, (size_t)n * 4));
        CUDA_TRY(h, cudaMalloc((void **)&w.blk_rand, (size_t)nblk * 8));
        CUDA_TRY(h, cudaMalloc((void **)&w.blk_len, (size_t)nblk * 24));
        CUDA_TRY(h, cudaMalloc((void **)&w.totals, 64));
        CUDA_TRY(h, cudaMalloc((void **)&w.status, 48));
        CUDA_TRY(h, cudaMemset(w.status, 0, 48));
        CUDA_TRY(h, cudaMalloc((void **)&w.jobs, (size_t)n * 2 * sizeof(uint2)));
        if (h->sp.data_type == 2) {                             // at most 16 CTAs of the simulate kernel per SM
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
            const size_t row_words = (size_t)(h->sp.nw[0] + h->sp.nw[1] + std::max(h->sp.nw[0], h->sp.nw[1]));
            CUDA_TRY(h, cudaMalloc((void **)&w.flow_scratch, (size_t)sms * 16 * kTpThreads * row_words * 4 + 256));
        }
        CUDA_TRY(h, cudaHostAlloc((void **)&w.h_totals, 128, cudaHostAllocMapped));
        CUDA_TRY(h, cudaHostGetDevicePointer((void **)&w.h_totals_dev, w.h_totals, 0));
        const uint64_t *cap = cap_now;
        for (int k = 0; k < 3; ++k) {
            w.out_cap[k] = align_up(cap[k] * (uint64_t)n + 256, 256);
            if (w.out_cap[k] >= (1ull << 32)) { h->last_error = "batch too large: a stream would exceed 4 GiB"; return DWGSIM_GPU_EINVAL; }
            for (int s = 0; s < 2; ++s) CUDA_TRY(h, cudaMalloc((void **)&w.out[s][k], w.out_cap[k]));
        }
        if (h->gz_mode) {
            w.gz_members_max = 1;
            for (int k = 0; k < 3; ++k) {
                w.gz_members[k] = (w.out_cap[k] + kGzMemberRaw - 1) / kGzMemberRaw + 1;
                w.gz_members_max = std::max(w.gz_members_max, w.gz_members[k]);
                w.gz_cap[k] = w.out_cap[k] + w.out_cap[k] / 4 + (1 << 20);
                CUDA_TRY(h, cudaMalloc((void **)&w.gz_slots[k], w.gz_members[k] * (uint64_t)kGzSlotStride));
                for (int s2 = 0; s2 < 2; ++s2) CUDA_TRY(h, cudaMalloc((void **)&w.gz_out[s2][k], w.gz_cap[k]));
            }
            CUDA_TRY(h, cudaMalloc((void **)&w.gz_sizes, 3 * w.gz_members_max * 8));
            CUDA_TRY(h, cudaMalloc((void **)&w.gz_offs, 3 * w.gz_members_max * 8));
            CUDA_TRY(h, cudaMalloc((void **)&w.gz_totals, 64));
            CUDA_TRY(h, cudaMalloc((void **)&w.gz_hist, 256 * 8));
        }
        w.cap_pairs = n;
    }
    if (want_pinned && h->pinned_slots == 0) {
        // Page-locking host memory costs about a second per GB, so the ring is sized for what the streams are expected to
        // need: the worst case of the FASTQ text, or 5/8 of it for gzip members (FASTQ under the literal-only code comes out
        // near 0.54; a batch that needs more makes run() enlarge the ring, see grow_ring)
        for (int k = 0; k < 3; ++k) h->pinned_cap[k] = h->gz_mode ? align_up(w.out_cap[k] / 8 * 5 + (64 << 10), 4096) : w.out_cap[k];
        for (int s = 0; s < h->ring; ++s)
            for (int k = 0; k < 3; ++k) CUDA_TRY(h, cudaMallocHost((void **)&h->pinned[s][k], h->pinned_cap[k]));
        h->pinned_slots = h->ring;
    }
    return DWGSIM_GPU_OK;
}

struct BatchResult {
    uint64_t bytes[3];
    int64_t n_random, n_failed;
    uint64_t status;
    float ms[3];
    int launches;
};

// phase 1 of a batch: simulate + count the random pairs (their number is needed before any name can be laid out,
// because rand_ii is a running count over all earlier pairs, src/dwgsim.c:1096)
int launch_simulate(dwgsim_gpu *h, int64_t first, int n, bool timed, int *launches)
{
    Workspace &w = h->ws;
    const SimParams &sp = h->sp;
    const int nblk = (n + kScanTile - 1) / kScanTile;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::min((n + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count * 8);
    const int cap0 = (sp.cap[0] + 15) & ~15, cap1 = (sp.cap[1] + 15) & ~15;
    const int flr = (sp.flow_order_len + 15) & ~15;
    const size_t smem_a = (size_t)kWarpsPerBlock * (cap0 + cap1 + flr) + flr;
    cudaStream_t st = h->s_compute;
    int extra_launches = 0;
    CUDA_TRY(h, cudaMemsetAsync(w.status, 0, 48, st));
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[0], st));
    if (h->ion_warp_kernel)
        simulate_pairs_kernel<<<grid, kThreads, smem_a, st>>>(sp, h->blob, first, h->gidx_origin, n, w.recs, w.seqs, w.status);
    else {
        // persistent grid: exactly the CTAs that are resident at once (a partial second wave would idle most SMs)
        const size_t smem_tp = tp_smem_bytes(sp);
        int occ_tp = 1;
        const tp_kernel_t tp = tp_kernel_of(sp);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tp, tp, kTpThreads, smem_tp);
        const int grid_tp = std::min((n + kTpThreads - 1) / kTpThreads, sm_count * std::min(std::max(occ_tp, 1), 16));
        JobLists J;
        J.retry = w.jobs; J.random = w.jobs + (size_t)w.cap_pairs; J.count = w.status + 2;
        // two passes (kernels.cuh "Job lists"): fresh pairs, then the retries and random pairs they left behind
        const int n_pass = sp.data_type == 2 ? 1 : 2;          // (Ion Torrent lanes finish their own retries)
        for (int pass = 0; pass < n_pass; ++pass) {
            tp<<<grid_tp, kTpThreads, smem_tp, st>>>(sp, h->blob, first, h->gidx_origin, n, pass, J, w.recs, w.seqs, w.status, w.flow_scratch);
        }
        extra_launches = n_pass - 1;
    }
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[1], st));
    layout_count_random_kernel<<<nblk, kThreads, 0, st>>>(w.recs, n, w.blk_rand);
    layout_scan_blocks_kernel<<<1, 1024, 0, st>>>(w.blk_rand, nblk, 1, w.totals);
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[5], st));
    CUDA_TRY(h, cudaGetLastError());
    *launches = 3 + extra_launches;
    return DWGSIM_GPU_OK;
}

// number of random pairs of the batch simulated last (synchronises the compute stream)
int read_random_count(dwgsim_gpu *h, int64_t *n_random)
{
    Workspace &w = h->ws;
    publish_words_kernel<<<1, 32, 0, h->s_compute>>>(w.h_totals_dev, w.totals, 1);
    CUDA_TRY(h, cudaStreamSynchronize(h->s_compute));
    *n_random = (int64_t)w.h_totals[0];
    return DWGSIM_GPU_OK;
}

// phase 2: record lengths -> offsets -> FASTQ text; results land in ws.h_totals after a sync
int launch_format(dwgsim_gpu *h, int64_t first, int n, int64_t rand_base, int slot, bool timed, int *launches,
                  const unsigned long long *rand_base_dev = nullptr)
{
    Workspace &w = h->ws;
    const SimParams &sp = h->sp;
    const int nblk = (n + kScanTile - 1) / kScanTile;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
    const int grid = std::min((n + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count * 8);
    const int cap0 = (sp.cap[0] + 15) & ~15, cap1 = (sp.cap[1] + 15) & ~15;
    const size_t smem_b = sp.fmt_v2 ? 0 : (size_t)format_smem_layout(sp).total;
    (void)cap0; (void)cap1;
    cudaStream_t st = h->s_compute;
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[4], st));
    layout_lengths_kernel<<<nblk, kThreads, 0, st>>>(sp, h->blob, w.recs, n, first, (unsigned long long)rand_base, rand_base_dev, w.blk_rand,
                                                     w.serial, w.lens, w.blk_len, w.names, w.name_len);
    layout_scan_blocks_kernel<<<3, 1024, 0, st>>>(w.blk_len, nblk, 3, w.totals + 1);
    layout_offsets_kernel<<<nblk, kThreads, 0, st>>>(n, w.blk_len, w.lens);
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[2], st));
    if (sp.fmt_v2) {
        const int fmt_warps = sp.fmt_warps > 0 ? sp.fmt_warps : kFmt2WarpsMax, fmt_threads = 32 * fmt_warps;
        const Format2Smem L2 = format2_smem_layout(sp);
        const size_t smem_2 = (size_t)L2.total;
        const int ctas = ((n + sp.tile_pairs - 1) / sp.tile_pairs + fmt_warps - 1) / fmt_warps;  // CTAs that have a mini-tile per warp
        const int grid_f = std::min(ctas, sm_count);
        format2_kernel_of(sp)<<<grid_f, fmt_threads, smem_2, st>>>(sp, L2, first, h->gidx_origin, n, w.recs, w.seqs, w.lens, w.totals + 1,
                                                                  w.names, w.name_len, w.out[slot][0], w.out[slot][1], w.out[slot][2]);
    } else {
        const int fmt_warps = sp.fmt_warps > 0 ? sp.fmt_warps : kFmtWarps, fmt_threads = 32 * fmt_warps;
        const int ntiles = ((n + sp.tile_pairs - 1) / sp.tile_pairs + fmt_warps - 1) / fmt_warps;    // CTAs that have a mini-tile per warp
        int occ_f = 1;
        const format_kernel_t fmt = format_kernel_of(sp);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, fmt, fmt_threads, smem_b);
        const int grid_f = std::min(ntiles, sm_count * std::max(occ_f, 1));
        fmt<<<grid_f, fmt_threads, smem_b, st>>>(sp, h->blob, first, h->gidx_origin, n, w.recs, w.seqs, w.serial, w.lens,
                                                 w.totals + 1, w.names, w.name_len, w.out[slot][0], w.out[slot][1], w.out[slot][2]);
    }
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev_t[3], st));
    CUDA_TRY(h, cudaGetLastError());
    // (a copy would queue behind the batch-sized device-to-host transfers of the previous batches on the copy engine)
    publish_batch_kernel<<<1, 32, 0, st>>>(w.h_totals_dev, w.totals, w.status, h->queue_active ? h->queue_dev : nullptr, h->queue_advance ? 1 : 0);
    *launches = 4;
    return DWGSIM_GPU_OK;
}

int launch_batch(dwgsim_gpu *h, int64_t first, int n, int64_t rand_base, int slot, bool timed, int *launches)
{
    int l1 = 0, l2 = 0, rc;
    if ((rc = launch_simulate(h, first, n, timed, &l1))) return rc;
    if ((rc = launch_format(h, first, n, rand_base, slot, timed, &l2))) return rc;
    *launches = l1 + l2;
    return DWGSIM_GPU_OK;
}

int collect_batch(dwgsim_gpu *h, bool timed, BatchResult *r)
{
    Workspace &w = h->ws;
    CUDA_TRY(h, cudaStreamSynchronize(h->s_compute));
    r->n_random = (int64_t)w.h_totals[0];
    for (int k = 0; k < 3; ++k) r->bytes[k] = w.h_totals[1 + k];
    r->status = w.h_totals[8];
    r->n_failed = (int64_t)w.h_totals[9];
    r->ms[0] = r->ms[1] = r->ms[2] = 0;
    if (timed) {
        float a = 0, b = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&r->ms[0], h->ev_t[0], h->ev_t[1]));
        CUDA_TRY(h, cudaEventElapsedTime(&a, h->ev_t[4], h->ev_t[2]));        // lengths + scan + offsets
        CUDA_TRY(h, cudaEventElapsedTime(&b, h->ev_t[1], h->ev_t[5]));        // count + scan
        r->ms[1] = a + b;
        CUDA_TRY(h, cudaEventElapsedTime(&r->ms[2], h->ev_t[2], h->ev_t[3]));
    }
    if (r->status & 1ull) {
        h->last_error = "failed to generate a read after 10001 trials";
        return DWGSIM_GPU_ETRIALS;
    }
    if (r->status & 2ull) {
        h->last_error = "Ion Torrent read grew past 2*len+64 bases";
        return DWGSIM_GPU_EOVERFLOW;
    }
    return DWGSIM_GPU_OK;
}


// fit the per-stream Huffman codes to the first batch and upload the tables
int gz_calibrate(dwgsim_gpu *h, int dslot, const uint64_t bytes[3])
{
    Workspace &w = h->ws;
    static const Crc32Tables ct;
    if (!h->gz_crc) {
        std::vector<uint32_t> v(1024 + 32);
        for (int s2 = 0; s2 < 4; ++s2) memcpy(&v[(size_t)s2 * 256], ct.t[s2], 1024);
        memcpy(&v[1024], ct.x2n, 128);
        int rc = upload(h, &h->gz_crc, v.data(), v.size());
        if (rc) return rc;
    }
    // the devices of a group all use the code the leader fitted to batch 0, so a group writes the bytes of a single device
    dwgsim_gpu *const lead = h->group ? h->group->leader : h;
    if (lead != h) {
        std::unique_lock<std::mutex> lk(h->group->mu);
        h->group->cv.wait(lk, [&]() { return lead->gz_hist_valid || h->group->failed; });
        if (!lead->gz_hist_valid) { h->last_error = "another device of the group failed"; return DWGSIM_GPU_ESTATE; }
    }
    for (int k = 0; k < 3; ++k) {
        uint64_t *hist = lead->gz_hist_host[k];
        if (lead == h) {
            memset(hist, 0, 256 * 8);
            if (bytes[k]) {
                CUDA_TRY(h, cudaMemsetAsync(w.gz_hist, 0, 256 * 8, h->s_compute));
                gz_histogram_kernel<<<592, 256, 0, h->s_compute>>>((const uint8_t *)w.out[dslot][k], bytes[k], w.gz_hist);
                CUDA_TRY(h, cudaMemcpyAsync(hist, w.gz_hist, 256 * 8, cudaMemcpyDeviceToHost, h->s_compute));
                CUDA_TRY(h, cudaStreamSynchronize(h->s_compute));
            }
        }
        const GzTables t = gz_build_tables(hist);
        cudaFree(h->gz_code[k]); cudaFree(h->gz_prefix[k]);
        h->gz_code[k] = nullptr; h->gz_prefix[k] = nullptr;
        int rc = upload(h, &h->gz_code[k], t.code, (size_t)257);
        if (rc) return rc;
        std::vector<uint8_t> pre = t.prefix;
        pre.resize((pre.size() + 11) & ~3ull, 0);               // whole words plus one
        if ((rc = upload(h, &h->gz_prefix[k], pre.data(), pre.size()))) return rc;
        h->gz_prefix_bits[k] = t.prefix_bits;
    }
    h->gz_ready = true;
    if (lead == h) {
        if (h->group) { std::lock_guard<std::mutex> g(h->group->mu); h->gz_hist_valid = true; h->group->cv.notify_all(); }
        else h->gz_hist_valid = true;
    }
    return DWGSIM_GPU_OK;
}

// raw streams of device slot `dslot` -> gzip members -> contiguous streams in ws.gz_out[dslot]; sizes in out_bytes
int gz_batch(dwgsim_gpu *h, int dslot, const uint64_t bytes[3], uint64_t out_bytes[3], int *launches)
{
    Workspace &w = h->ws;
    int rc;
    if (!h->gz_ready && (rc = gz_calibrate(h, dslot, bytes))) return rc;
    cudaStream_t st = h->s_compute;
    CUDA_TRY(h, cudaEventRecord(h->ev_t[6], st));
    CUDA_TRY(h, cudaMemsetAsync(w.gz_totals, 0, 64, st));
    for (int k = 0; k < 3; ++k) {
        if (!bytes[k]) continue;
        const int nm = (int)((bytes[k] + kGzMemberRaw - 1) / kGzMemberRaw);
        unsigned long long *sizes = w.gz_sizes + (size_t)k * w.gz_members_max, *offs = w.gz_offs + (size_t)k * w.gz_members_max;
        GzDeviceTables T{h->gz_code[k], h->gz_prefix[k], h->gz_prefix_bits[k]};
        gz_compress_kernel<<<nm, kGzThreads, 0, st>>>((const uint8_t *)w.out[dslot][k], bytes[k], T, h->gz_crc, w.gz_slots[k], sizes);
        CUDA_TRY(h, cudaMemcpyAsync(offs, sizes, (size_t)nm * 8, cudaMemcpyDeviceToDevice, st));
        layout_scan_blocks_kernel<<<1, 1024, 0, st>>>(offs, nm, 1, w.gz_totals + k);
        gz_compact_kernel<<<nm, 256, 0, st>>>(w.gz_slots[k], sizes, offs, (uint8_t *)w.gz_out[dslot][k], w.gz_cap[k]);
        *launches += 3;
    }
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev_t[7], st));
    publish_words_kernel<<<1, 32, 0, st>>>(w.h_totals_dev + 12, w.gz_totals, 3);
    CUDA_TRY(h, cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, h->ev_t[6], h->ev_t[7]); h->ms_gz += ms; }
    for (int k = 0; k < 3; ++k) {
        out_bytes[k] = bytes[k] ? w.h_totals[12 + k] : 0;
        if (out_bytes[k] > w.gz_cap[k]) { h->last_error = "compressed stream larger than its buffer"; return DWGSIM_GPU_EOVERFLOW; }
    }
    return DWGSIM_GPU_OK;
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------------------------
extern "C" {

int dwgsim_gpu_abi_version(void) { return DWGSIM_GPU_ABI_VERSION; }

const char *dwgsim_gpu_strerror(int code)
{
    switch (code) {
        case DWGSIM_GPU_OK: return "ok";
        case DWGSIM_GPU_EINVAL: return "invalid argument";
        case DWGSIM_GPU_ENODEV: return "no CUDA device (the read-pair path has no CPU fallback)";
        case DWGSIM_GPU_ECUDA: return "CUDA error";
        case DWGSIM_GPU_ENOMEM: return "out of memory";
        case DWGSIM_GPU_ETRIALS: return "failed to generate a read after 10001 trials";
        case DWGSIM_GPU_ESINK: return "output sink failed";
        case DWGSIM_GPU_EUNSUPPORTED: return "option not supported on the device path";
        case DWGSIM_GPU_EOVERFLOW: return "Ion Torrent read grew past the device bound";
        case DWGSIM_GPU_ESTATE: return "call order violated";
        default: return "unknown error";
    }
}

const char *dwgsim_gpu_last_error(const dwgsim_gpu_t *h) { return h ? h->last_error.c_str() : ""; }

int dwgsim_gpu_create(dwgsim_gpu_t **out, const dwgsim_gpu_params_t *p, int device)
{
    if (!out || !p) return DWGSIM_GPU_EINVAL;
    *out = nullptr;
    // the ranges dwgsim_opt_parse enforces (src/dwgsim_opt.c:307-371)
    if (p->length[0] < 1 || p->length[1] < 0 || p->dist < 0 || p->std_dev < 0) return DWGSIM_GPU_EINVAL;
    if (p->data_type < 0 || p->data_type > 2 || p->strandedness < 0 || p->strandedness > 2) return DWGSIM_GPU_EINVAL;
    if (p->read_one_strand < 0 || p->read_one_strand > 2 || p->max_n < 0) return DWGSIM_GPU_EINVAL;
    if (p->rand_read < 0 || p->rand_read > 1 || p->mut_freq < 0 || p->mut_freq > 1) return DWGSIM_GPU_EINVAL;
    if (p->reads_output_type < 0 || p->reads_output_type > 2 || p->quality_std < 0) return DWGSIM_GPU_EINVAL;
    if (p->data_type == 2 && (!p->flow_order || p->flow_order_len <= 0)) return DWGSIM_GPU_EINVAL;
    if (p->length[0] > 30000 || p->length[1] > 30000) return DWGSIM_GPU_EUNSUPPORTED;   // 16-bit lengths in PairRec
    if (p->std_dev > 1.0e6) return DWGSIM_GPU_EUNSUPPORTED;                              // insert-size table size
    if (p->data_type == 2) {
        // every base must have a flow (the reference loops forever otherwise, src/dwgsim.c:283-286)
        int seen = 0;
        if (p->flow_order_len > 1024) return DWGSIM_GPU_EUNSUPPORTED;
        for (int i = 0; i < p->flow_order_len; ++i) if (p->flow_order[i] >= 0 && p->flow_order[i] < 4) seen |= 1 << p->flow_order[i];
        if (seen != 15) return DWGSIM_GPU_EUNSUPPORTED;
        for (int e = 0; e < 2; ++e)                                                      // src/dwgsim_opt.c:338-343
            if (p->length[e] > 0 && p->e_by[e] != 0.0) return DWGSIM_GPU_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return DWGSIM_GPU_ENODEV;
    dwgsim_gpu *h = new dwgsim_gpu();
    h->p = *p;
    for (int e = 0; e < 2; ++e) if (p->length[e] == 0) h->p.e_by[e] = 0.0;   // the reference leaves 0/0 here (src/dwgsim_opt.c:460)
    h->device = device;
    if (p->read_prefix) { h->prefix_s = std::string(p->read_prefix) + "_"; }
    if (p->flow_order) h->flow_order.assign(p->flow_order, p->flow_order + p->flow_order_len);
    h->p.read_prefix = nullptr; h->p.flow_order = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking) != cudaSuccess) { delete h; return DWGSIM_GPU_ECUDA; }
    for (auto &e : h->ev_t) cudaEventCreate(&e);
    derive_tables(h);
    int rc = upload_tables(h);
    if (rc) { dwgsim_gpu_destroy(h); return rc; }
    {   // shared memory per CTA: read codes (+ flow mask) per warp; names + qualities per warp
        const SimParams &sp = h->sp;
        const int cap0 = (sp.cap[0] + 15) & ~15, cap1 = (sp.cap[1] + 15) & ~15, flr = (sp.flow_order_len + 15) & ~15;
        const size_t smem_a = (size_t)kWarpsPerBlock * (cap0 + cap1 + flr) + flr;
        // the reference window of the thread-per-pair kernel costs shared memory: keep it when at least four CTAs (16 warps)
        // still fit an SM or when it does not cost a CTA (long Ion Torrent rows leave room for two or three CTAs only)
        {
            auto ctas = [&](size_t bytes) { return (int)std::min<size_t>(kTpMinBlocks, (227 * 1024) / (bytes + 1024)); };
            h->sp.tp_tables = 3;
            h->sp.win_slots = 0;
            const int without = ctas(tp_smem_bytes(h->sp));
            h->sp.win_slots = std::max(window_slots(sp.len[0]), window_slots(sp.len[1]));
            const int with = ctas(tp_smem_bytes(h->sp));
            if (const char *e = getenv("DWGSIM_WINDOW")) { if (atoi(e) == 0) h->sp.win_slots = 0; }
            else if (with < 4 && with < without) h->sp.win_slots = 0;
            // the sampling tables leave shared memory (guides first, then the error tables, then the insert-size CDF) when
            // that lets one more CTA fit: a pair reads a handful of their entries, an SM gains four warps
            if (sp.data_type != 2) {
                h->sp.tp_tables = 0;
                const int without_tables = ctas(tp_smem_bytes(h->sp));
                h->sp.tp_tables = 3;
                if (ctas(tp_smem_bytes(h->sp)) < without_tables) h->sp.tp_tables = 0;
                if (const char *e = getenv("DWGSIM_TP_TABLES")) h->sp.tp_tables = atoi(e) >= 3 ? 3 : 0;
            }
        }
        const size_t smem_tp = tp_smem_bytes(sp);
        // one staging row per thread in shared memory bounds the combined read length (about 3,400 symbols); Ion Torrent
        // rows hold 2*len+64 symbols and longer ones fall back to the warp-per-pair kernel
        const bool tp_fits = smem_tp <= 220 * 1024;
        const char *force = getenv("DWGSIM_ION_KERNEL");
        h->ion_warp_kernel = sp.data_type == 2 && (!tp_fits || (force && strcmp(force, "warp") == 0));
        if (sp.data_type != 2 && !tp_fits) { dwgsim_gpu_destroy(h); return DWGSIM_GPU_EUNSUPPORTED; }
        if (!h->ion_warp_kernel &&
            cudaFuncSetAttribute(tp_kernel_of(sp), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tp) != cudaSuccess) {
            dwgsim_gpu_destroy(h); return DWGSIM_GPU_ECUDA;
        }
        if (smem_a > 227 * 1024) { dwgsim_gpu_destroy(h); return DWGSIM_GPU_EUNSUPPORTED; }
        if (cudaFuncSetAttribute(simulate_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a) != cudaSuccess) {
            dwgsim_gpu_destroy(h); return DWGSIM_GPU_ECUDA;
        }
    }
    *out = h;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_create_group(dwgsim_gpu_t **out, const dwgsim_gpu_params_t *p, const int32_t *devices, int32_t n_devices)
{
    if (!out || !p || !devices || n_devices < 1 || n_devices > 64) return DWGSIM_GPU_EINVAL;
    *out = nullptr;
    // (a device may appear more than once: its ranks then share it, each with its own streams and buffers)
    dwgsim_gpu *lead = nullptr;
    int rc = dwgsim_gpu_create(&lead, p, devices[0]);
    if (rc) return rc;
    for (int i = 1; i < n_devices; ++i) {
        dwgsim_gpu *q = nullptr;
        if ((rc = dwgsim_gpu_create(&q, p, devices[i]))) { dwgsim_gpu_destroy(lead); return rc; }
        lead->peers.push_back(q);
        // direct device-to-device copies of the genome blob (NVLink between the GPUs of one box); without peer access the
        // runtime stages the copy through the host
        int can = 0;
        if (devices[i] != devices[0] && cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
            cudaSetDevice(devices[i]);
            if (cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();   // (already enabled: fine)
        }
    }
    cudaSetDevice(devices[0]);
    *out = lead;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_group_size(const dwgsim_gpu_t *h) { return h ? 1 + (int)h->peers.size() : 0; }

void dwgsim_gpu_destroy(dwgsim_gpu_t *h)
{
    if (!h) return;
    for (dwgsim_gpu *q : h->peers) dwgsim_gpu_destroy(q);
    h->peers.clear();
    cudaSetDevice(h->device);
    free_workspace(h);
    free_blob(h);
    cudaFree(h->blob_spare);
    cudaFree(h->queue_dev);
    cudaFree(h->dt.isize_cdf); cudaFree(h->dt.qdelta_cdf); cudaFree(h->dt.qguide); cudaFree(h->dt.qtab); cudaFree(h->dt.isize_guide); cudaFree(h->dt.gap_guide[0]); cudaFree(h->dt.gap_guide[1]);
    for (int e = 0; e < 2; ++e) { cudaFree(h->dt.err_gap[e]); cudaFree(h->dt.err_acc[e]); cudaFree(h->dt.qbase[e]); }
    cudaFree(h->dt.flow_order); cudaFree(h->dt.prefix); cudaFree(h->dt.flow_gap[0]); cudaFree(h->dt.flow_gap[1]);
    for (int k = 0; k < 3; ++k) { cudaFree(h->gz_code[k]); cudaFree(h->gz_prefix[k]); }
    cudaFree(h->gz_crc);
    for (auto &e : h->ev_t) if (e) cudaEventDestroy(e);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    delete h;
}

struct dwgsim_gpu_packed {
    HostContig c;
    double ms_pack = 0;
};

// Packing is host-only work on the caller's arrays: it may run on another thread while dwgsim_gpu_run is in flight on
// the same handle (it reads nothing of the handle but its packer-thread count and read-prefix length).
int dwgsim_gpu_pack_contig(const dwgsim_gpu_t *h, int32_t contig_i, const char *name, const uint8_t *seq_ascii, int32_t len,
                           const uint64_t *hap1, const uint64_t *hap2, uint8_t *const *ins1, int32_t ins1_n,
                           uint8_t *const *ins2, int32_t ins2_n, int64_t n_pairs, dwgsim_gpu_packed_t **out)
{
    if (!h || !out || !name || !seq_ascii || !hap1 || !hap2 || len <= 0 || n_pairs < 0) return DWGSIM_GPU_EINVAL;
    *out = nullptr;
    const size_t nl = strlen(name);
    if (nl + h->prefix_s.size() + 128 > 1024) return DWGSIM_GPU_EUNSUPPORTED;       // read name too long for the device formatter
    const double t0 = now_ms();
    dwgsim_gpu_packed *p = new dwgsim_gpu_packed();
    HostContig &c = p->c;
    c.name = name; c.contig_i = contig_i; c.len = len; c.n_pairs = n_pairs;
    const uint64_t *hap[2] = {hap1, hap2};
    uint8_t *const *ins[2] = {ins1, ins2};
    const int32_t ins_n[2] = {ins1_n, ins2_n};
    std::string err;
    const int rc = pack_contig(h->host_threads, err, c, seq_ascii, hap, ins, ins_n);
    if (rc) { delete p; return rc; }
    p->ms_pack = now_ms() - t0;
    *out = p;
    return DWGSIM_GPU_OK;
}

void dwgsim_gpu_packed_free(dwgsim_gpu_packed_t *p) { delete p; }

int dwgsim_gpu_add_packed(dwgsim_gpu_t *h, dwgsim_gpu_packed_t *p)
{
    if (!h || !p) return DWGSIM_GPU_EINVAL;
    if (h->blob) { h->last_error = "add_contig after the genome was finalized: call run() first"; return DWGSIM_GPU_ESTATE; }
    h->max_name_len = std::max(h->max_name_len, (int)p->c.name.size());
    h->ms_pack += p->ms_pack;
    h->queue.emplace_back(std::move(p->c));
    delete p;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_add_contig(dwgsim_gpu_t *h, int32_t contig_i, const char *name, const uint8_t *seq_ascii, int32_t len,
                          const uint64_t *hap1, const uint64_t *hap2, uint8_t *const *ins1, int32_t ins1_n,
                          uint8_t *const *ins2, int32_t ins2_n, int64_t n_pairs)
{
    if (!h || !name || !seq_ascii || !hap1 || !hap2 || len <= 0 || n_pairs < 0) return DWGSIM_GPU_EINVAL;
    if (h->blob) { h->last_error = "add_contig after the genome was finalized: call run() first"; return DWGSIM_GPU_ESTATE; }
    dwgsim_gpu_packed_t *p = nullptr;
    const int rc = dwgsim_gpu_pack_contig(h, contig_i, name, seq_ascii, len, hap1, hap2, ins1, ins1_n, ins2, ins2_n, n_pairs, &p);
    if (rc == DWGSIM_GPU_EUNSUPPORTED && !p) h->last_error = "read name too long for the device formatter, or an insertion longer than 2^27-1 bases";
    else if (rc) h->last_error = "packing the contig failed (long insertion index out of range?)";
    if (rc) return rc;
    return dwgsim_gpu_add_packed(h, p);
}

int dwgsim_gpu_warm(dwgsim_gpu_t *h)
{
    if (!h) return DWGSIM_GPU_EINVAL;
    // the batch workspace and the pinned ring (the slow part: page-locking a GB takes about a second) of every device,
    // sized for the names seen so far plus some room; run() allocates the same on first use otherwise
    std::vector<dwgsim_gpu *> all{h};
    all.insert(all.end(), h->peers.begin(), h->peers.end());
    std::vector<int> rcs(all.size(), DWGSIM_GPU_OK);
    std::vector<std::thread> th;
    auto one = [&](size_t i) {
        dwgsim_gpu *q = all[i];
        cudaSetDevice(q->device);
        int rc = update_caps(q);
        if (rc == DWGSIM_GPU_OK) rc = ensure_workspace(q, q->batch_pairs, true);
        rcs[i] = rc;
    };
    for (size_t i = 1; i < all.size(); ++i) th.emplace_back(one, i);
    one(0);
    for (auto &t : th) t.join();
    cudaSetDevice(h->device);
    for (size_t i = 0; i < all.size(); ++i) if (rcs[i]) { h->last_error = all[i]->last_error; return rcs[i]; }
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_host_threads(dwgsim_gpu_t *h, int32_t n)
{
    if (!h || n < 0 || n > 1024) return DWGSIM_GPU_EINVAL;
    h->host_threads = n;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_regions(dwgsim_gpu_t *h, const uint32_t *start, const uint32_t *end, int32_t n, int32_t sample_len)
{
    if (!h || n < 0 || (n > 0 && (!start || !end)) || sample_len <= 0) return DWGSIM_GPU_EINVAL;
    if (h->blob || h->queue.empty()) { h->last_error = "set_regions applies to the contig just queued with add_contig"; return DWGSIM_GPU_ESTATE; }
    if (h->p.amplicons) { h->last_error = "regions cannot be combined with amplicon mode"; return DWGSIM_GPU_EINVAL; }
    HostContig &c = h->queue.back();
    std::vector<Region> r((size_t)n);
    uint32_t cum = 0;
    for (int32_t i = 0; i < n; ++i) {
        // the invariants regions_bed_init leaves behind (src/regions_bed.c:66-97): inside the contig, sorted, merged
        if (end[i] < start[i] || end[i] > (uint32_t)c.len || (i > 0 && start[i] <= end[i - 1])) {
            h->last_error = "regions must lie inside the contig, sorted by start and not touch or overlap";
            return DWGSIM_GPU_EINVAL;
        }
        r[(size_t)i] = Region{start[i], end[i], cum, 0u};
        cum += end[i] - start[i];
    }
    c.regions.swap(r);
    c.sample_len = sample_len;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_batch(dwgsim_gpu_t *h, int64_t pairs_per_batch, int32_t ring_slots)
{
    if (!h || pairs_per_batch < 1 || pairs_per_batch > (1 << 24) || ring_slots < 2 || ring_slots > 8) return DWGSIM_GPU_EINVAL;
    for (dwgsim_gpu *q : h->peers) dwgsim_gpu_set_batch(q, pairs_per_batch, ring_slots);
    cudaSetDevice(h->device);
    free_workspace(h);
    h->batch_pairs = pairs_per_batch; h->ring = ring_slots;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_compression(dwgsim_gpu_t *h, int32_t mode)
{
    if (!h || mode < 0 || mode > 1) return DWGSIM_GPU_EINVAL;
    for (dwgsim_gpu *q : h->peers) dwgsim_gpu_set_compression(q, mode);
    cudaSetDevice(h->device);
    if (mode != h->gz_mode) free_workspace(h);
    h->gz_mode = mode;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_shard(dwgsim_gpu_t *h, int32_t rank, int32_t world)
{
    if (!h || world < 1 || rank < 0 || rank >= world) return DWGSIM_GPU_EINVAL;
    if (!h->peers.empty()) { h->last_error = "a device group shards its batches itself"; return DWGSIM_GPU_ESTATE; }
    h->shard_rank = rank; h->shard_world = world;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_set_origin(dwgsim_gpu_t *h, int64_t first_pair_index, int64_t first_rand_serial)
{
    if (!h || first_pair_index < 0 || first_rand_serial < 0) return DWGSIM_GPU_EINVAL;
    h->gidx_origin = first_pair_index; h->rand_serial = first_rand_serial;
    for (dwgsim_gpu *q : h->peers) { q->gidx_origin = first_pair_index; q->rand_serial = first_rand_serial; }
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_genome_finalize(dwgsim_gpu_t *h)
{
    if (!h) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    return finalize_genome(h);
}

int dwgsim_gpu_genome_blob(const dwgsim_gpu_t *h, uint64_t *device_ptr, uint64_t *n_bytes)
{
    if (!h || !h->blob) return DWGSIM_GPU_ESTATE;
    if (device_ptr) *device_ptr = (uint64_t)(uintptr_t)h->blob;
    if (n_bytes) *n_bytes = h->blob_bytes;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_genome_import(dwgsim_gpu_t *h, uint64_t device_ptr, uint64_t n_bytes, int32_t take_ownership)
{
    if (!h || !device_ptr || n_bytes < sizeof(BlobHeader)) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    BlobHeader hd;
    CUDA_TRY(h, cudaMemcpy(&hd, (const void *)(uintptr_t)device_ptr, sizeof hd, cudaMemcpyDeviceToHost));
    if (hd.magic != kBlobMagic || hd.version != kBlobVersion || hd.n_bytes != n_bytes) { h->last_error = "not a genome blob"; return DWGSIM_GPU_EINVAL; }
    std::vector<ContigDesc> cds(hd.n_contigs);
    CUDA_TRY(h, cudaMemcpy(cds.data(), (const uint8_t *)(uintptr_t)device_ptr + hd.contigs_off, hd.n_contigs * sizeof(ContigDesc),
                           cudaMemcpyDeviceToHost));
    free_blob(h);
    h->queue.clear();
    h->blob = (uint8_t *)(uintptr_t)device_ptr; h->blob_bytes = n_bytes; h->blob_owned = take_ownership != 0;
    h->blob_pairs = hd.total_pairs;
    h->sp.regions = (int32_t)(hd.flags & 1u);
    for (auto &d : cds) h->max_name_len = std::max(h->max_name_len, (int)d.name_len);
    return update_caps(h);
}

int64_t dwgsim_gpu_genome_pairs(const dwgsim_gpu_t *h)
{
    if (!h) return -1;
    if (h->blob) return h->blob_pairs;
    int64_t n = 0;
    for (auto &c : h->queue) n += c.n_pairs;
    return n;
}

void *dwgsim_gpu_cuda_stream(const dwgsim_gpu_t *h) { return h ? (void *)h->s_compute : nullptr; }

int dwgsim_gpu_tables(const dwgsim_gpu_t *h, dwgsim_gpu_tables_t *t)
{
    if (!h || !t) return DWGSIM_GPU_EINVAL;
    t->thr_genomic = h->thr_genomic; t->thr_hap0 = h->thr_hap0;
    t->isize_lo = h->isize_lo; t->isize_n = (int32_t)h->isize_cdf.size(); t->isize_cdf = h->isize_cdf.data();
    t->qdelta_lo = h->qdelta_lo; t->qdelta_n = (int32_t)h->qdelta_cdf.size(); t->qdelta_cdf = h->qdelta_cdf.data();
    for (int e = 0; e < 2; ++e) {
        t->n_cycles[e] = (int32_t)h->err_gap[e].size() - 1;
        t->err_gap[e] = h->err_gap[e].data(); t->err_acc[e] = h->err_acc[e].data(); t->qbase[e] = h->qbase[e].data();
        t->flow_thr[e] = h->flow_thr[e];
        t->flow_gap[e] = h->flow_gap[e].data(); t->flow_gap_n[e] = (int32_t)h->flow_gap[e].size();
    }
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_resident_begin(dwgsim_gpu_t *h, int64_t first, int64_t n, int64_t *n_random)
{
    if (!h || n < 1 || first < 0) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    int rc = finalize_genome(h);
    if (rc) return rc;
    if (first + n > h->blob_pairs || n > (1 << 24)) return DWGSIM_GPU_EINVAL;
    if ((rc = ensure_workspace(h, std::max<int64_t>(n, h->ws.cap_pairs), false))) return rc;
    int launches = 0;
    if ((rc = launch_simulate(h, first, (int)n, true, &launches))) return rc;
    h->pending_first = first; h->pending_n = (int)n; h->pending_launches = launches;
    if (n_random) return read_random_count(h, n_random);
    return DWGSIM_GPU_OK;
}

namespace {
int resident_finish(dwgsim_gpu_t *h, int64_t rand_serial_base, const unsigned long long *rand_base_dev, dwgsim_gpu_batch_t *out)
{
    if (!h || !out || h->pending_first < 0) return DWGSIM_GPU_ESTATE;
    cudaSetDevice(h->device);
    int launches = 0, rc;
    const int64_t first = h->pending_first;
    const int n = h->pending_n;
    h->pending_first = -1;
    if ((rc = launch_format(h, first, n, rand_serial_base, 0, true, &launches, rand_base_dev))) return rc;
    BatchResult r;
    rc = collect_batch(h, true, &r);
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 3; ++k) { out->dev_ptr[k] = (uint64_t)(uintptr_t)h->ws.out[0][k]; out->n_bytes[k] = r.bytes[k]; h->last_bytes[k] = r.bytes[k]; }
    h->last_slot = 0;
    out->n_pairs = n; out->n_random = r.n_random; out->n_failed_attempts = r.n_failed;
    out->ms_simulate = r.ms[0]; out->ms_layout = r.ms[1]; out->ms_format = r.ms[2];
    out->n_launches = launches + h->pending_launches;
    return rc;
}
}  // namespace

int dwgsim_gpu_resident_finish(dwgsim_gpu_t *h, int64_t rand_serial_base, dwgsim_gpu_batch_t *out)
{
    return resident_finish(h, rand_serial_base, nullptr, out);
}

// ---- batches queued back to back: no host round trip between them -----------------------------------------------------
namespace {
int queue_ready(dwgsim_gpu_t *h)
{
    if (h->queue_dev) return DWGSIM_GPU_OK;
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMalloc((void **)&h->queue_dev, 32));
    CUDA_TRY(h, cudaMemset(h->queue_dev, 0, 32));
    return DWGSIM_GPU_OK;
}
// layout + format of the batch begun last, queued behind its simulate passes; nothing is waited for
int queue_finish(dwgsim_gpu_t *h, const unsigned long long *rand_base_dev, bool advance)
{
    if (h->pending_first < 0) return DWGSIM_GPU_ESTATE;
    int launches = 0;
    const int64_t first = h->pending_first;
    const int n = h->pending_n;
    h->pending_first = -1;
    h->queue_active = true; h->queue_advance = advance;
    const int rc = launch_format(h, first, n, 0, 0, true, &launches, rand_base_dev);
    h->queue_active = h->queue_advance = false;
    if (rc) return rc;
    h->queued_launches += launches + h->pending_launches;
    h->queued_n = n;
    return DWGSIM_GPU_OK;
}
}  // namespace

int dwgsim_gpu_resident_set_running(dwgsim_gpu_t *h, int64_t rand_serial)
{
    if (!h || rand_serial < 0) return DWGSIM_GPU_EINVAL;
    int rc = queue_ready(h);
    if (rc) return rc;
    const unsigned long long v = (unsigned long long)rand_serial;
    CUDA_TRY(h, cudaMemcpyAsync(h->queue_dev, &v, 8, cudaMemcpyHostToDevice, h->s_compute));
    CUDA_TRY(h, cudaStreamSynchronize(h->s_compute));
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_resident_enqueue(dwgsim_gpu_t *h, int64_t first, int64_t n)
{
    if (!h) return DWGSIM_GPU_EINVAL;
    int rc = queue_ready(h);
    if (rc || (rc = dwgsim_gpu_resident_begin(h, first, n, nullptr))) return rc;
    return queue_finish(h, h->queue_dev, true);
}

int dwgsim_gpu_resident_finish_async(dwgsim_gpu_t *h, uint64_t rand_serial_base_device_ptr)
{
    if (!h || !rand_serial_base_device_ptr) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    int rc = queue_ready(h);
    if (rc) return rc;
    return queue_finish(h, reinterpret_cast<const unsigned long long *>((uintptr_t)rand_serial_base_device_ptr), false);
}

int dwgsim_gpu_resident_finish_gathered(dwgsim_gpu_t *h, uint64_t counts_device_ptr, int32_t world, int32_t rank)
{
    if (!h || !counts_device_ptr || world < 1 || rank < 0 || rank >= world) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    int rc = queue_ready(h);
    if (rc) return rc;
    gathered_base_kernel<<<1, 32, 0, h->s_compute>>>(reinterpret_cast<const unsigned long long *>((uintptr_t)counts_device_ptr), world, rank, h->queue_dev);
    return queue_finish(h, h->queue_dev + 2, false);
}

int dwgsim_gpu_resident_wait(dwgsim_gpu_t *h, dwgsim_gpu_batch_t *out)
{
    if (!h || !out || !h->queue_dev || h->queued_n <= 0) return DWGSIM_GPU_ESTATE;
    cudaSetDevice(h->device);
    BatchResult r;
    int rc = collect_batch(h, true, &r);
    const unsigned long long bits = h->ws.h_totals[10];            // error bits of every batch since the last wait
    CUDA_TRY(h, cudaMemsetAsync(h->queue_dev + 1, 0, 8, h->s_compute));
    if (rc == DWGSIM_GPU_OK && (bits & 1ull)) { h->last_error = "failed to generate a read after 10001 trials"; rc = DWGSIM_GPU_ETRIALS; }
    if (rc == DWGSIM_GPU_OK && (bits & 2ull)) { h->last_error = "Ion Torrent read grew past 2*len+64 bases"; rc = DWGSIM_GPU_EOVERFLOW; }
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 3; ++k) { out->dev_ptr[k] = (uint64_t)(uintptr_t)h->ws.out[0][k]; out->n_bytes[k] = r.bytes[k]; h->last_bytes[k] = r.bytes[k]; }
    h->last_slot = 0;
    out->n_pairs = h->queued_n; out->n_random = r.n_random; out->n_failed_attempts = r.n_failed;
    out->ms_simulate = r.ms[0]; out->ms_layout = r.ms[1]; out->ms_format = r.ms[2];
    out->n_launches = h->queued_launches;
    h->queued_launches = 0; h->queued_n = 0;
    return rc;
}

int dwgsim_gpu_resident_count_ptr(dwgsim_gpu_t *h, uint64_t *count_device_ptr)
{
    if (!h || !count_device_ptr || !h->ws.totals) return DWGSIM_GPU_ESTATE;
    *count_device_ptr = (uint64_t)(uintptr_t)h->ws.totals;        // totals[0]: written by the scan that follows the simulate passes
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_resident_finish_dev(dwgsim_gpu_t *h, uint64_t rand_serial_base_device_ptr, dwgsim_gpu_batch_t *out)
{
    if (!rand_serial_base_device_ptr) return DWGSIM_GPU_EINVAL;
    return resident_finish(h, 0, reinterpret_cast<const unsigned long long *>((uintptr_t)rand_serial_base_device_ptr), out);
}

int dwgsim_gpu_simulate_resident(dwgsim_gpu_t *h, int64_t first, int64_t n, int64_t rand_serial_base, dwgsim_gpu_batch_t *out)
{
    if (!out) return DWGSIM_GPU_EINVAL;
    int rc = dwgsim_gpu_resident_begin(h, first, n, nullptr);      // no host sync between the phases
    if (rc) return rc;
    return dwgsim_gpu_resident_finish(h, rand_serial_base, out);
}

int dwgsim_gpu_set_exchange(dwgsim_gpu_t *h, dwgsim_gpu_exchange_fn fn, void *user)
{
    if (!h) return DWGSIM_GPU_EINVAL;
    h->exchange = fn; h->exchange_user = user;
    return DWGSIM_GPU_OK;
}

int dwgsim_gpu_copy_stream(dwgsim_gpu_t *h, int file_id, char *dst, uint64_t cap)
{
    if (!h || file_id < 0 || file_id > 2 || !dst) return DWGSIM_GPU_EINVAL;
    if (cap < h->last_bytes[file_id]) return DWGSIM_GPU_EINVAL;
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMemcpy(dst, h->ws.out[h->last_slot][file_id], h->last_bytes[file_id], cudaMemcpyDeviceToHost));
    return DWGSIM_GPU_OK;
}

namespace {
int run_group(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats);
int run_one(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats);
}  // namespace

int dwgsim_gpu_run(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats)
{
    if (!h || !sink) return DWGSIM_GPU_EINVAL;
    return h->peers.empty() ? run_one(h, sink, user, stats) : run_group(h, sink, user, stats);
}

namespace {
int run_one(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats)
{
    cudaSetDevice(h->device);
    const double t_start = now_ms();
    dwgsim_gpu_stats_t st;
    memset(&st, 0, sizeof st);
    const int64_t h2d_before = h->h2d_bytes;
    int rc = finalize_genome(h);
    if (rc) return rc;
    const int64_t total = h->blob_pairs;
    const int64_t B = std::min<int64_t>(h->batch_pairs, std::max<int64_t>(total, 1));
    if ((rc = ensure_workspace(h, B, true))) { free_blob(h); return rc; }
    Workspace &w = h->ws;
    const int world = h->shard_world, rank = h->shard_rank;
    if (world > 1 && !h->exchange) { h->last_error = "sharded run needs dwgsim_gpu_set_exchange"; free_blob(h); return DWGSIM_GPU_ESTATE; }
    cudaEvent_t copied[8];
    for (int s = 0; s < h->pinned_slots; ++s) cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming);
    cudaEvent_t computed;
    cudaEventCreateWithFlags(&computed, cudaEventDisableTiming);
    // every exit below goes through the clean-up at the end (events, copy stream, genome, counters)
#define RUN_TRY(expr)                                                                                          \
    {                                                                                                          \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            h->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                               \
            rc = _e == cudaErrorMemoryAllocation ? DWGSIM_GPU_ENOMEM : DWGSIM_GPU_ECUDA;                       \
            break;                                                                                             \
        }                                                                                                      \
    }
    // Batches whose device-to-host copy is in flight or finished but not handed to the sink yet, oldest first.  Two of them
    // keep the copy engine busy: the copy of batch j is enqueued while that of batch j-1 still runs, and the sink gets batch
    // j-2 meanwhile (with a ring of two pinned slots: one).
    struct Pending { uint64_t bytes[3] = {0, 0, 0}; int pslot = 0; int64_t batch = 0; };
    std::vector<Pending> pend;
    const size_t depth = h->pinned_slots >= 3 ? 2 : 1;
    GroupSync *const grp = h->group;                            // ranks of a device group: the sink sees the batches in order
    double t_copy_wait = 0, t_turn_wait = 0, t_sink = 0, t_collect = 0, t_gz = 0, t_launch = 0;   // DWGSIM_RUN_TIMING
    auto drain_one = [&]() -> int {
        const Pending pd = pend.front();
        pend.erase(pend.begin());
        double t0 = now_ms();
        if (cudaEventSynchronize(copied[pd.pslot]) != cudaSuccess) { h->last_error = "device to host copy failed"; return DWGSIM_GPU_ECUDA; }
        double t1 = now_ms();
        t_copy_wait += t1 - t0;
        if (grp && !grp->wait_turn(pd.batch)) { h->last_error = "another device of the group failed"; return DWGSIM_GPU_ESTATE; }
        t0 = now_ms();
        t_turn_wait += t0 - t1;
        for (int k = 0; k < 3; ++k)
            if (pd.bytes[k]) {
                if (sink(user, k, h->pinned[pd.pslot][k], (size_t)pd.bytes[k])) { h->last_error = "sink callback failed"; return DWGSIM_GPU_ESINK; }
                st.bytes[k] += (int64_t)pd.bytes[k];
            }
        if (grp) grp->pass_turn();
        t_sink += now_ms() - t0;
        return DWGSIM_GPU_OK;
    };
    auto drain_to = [&](size_t keep) -> int { int r = DWGSIM_GPU_OK; while (pend.size() > keep && r == DWGSIM_GPU_OK) r = drain_one(); return r; };
    int launches = 0;
    // batch b covers the pairs [starts[b], starts[b + 1]).  A run starts with shorter batches (B/4, B/2, then B): the first
    // device-to-host copy, which everything after it queues behind, begins after a quarter of a batch's kernel time.  The
    // devices of a group follow the same schedule (their output is that of one device, gzip members included); runs sharded
    // across processes keep equal batches: batch index -> rank is part of the exchange protocol (dwgsim_b200/shard.py).
    std::vector<int64_t> starts;
    {
        static const bool no_ramp = getenv("DWGSIM_NO_RAMP") != nullptr;
        int64_t sz = ((world == 1 || h->group) && !no_ramp) ? std::max<int64_t>(B / 4, 1) : B;
        for (int64_t at = 0; at < total; at += std::min(sz, total - at), sz = std::min(B, sz * 2)) starts.push_back(at);
        starts.push_back(total);
    }
    const int64_t n_batches = (int64_t)starts.size() - 1;
    const int64_t n_rounds = (n_batches + world - 1) / world;
    int64_t mine = 0;                                           // batches this rank has processed
    for (int64_t round = 0; round < n_rounds && rc == DWGSIM_GPU_OK; ++round) {
        const int64_t bi = round * world + rank;               // batch index owned by this rank in this round
        const bool active = bi < n_batches;
        const int64_t first = active ? starts[(size_t)bi] : total;
        const int n = active ? (int)(starts[(size_t)bi + 1] - first) : 0;
        const int dslot = (int)(mine & 1), pslot = (int)(mine % h->pinned_slots);
        int l = 0;
        int64_t rand_base = h->rand_serial, my_random = 0, round_total = 0;
        // the device slot was last used two batches ago: its copy to the host has to be over before the kernels write it again
        if (active && mine >= 2) RUN_TRY(cudaStreamWaitEvent(h->s_compute, copied[(int)((mine - 2) % h->pinned_slots)], 0));
        if (world == 1) {
            if ((rc = launch_batch(h, first, n, rand_base, dslot, true, &l))) break;
            launches += l;
            if ((rc = drain_to(depth - 1))) break;             // the sink works on an earlier batch meanwhile
        } else {
            if (active) {
                if ((rc = launch_simulate(h, first, n, true, &l))) break;
                launches += l;
            }
            if ((rc = drain_to(depth - 1))) break;
            if (active && (rc = read_random_count(h, &my_random))) break;
            // the one exchange of the path: random-pair counts of this round, in batch order
            int64_t before_me = 0;
            if (h->exchange(h->exchange_user, round, my_random, &before_me, &round_total)) { h->last_error = "exchange callback failed"; rc = DWGSIM_GPU_ESINK; break; }
            rand_base = h->rand_serial + before_me;
            if (active) {
                if ((rc = launch_format(h, first, n, rand_base, dslot, true, &l))) break;
                launches += l;
            }
        }
        if (!active) { h->rand_serial += round_total; continue; }
        BatchResult r;
        const double tc0 = now_ms();
        if ((rc = collect_batch(h, true, &r))) break;
        t_collect += now_ms() - tc0;
        st.ms_simulate += r.ms[0]; st.ms_layout += r.ms[1]; st.ms_format += r.ms[2];
        st.n_random += r.n_random; st.n_failed_attempts += r.n_failed; st.n_pairs += n;
        h->rand_serial += world == 1 ? r.n_random : round_total;
        uint64_t send[3] = {r.bytes[0], r.bytes[1], r.bytes[2]};
        for (int k = 0; k < 3; ++k) st.raw_bytes[k] += (int64_t)r.bytes[k];
        if (h->gz_mode) {                                          // gzip members on the device: fewer bytes over PCIe
            int lz = 0;
            const double tg0 = now_ms();
            if ((rc = gz_batch(h, dslot, r.bytes, send, &lz))) break;
            t_gz += now_ms() - tg0;
            launches += lz;
        }
        bool grow = false;
        for (int k = 0; k < 3; ++k) grow = grow || send[k] > h->pinned_cap[k];
        if (grow) {
            // a batch compressed worse than the ring was sized for: hand over what is pending, then page-lock larger slots
            if ((rc = drain_to(0))) break;
            cudaStreamSynchronize(h->s_copy);
            for (int k = 0; k < 3; ++k) {
                if (send[k] <= h->pinned_cap[k]) continue;
                h->pinned_cap[k] = align_up(send[k] + send[k] / 4, 4096);
                for (int s2 = 0; s2 < h->pinned_slots && rc == DWGSIM_GPU_OK; ++s2) {
                    cudaFreeHost(h->pinned[s2][k]); h->pinned[s2][k] = nullptr;
                    if (cudaMallocHost((void **)&h->pinned[s2][k], h->pinned_cap[k]) != cudaSuccess) { h->last_error = "out of pinned host memory"; rc = DWGSIM_GPU_ENOMEM; }
                }
            }
            if (rc) break;
        }
        RUN_TRY(cudaEventRecord(computed, h->s_compute));
        RUN_TRY(cudaStreamWaitEvent(h->s_copy, computed, 0));
        bool copy_failed = false;
        for (int k = 0; k < 3 && !copy_failed; ++k)
            if (send[k]) {
                if (cudaMemcpyAsync(h->pinned[pslot][k], h->gz_mode ? w.gz_out[dslot][k] : w.out[dslot][k], send[k],
                                    cudaMemcpyDeviceToHost, h->s_copy) != cudaSuccess) copy_failed = true;
                st.d2h_bytes += (int64_t)send[k];
            }
        if (copy_failed) { h->last_error = "device to host copy failed"; rc = DWGSIM_GPU_ECUDA; break; }
        RUN_TRY(cudaEventRecord(copied[pslot], h->s_copy));
        Pending pd;
        pd.pslot = pslot; pd.batch = bi;
        for (int k = 0; k < 3; ++k) pd.bytes[k] = send[k];
        pend.push_back(pd);
        ++st.n_batches; ++mine;
    }
#undef RUN_TRY
    if (rc == DWGSIM_GPU_OK) rc = drain_to(0);
    cudaStreamSynchronize(h->s_copy);
    cudaStreamSynchronize(h->s_compute);
    for (int s = 0; s < h->pinned_slots; ++s) cudaEventDestroy(copied[s]);
    cudaEventDestroy(computed);
    h->gidx_origin += total;
    free_blob(h);
    st.n_launches = launches;
    st.h2d_bytes = h->h2d_bytes - h2d_before;
    st.ms_pack = h->ms_pack; h->ms_pack = 0;
    st.ms_compress = h->ms_gz; h->ms_gz = 0;
    st.ms_total = now_ms() - t_start;
    if (getenv("DWGSIM_RUN_TIMING"))
        fprintf(stderr, "[dwgsim_gpu_run] rank %d: %lld pairs in %d batches, %.1f ms: waits copy %.1f turn %.1f, sink %.1f, collect %.1f, gz %.1f (launch %.1f) | kernels %.1f gz kernels %.1f | d2h %.1f MB\n",
                rank, (long long)st.n_pairs, st.n_batches, st.ms_total, t_copy_wait, t_turn_wait, t_sink, t_collect, t_gz, t_launch,
                st.ms_simulate + st.ms_layout + st.ms_format, st.ms_compress, st.d2h_bytes / 1e6);
    if (stats) *stats = st;
    return rc;
}

int group_exchange(void *user, int64_t round, int64_t my_random, int64_t *before_me, int64_t *round_total)
{
    (void)round;
    dwgsim_gpu *h = (dwgsim_gpu *)user;
    return h->group->exchange(h->shard_rank, my_random, before_me, round_total) ? 0 : 1;
}

// dwgsim_gpu_run of a device group: the leader packs and uploads the genome, the peers receive a copy of the blob device to
// device (NVLink when the devices are peers), then every device runs its share of the batches (b % world == rank) on its
// own host thread; the sink sees the batches in order, so the bytes are those of a single-device run.
int run_group(dwgsim_gpu_t *h, dwgsim_gpu_sink_fn sink, void *user, dwgsim_gpu_stats_t *stats)
{
    const int world = 1 + (int)h->peers.size();
    cudaSetDevice(h->device);
    const double t_start = now_ms();
    int rc = finalize_genome(h);
    if (rc) return rc;
    for (dwgsim_gpu *q : h->peers) {                                // the blob, device to device
        cudaSetDevice(q->device);
        free_blob(q);
        uint8_t *dst = nullptr;
        if (q->blob_spare && q->blob_spare_cap >= h->blob_bytes) { dst = q->blob_spare; q->blob_spare = nullptr; q->blob_spare_cap = 0; }
        else if (cudaMalloc((void **)&dst, h->blob_bytes) != cudaSuccess) { h->last_error = "out of device memory for the genome copy"; return DWGSIM_GPU_ENOMEM; }
        cudaError_t e = q->device == h->device ? cudaMemcpyAsync(dst, h->blob, h->blob_bytes, cudaMemcpyDeviceToDevice, q->s_compute)
                                               : cudaMemcpyPeerAsync(dst, q->device, h->blob, h->device, h->blob_bytes, q->s_compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(q->s_compute);
        if (e != cudaSuccess) { cudaFree(dst); h->last_error = std::string("peer copy of the genome: ") + cudaGetErrorString(e); return DWGSIM_GPU_ECUDA; }
        q->max_name_len = std::max(q->max_name_len, h->max_name_len);
        if ((rc = dwgsim_gpu_genome_import(q, (uint64_t)(uintptr_t)dst, h->blob_bytes, 1))) { h->last_error = q->last_error; return rc; }
        q->gidx_origin = h->gidx_origin; q->rand_serial = h->rand_serial;
        q->batch_pairs = h->batch_pairs;
    }
    cudaSetDevice(h->device);
    GroupSync sync;
    sync.reset(world);
    sync.leader = h;
    std::vector<dwgsim_gpu *> ranks{h};
    ranks.insert(ranks.end(), h->peers.begin(), h->peers.end());
    std::vector<int> rcs((size_t)world, DWGSIM_GPU_OK);
    std::vector<dwgsim_gpu_stats_t> sts((size_t)world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r) {
        dwgsim_gpu *q = ranks[(size_t)r];
        q->group = &sync; q->shard_rank = r; q->shard_world = world;
        q->exchange = group_exchange; q->exchange_user = q;
    }
    for (int r = 1; r < world; ++r)
        th.emplace_back([&, r]() { rcs[(size_t)r] = run_one(ranks[(size_t)r], sink, user, &sts[(size_t)r]); if (rcs[(size_t)r]) sync.fail(); });
    rcs[0] = run_one(h, sink, user, &sts[0]);
    if (rcs[0]) sync.fail();
    for (auto &t : th) t.join();
    dwgsim_gpu_stats_t st = sts[0];
    rc = rcs[0];
    for (int r = 0; r < world; ++r) {
        dwgsim_gpu *q = ranks[(size_t)r];
        q->group = nullptr; q->shard_rank = 0; q->shard_world = 1; q->exchange = nullptr; q->exchange_user = nullptr;
        if (rcs[(size_t)r] && (rc == DWGSIM_GPU_OK || rc == DWGSIM_GPU_ESTATE)) { rc = rcs[(size_t)r]; h->last_error = q->last_error; }
        if (r == 0) continue;
        const dwgsim_gpu_stats_t &s2 = sts[(size_t)r];
        st.n_pairs += s2.n_pairs; st.n_random += s2.n_random; st.n_failed_attempts += s2.n_failed_attempts;
        for (int k = 0; k < 3; ++k) { st.bytes[k] += s2.bytes[k]; st.raw_bytes[k] += s2.raw_bytes[k]; }
        st.d2h_bytes += s2.d2h_bytes; st.n_launches += s2.n_launches; st.n_batches += s2.n_batches;
        // device time: the slowest device bounds the run
        st.ms_simulate = std::max(st.ms_simulate, s2.ms_simulate); st.ms_layout = std::max(st.ms_layout, s2.ms_layout);
        st.ms_format = std::max(st.ms_format, s2.ms_format); st.ms_compress = std::max(st.ms_compress, s2.ms_compress);
    }
    cudaSetDevice(h->device);
    st.ms_total = now_ms() - t_start;
    if (stats) *stats = st;
    return rc;
}
}  // namespace

// ---- synthetic genome (benchmarks): built procedurally on the host, then packed like any other ---------
namespace {
struct SplitMix {
    uint64_t s;
    explicit SplitMix(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double unif() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

void synth_contig(HostContig &c, uint64_t seed, double mut_rate, double indel_frac, double n_frac)
{
    const int len = c.len;
    const size_t nw2 = ((size_t)len + 15) / 16 + 1, nwm = ((size_t)len + 31) / 32 + 1;
    c.ref2.resize(nw2); c.nmask.assign(nwm, 0);
    SplitMix rng(seed * 0x100000001B3ull + (uint64_t)c.contig_i * 0x9E3779B97F4A7C15ull + 1);
    for (size_t w = 0; w + 1 < nw2; w += 2) { uint64_t r = rng.next(); c.ref2[w] = (uint32_t)r; c.ref2[w + 1] = (uint32_t)(r >> 32); }
    c.ref2[nw2 - 1] = (uint32_t)rng.next();
    // N runs: 10 kb telomeres and one long run a third of the way in (together ~n_frac of the contig)
    auto set_n = [&](int64_t a, int64_t b) {
        a = std::max<int64_t>(a, 0); b = std::min<int64_t>(b, len);
        for (int64_t p = a; p < b; ++p) { c.nmask[p >> 5] |= 1u << (p & 31); c.ref2[p >> 4] &= ~(3u << ((p & 15) << 1)); }
    };
    if (n_frac > 0 && len > 100000) {
        const int64_t tel = 10000, mid = std::max<int64_t>((int64_t)(n_frac * len) - 2 * tel, 0);
        set_n(0, tel); set_n(len - tel, len); set_n(len / 3, len / 3 + mid);
    }
    auto is_n = [&](int64_t p) { return (c.nmask[p >> 5] >> (p & 31)) & 1u; };
    auto base_at = [&](int64_t p) { return (c.ref2[p >> 4] >> ((p & 15) << 1)) & 3u; };
    const int nblk = (len >> kBlkShift) + 2;
    for (int hh = 0; hh < 2; ++hh) { c.ev[hh].clear(); c.pool[hh].assign(4, 0); c.blk[hh].assign((size_t)nblk, 0); }
    if (mut_rate > 0) {
        const double lg = log1p(-std::min(mut_rate, 0.999999));
        int64_t p = -1;
        for (;;) {
            double u = rng.unif();
            if (u <= 0) u = 1e-300;
            p += 1 + (int64_t)floor(log(u) / lg);
            if (p >= len) break;
            if (is_n(p)) continue;
            const uint64_t r = rng.next();
            const int which = (r & 0xff) < 85 ? 3 : (((r >> 8) & 1) ? 1 : 2);     // 1/3 homozygous
            const uint32_t rb = base_at(p);
            if (rng.unif() >= indel_frac) {
                Event e{(uint32_t)p, kEvSubst | (((rb + 1 + (uint32_t)((r >> 16) % 3)) & 3u) << 2), 0};
                for (int hh = 0; hh < 2; ++hh) if (which & (1 << hh)) c.ev[hh].push_back(e);
            } else if ((r >> 20) & 1) {
                int L = 1;
                while (rng.unif() < 0.3) ++L;
                int64_t q = p;
                for (; q < p + L && q < len && !is_n(q); ++q) {
                    Event e{(uint32_t)q, kEvDelete | (base_at(q) << 2), 0};
                    for (int hh = 0; hh < 2; ++hh) if (which & (1 << hh)) c.ev[hh].push_back(e);
                }
                p = q - 1;
            } else {
                uint32_t n = 1;
                while (rng.unif() < 0.3 && n < 26) ++n;
                Event e{(uint32_t)p, kEvInsert | (rb << 2) | (n << 5), rng.next() & ((n >= 32) ? ~0ull : ((1ull << (2 * n)) - 1))};
                for (int hh = 0; hh < 2; ++hh) if (which & (1 << hh)) c.ev[hh].push_back(e);
            }
        }
    }
    for (int hh = 0; hh < 2; ++hh) {
        size_t e = 0;
        for (int b = 0; b < nblk; ++b) {
            const uint64_t start = (uint64_t)b << kBlkShift;
            while (e < c.ev[hh].size() && c.ev[hh][e].pos < start) ++e;
            c.blk[hh][b] = (uint32_t)e;
        }
    }
}

struct CountSink { int64_t bytes[3]; int64_t calls; uint64_t first_bytes[3]; };
}  // namespace

int dwgsim_gpu_genome_synthetic(dwgsim_gpu_t *h, int32_t n_contigs, const int32_t *lengths, uint64_t seed, double mut_rate,
                                double indel_frac, double n_frac, double coverage)
{
    if (!h || n_contigs < 1 || !lengths || mut_rate < 0 || mut_rate >= 1 || coverage < 0) return DWGSIM_GPU_EINVAL;
    if (h->blob) { h->last_error = "genome already finalized"; return DWGSIM_GPU_ESTATE; }
    const double t0 = now_ms();
    const size_t base = h->queue.size();
    h->queue.resize(base + (size_t)n_contigs);
    for (int i = 0; i < n_contigs; ++i) {
        HostContig &c = h->queue[base + i];
        c.name = "chr" + std::to_string(i + 1);
        c.contig_i = (int32_t)(base + i); c.len = lengths[i];
        // pair budget by coverage, src/dwgsim.c:589
        c.n_pairs = (int64_t)(uint64_t)(c.len * coverage / ((long double)(h->p.length[0] + h->p.length[1])) / (1.0 - h->p.rand_read) + 0.5);
        h->max_name_len = std::max(h->max_name_len, (int)c.name.size());
    }
    std::vector<std::thread> th;
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)n_contigs));
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (int i = (int)t; i < n_contigs; i += (int)nt) synth_contig(h->queue[base + i], seed, mut_rate, indel_frac, n_frac); });
    for (auto &x : th) x.join();
    h->ms_pack += now_ms() - t0;
    return DWGSIM_GPU_OK;
}

// ---- host reference of the device gzip member format (CPU tests of the Huffman / CRC tables) ------------------
int dwgsim_gpu_gz_host_encode(const uint8_t *data, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n)
{
    if ((!data && n) || !out || !out_n) return DWGSIM_GPU_EINVAL;
    uint64_t hist[256] = {0};
    for (uint64_t i = 0; i < n; ++i) hist[data[i]]++;
    const GzTables t = gz_build_tables(hist);
    static const Crc32Tables ct;
    std::vector<uint8_t> v;
    gz_encode_host(t, ct, data, (size_t)n, v);
    if (v.size() > cap) return DWGSIM_GPU_EINVAL;
    memcpy(out, v.data(), v.size());
    *out_n = v.size();
    return DWGSIM_GPU_OK;
}

// ---- ready-made sinks (plain C callbacks a host can pass to dwgsim_gpu_run) --------------------------------
// user = int64_t[4]: bytes per file id, number of calls
int dwgsim_gpu_sink_count(void *user, int file_id, const char *buf, size_t n)
{
    int64_t *c = (int64_t *)user;
    (void)buf;
    c[file_id] += (int64_t)n; c[3] += 1;
    return 0;
}
// user = int[3]: one file descriptor per file id (-1 = discard)
int dwgsim_gpu_sink_fd(void *user, int file_id, const char *buf, size_t n)
{
    const int fd = ((const int *)user)[file_id];
    if (fd < 0) return 0;
    while (n) {
        ssize_t w = write(fd, buf, n);
        if (w <= 0) return 1;
        buf += w; n -= (size_t)w;
    }
    return 0;
}

// ---- file sink: one writer thread per file -------------------------------------------------------------------
// Copying into the page cache is what bounds a file sink (a few GB/s per file: buffered writes to one file serialise
// on its inode lock), so the three files of a batch are written side by side and while the device works on the next
// batch: a chunk is handed to the file's writer and joined at the next chunk for that file (the buffer stays valid
// until then, see dwgsim_gpu_sink_fn) or by dwgsim_gpu_file_sink_close.
int dwgsim_gpu_pwrite_all(int fd, const char *buf, size_t n, int64_t offset)
{
    if (fd < 0) return 0;
    while (n) {
        ssize_t w = pwrite(fd, buf, n, (off_t)offset);
        if (w < 0 && errno == ESPIPE) w = write(fd, buf, n);         // not seekable (pipe, character device)
        if (w <= 0) return 1;
        buf += w; n -= (size_t)w; offset += w;
    }
    return 0;
}

struct dwgsim_gpu_file_sink {
    int fd[3];
    int64_t offset[3];
    std::thread writer[3];
    int status[3];
};

dwgsim_gpu_file_sink_t *dwgsim_gpu_file_sink_open(const int32_t fd[3], const int64_t offset[3])
{
    dwgsim_gpu_file_sink *f = new dwgsim_gpu_file_sink();
    for (int k = 0; k < 3; ++k) { f->fd[k] = fd ? fd[k] : -1; f->offset[k] = offset ? offset[k] : 0; f->status[k] = 0; }
    return f;
}
static int file_sink_join(dwgsim_gpu_file_sink *f, int k)
{
    if (f->writer[k].joinable()) f->writer[k].join();
    return f->status[k];
}
int dwgsim_gpu_sink_files(void *user, int file_id, const char *buf, size_t n)
{
    dwgsim_gpu_file_sink *f = (dwgsim_gpu_file_sink *)user;
    if (!f || file_id < 0 || file_id > 2) return 1;
    if (file_sink_join(f, file_id)) return 1;
    const int fd = f->fd[file_id];
    const int64_t at = f->offset[file_id];
    f->offset[file_id] += (int64_t)n;
    if (fd < 0 || n == 0) return 0;
    int *st = &f->status[file_id];
    f->writer[file_id] = std::thread([fd, buf, n, at, st]() { if (dwgsim_gpu_pwrite_all(fd, buf, n, at)) *st = 1; });
    return 0;
}
int dwgsim_gpu_file_sink_close(dwgsim_gpu_file_sink_t *f, int64_t offset_out[3])
{
    if (!f) return 0;
    int rc = 0;
    for (int k = 0; k < 3; ++k) { if (file_sink_join(f, k)) rc = 1; if (offset_out) offset_out[k] = f->offset[k]; }
    delete f;
    return rc;
}

}  // extern "C"
