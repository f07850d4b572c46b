// sparse_search.cu — host driver of the sparse search (the C++ side of cpp/sparse.cpp's
// FullState(paulis).evolve(), cpp/state.cpp:636-687).  The host only orchestrates: it uploads the term
// table, launches the kernels of sparse_kernels.cuh and reads back one record.  There is no CPU
// search path; without a CUDA device every entry point fails with SGS_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.hpp"
#include "hamiltonian.hpp"
#include "nccl_glue.hpp"
#include "sparse_kernels.cuh"
#include "arena.hpp"
#include "sparse_search.hpp"

namespace sgs {

namespace {

struct DevCtx {
    int device = 0;
    int shard = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evh0 = nullptr, evh1 = nullptr, evA = nullptr, evB = nullptr;
    int sms = 0;
    // device buffers
    uint64_t *rows = nullptr, *adj = nullptr;
    double *wgt = nullptr, *wabs = nullptr;
    int *order = nullptr;
    BestRec *wbest = nullptr, *gbest = nullptr;
    ShardRec *shardrec = nullptr, *allrec = nullptr;   // this shard's contribution; every shard's (after the all-gather)
    SparseStats *stats = nullptr;
    uint16_t *qA = nullptr, *qB = nullptr, *slow = nullptr, *work = nullptr;
    unsigned int *counters = nullptr;   // [0] next-level count, [1] item counter, [2] slow count, [8..15] root class counts, [16..23] cursors
    unsigned int *perm = nullptr;       // [qcap] heavy-first order of the subtree roots
    unsigned long long *inc = nullptr;  // branch-and-bound mode: incumbent energy key
    uint16_t *retry[2] = {nullptr, nullptr};   // [retry_cap * S] roots of the 2nd / 3rd k_dfs round (counters[24..27]: count, valid end)
    FinalOut *fin = nullptr;
    int8_t *term_signs = nullptr;
    uint64_t *memX = nullptr;
    double *memA = nullptr;
    int *memT = nullptr, *memCount = nullptr;
    int nwbest = 0;
    unsigned long long *wtime = nullptr;
    void *arena = nullptr;
    // results of the local search
    SparseStats hstats{};
    BestRec hbest{};
    float ms_device = 0.f, ms_hot = 0.f;
    unsigned launches = 0;
    bool had_slow = false;
    Status status;
};


}  // namespace


// order-encoded +1e300 (enc_double in sparse_kernels.cuh): "no incumbent yet" of the branch-and-bound mode
static const unsigned long long kNoIncumbent = [] {
    const double x = 1e300;
    unsigned long long u;
    memcpy(&u, &x, 8);
    return u | 0x8000000000000000ULL;
}();

struct SparsePlan::Impl {
    Hamiltonian H;
    sgs_opts opts{};
    int W = 1, T = 0, TW = 0, S = 0, n = 0;
    int nshards = 1;
    unsigned qcap = 0, slow_cap = 0, chunk = 0;
    unsigned retry_cap = 1u << 20;      // roots handed to the next k_dfs round (multi-shard runs); SGS_RETRY_CAP for tests
    int dfs_cap = 0;
    bool no_dfs_expand = false;         // the expansion-mode launch overflowed its queue once: use k_general for every level
    std::vector<DevCtx> devs;
    std::shared_ptr<NcclWorld> nccl;
    // Replay: the launch sequence of a first (adaptive, host-synchronised) run that needed no work from
    // the exact kernel below the subtree roots is deterministic, so later runs of the same plan enqueue it
    // back to back — expansion levels, walk, reduction, collective, finalisation — and synchronise once.
    struct Replay {
        bool ok = false;
        int depth = 0;
        std::vector<unsigned> counts;    // nodes per expanded level; counts[depth] = subtree roots
        int dfs_cap = 0, dfs_wpb = 8, gen_wpb = 4;
        size_t dfs_pw = 0, gen_pw = 0;
        unsigned dfs_grid = 0, gen_maxgrid = 0;
        // last expansion level done by the walk kernel in expansion mode (k_dfs, kDfsExpand)
        bool dfs_expand = false;
        int exp_cap = 0, exp_wpb = 4;
        size_t exp_pw = 0;
        unsigned exp_grid = 0;
        unsigned long long roots_per_shard = 0;      // the walk's proxy for its own length (hand-off rounds or not): same in both launch sequences
    } rp;

    ~Impl();
    Status init();
    Status init_device(DevCtx &c);
    template <int W_> Status search_local(DevCtx &c);
    template <int W_> Status launch_dfs(DevCtx &c, const SparseDev &D, unsigned grid, int wpb, size_t pw, int cap, const uint16_t *items,
                                        unsigned count, int nsh, const unsigned *count_dev, const unsigned *perm, const unsigned *valid0,
                                        unsigned long long roots_per_shard);
    template <int W_> Status launch_expand(DevCtx &c, const SparseDev &D, unsigned grid, int wpb, size_t pw, int cap, const uint16_t *items,
                                           unsigned count, uint16_t *out);
    template <int W_> Status finalize(DevCtx &c, const ShardRec *src, int nrec, sgs_result *out, SparseStats *tot, BestRec *best);
    template <int W_> Status run_replay(sgs_result *out);
    void fill_stats(const SparseStats &tot, sgs_result *out);
    void fill_empty(const SparseStats &tot, sgs_result *out);
    Status run(sgs_result *out);
    SparseDev dev_view(const DevCtx &c) const;
};

SparsePlan::Impl::~Impl()
{
    nccl.reset();
    for (auto &c : devs) {
        cudaSetDevice(c.device);
        if (c.arena) sparse_arena().release(c.device, c.arena);
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.evh0) cudaEventDestroy(c.evh0);
        if (c.evh1) cudaEventDestroy(c.evh1);
        if (c.evA) cudaEventDestroy(c.evA);
        if (c.evB) cudaEventDestroy(c.evB);
        if (c.stream) cudaStreamDestroy(c.stream);
    }
}

SparseDev SparsePlan::Impl::dev_view(const DevCtx &c) const
{
    SparseDev D{};
    D.n = n; D.T = T; D.TW = TW; D.S = S;
    D.wtime = c.wtime;
    D.rows = c.rows; D.wgt = c.wgt; D.wabs = c.wabs; D.adj = c.adj;
    D.wbest = c.wbest; D.stats = c.stats;
    D.slow = c.slow; D.slow_count = c.counters + 2; D.slow_cap = slow_cap;
    D.trace = getenv("SGS_TRACE") ? 1 : 0;
    D.inc = opts.prune ? c.inc : nullptr;
    D.split_cliques = kSplitCliques;
    if (const char *v = getenv("SGS_SPLIT_CLIQUES")) if (atof(v) > 0) D.split_cliques = (float)atof(v);
    return D;
}

Status SparsePlan::Impl::init()
{
    n = H.n; T = H.T();
    if (H.wide || n + 1 > SGS_MAX_QUBITS) return Status::fail(SGS_ERR_LIMIT, "sparse search supports n <= 127 (the k-local DP takes any n)");
    if (T > 65535) return Status::fail(SGS_ERR_LIMIT, "sparse search supports at most 65535 terms");
    if (T > 8192) return Status::fail(SGS_ERR_LIMIT, "sparse search supports at most 8192 terms (shared-memory lists)");
    {
        // Input check for the rank limit (kMaxG generators per group).  A greedy pass collects terms that commute
        // pairwise and are linearly independent; if it gathers more than kMaxG, groups of a higher rank exist, the
        // tree has more than 2^kMaxG leaves and the search could neither finish nor represent its answer: refuse
        // now.  (Instances that pass may still meet the limit during the search; the kernels then stop with the
        // same error instead of writing a record that does not fit.)
        std::vector<PackedRow> chosen, basis;
        auto anticommute = [](const PackedRow &a, const PackedRow &b) {
            uint64_t acc = 0;
            for (int k = 0; k < kMaxWords; k++) {
                const uint64_t sw = ((b.w[k] >> 1) & 0x5555555555555555ULL) | ((b.w[k] & 0x5555555555555555ULL) << 1);
                acc ^= a.w[k] & sw;
            }
            return __builtin_popcountll(acc) & 1;
        };
        for (int p = 0; p < T && (int)chosen.size() <= kMaxG; p++) {
            const PackedRow &t = H.terms[H.order[p]];
            bool ok = true;
            for (const PackedRow &c : chosen) if (anticommute(t, c)) { ok = false; break; }
            if (!ok) continue;
            PackedRow r = t;
            for (const PackedRow &b : basis) {          // basis rows are reduced: each has a pivot no other row has
                int pw = 0;
                while (pw < kMaxWords && b.w[pw] == 0) pw++;
                const uint64_t pb = b.w[pw] & (0 - b.w[pw]);
                if (r.w[pw] & pb) for (int k = 0; k < kMaxWords; k++) r.w[k] ^= b.w[k];
            }
            bool zero = true;
            for (int k = 0; k < kMaxWords; k++) zero = zero && r.w[k] == 0;
            if (zero) continue;
            chosen.push_back(t);
            basis.push_back(r);
        }
        if ((int)chosen.size() > kMaxG)
            return Status::fail(SGS_ERR_LIMIT, "sparse search: the Hamiltonian has more than 64 independent pairwise-commuting terms "
                                               "(groups of rank > 64; the search space exceeds 2^64 groups)");
    }
    W = (2 * n + 63) / 64;
    if (W == 3) W = 4;
    TW = (T + 63) / 64;
    if (TW < 1) TW = 1;
    const int maxrank = std::min(std::min(n, std::max(T, 1)), kMaxG);
    S = 2 + maxrank;
    if (S & 1) S++;
    const unsigned long long want = std::max<unsigned long long>(1ull << 18, 64ull * T * T);
    qcap = (unsigned)std::min<unsigned long long>(want, 1ull << 25);
    slow_cap = 1u << 22;
    chunk = 1u << 16;
    if (const char *v = getenv("SGS_RETRY_CAP")) if (atoi(v) > 0) retry_cap = (unsigned)atoi(v);
    dfs_cap = 128;     // re-derived per search from the measured branching of the last expanded level

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return Status::fail(SGS_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                                   " (libsgs has no CPU search path)");
    const int ng = std::max(1, opts.ngpus);
    if (opts.device < 0 || opts.device + ng > ndev) return Status::fail(SGS_ERR_ARG, "device ordinal out of range");
    const int world = std::max(1, opts.world);
    nshards = world * ng;
    devs.resize(ng);
    for (int i = 0; i < ng; i++) {
        devs[i].device = opts.device + i;
        devs[i].shard = std::max(0, opts.rank) * ng + i;
        Status s = init_device(devs[i]);
        if (!s.ok()) return s;
    }
    if (nshards > 1 && (opts.nccl_id != nullptr || (world == 1 && ng > 1))) {
        // communicators are cached for the life of the process, keyed by (unique id, rank, world, devices):
        // repeated one-shot searches with the same id reuse the communicator (an id can bootstrap only one)
        std::vector<int> ids;
        for (auto &d : devs) ids.push_back(d.device);
        Status s = NcclWorld::get(ids, std::max(0, opts.rank), world, opts.nccl_id, nccl);
        if (!s.ok()) return s;
    }
    return Status();
}

Status SparsePlan::Impl::init_device(DevCtx &c)
{
    SGS_CUDA_TRY(cudaSetDevice(c.device));
    SGS_CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, c.device));   // (cudaGetDeviceProperties costs milliseconds)
    SGS_CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    SGS_CUDA_TRY(cudaEventCreate(&c.ev0));
    SGS_CUDA_TRY(cudaEventCreate(&c.ev1));
    SGS_CUDA_TRY(cudaEventCreate(&c.evh0));
    SGS_CUDA_TRY(cudaEventCreate(&c.evh1));
    SGS_CUDA_TRY(cudaEventCreate(&c.evA));
    SGS_CUDA_TRY(cudaEventCreate(&c.evB));
    const int Tn = std::max(T, 1);
    c.nwbest = c.sms * 64;
    // one arena, carved into 256-byte aligned buffers
    struct Carve { void **pp; size_t bytes; };
    std::vector<Carve> carve;
    auto want = [&](auto **pp, size_t count) { carve.push_back({(void **)pp, ((count * sizeof(**pp)) + 255) & ~(size_t)255}); };
    want(&c.rows, (size_t)Tn * W);
    want(&c.adj, (size_t)Tn * TW);
    want(&c.wgt, Tn);
    want(&c.wabs, Tn);
    want(&c.order, Tn);
    want(&c.wbest, c.nwbest);
    want(&c.wtime, (size_t)c.nwbest * 2);
    want(&c.shardrec, 1);
    want(&c.allrec, std::max(nshards, 1));
    want(&c.qA, (size_t)qcap * S);
    want(&c.qB, (size_t)qcap * S);
    want(&c.slow, (size_t)slow_cap * S);
    want(&c.work, (size_t)chunk * S);
    want(&c.perm, qcap);
    want(&c.inc, 1);
    want(&c.retry[0], (size_t)retry_cap * S);
    want(&c.retry[1], (size_t)retry_cap * S);
    // result span: everything the host reads after a search, contiguous so that one copy fetches it
    want(&c.fin, 1);
    want(&c.term_signs, Tn);
    want(&c.stats, 1);
    want(&c.gbest, 1);
    want(&c.counters, 32);
    want(&c.memX, Tn);
    want(&c.memA, Tn);
    want(&c.memT, Tn);
    want(&c.memCount, 1);
    size_t total = 0;
    for (auto &cv : carve) total += cv.bytes;
    SGS_CUDA_TRY(sparse_arena().acquire(c.device, total, &c.arena));
    {
        size_t off = 0;
        for (auto &cv : carve) { *cv.pp = (char *)c.arena + off; off += cv.bytes; }
    }

    // term table in paulis_list order (cpp/state.cpp:88-96)
    std::vector<uint64_t> rows((size_t)Tn * W, 0);
    std::vector<double> wgt(Tn, 0.0), wabs(Tn, 0.0);
    for (int p = 0; p < T; p++) {
        const int f = H.order[p];
        for (int k = 0; k < W; k++) rows[(size_t)p * W + k] = H.terms[f].w[k];
        wgt[p] = H.weights[f];
        wabs[p] = std::fabs(H.weights[f]);
    }
    SGS_CUDA_TRY(cudaMemcpyAsync(c.rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice, c.stream));
    SGS_CUDA_TRY(cudaMemcpyAsync(c.wgt, wgt.data(), wgt.size() * 8, cudaMemcpyHostToDevice, c.stream));
    SGS_CUDA_TRY(cudaMemcpyAsync(c.wabs, wabs.data(), wabs.size() * 8, cudaMemcpyHostToDevice, c.stream));
    if (T > 0) SGS_CUDA_TRY(cudaMemcpyAsync(c.order, H.order.data(), (size_t)T * 4, cudaMemcpyHostToDevice, c.stream));
    SGS_CUDA_TRY(cudaStreamSynchronize(c.stream));
    if (T > 0) {
        const int total = T * TW;
        if (W == 1) k_adjacency<1><<<(total + 127) / 128, 128, 0, c.stream>>>(T, TW, c.rows, c.adj);
        else if (W == 2) k_adjacency<2><<<(total + 127) / 128, 128, 0, c.stream>>>(T, TW, c.rows, c.adj);
        else k_adjacency<4><<<(total + 127) / 128, 128, 0, c.stream>>>(T, TW, c.rows, c.adj);
        SGS_CUDA_TRY(cudaGetLastError());
        SGS_CUDA_TRY(cudaStreamSynchronize(c.stream));
    }
    return Status();
}

// The walk below the subtree roots.  With several shards the walk is short, so a root that cannot be certified
// at once (a commuting dependency below it) would pin one warp for most of the kernel: such roots hand their
// children to a second launch (counts stay on the device).
template <int W_>
Status SparsePlan::Impl::launch_dfs(DevCtx &c, const SparseDev &D, unsigned grid, int wpb, size_t pw, int cap, const uint16_t *items,
                                    unsigned count, int nsh, const unsigned *count_dev, const unsigned *perm, const unsigned *valid0,
                                    unsigned long long roots_per_shard)
{
    cudaStream_t st = c.stream;
    // The hand-off rounds pay only when the whole walk of a shard is about as short as its longest subtree (a few
    // 100 us on one warp), which then alone decides how long the launch lasts: config 4 at 8 shards, 70k roots per shard,
    // walk 0.55 -> 0.39 ms.  A longer walk hides that tail behind the dynamic fetch (roots come heaviest first) and
    // re-rooting the children is pure cost: n=30/T=400 at 8 shards (170k roots per shard) 1.46 ms without the rounds,
    // 1.57 with them; n=40/T=500 at 8 shards 6.6 vs 9.4 ms; config 4 at 2 shards 0.85 vs 1.05 ms.  The root count per
    // resident warp is the proxy for the length of the walk.
    const bool split = ((nshards > 1 && roots_per_shard < 24ull * grid * (unsigned)wpb) || getenv("SGS_SPLIT")) && !getenv("SGS_NOSPLIT");
    unsigned *rc = c.counters + 24;      // [0] round-2 count, [1] its valid end, [2] round-3 count, [3] its valid end, [4], [5] item counters
    if (split) {
        SGS_CUDA_TRY(cudaMemsetAsync(rc, 0, 6 * sizeof(unsigned), st));
        SGS_CUDA_TRY(cudaMemsetAsync(rc + 1, 0xFF, sizeof(unsigned), st));
        SGS_CUDA_TRY(cudaMemsetAsync(rc + 3, 0xFF, sizeof(unsigned), st));
    }
    const bool prune = opts.prune != 0;
#define SGS_DFS(SPL, PRN, ...) k_dfs<W_, SPL, PRN><<<grid, wpb * 32, pw * wpb, st>>>(__VA_ARGS__)
    if (split) {
        // one hand-off round: the second launch walks whatever it meets itself (a child of an uncertifiable root that
        // is uncertifiable again has ~25 alive terms: a third, nearly empty launch cost more than it saved — config 4 at 8
        // shards 0.395 -> 0.364 ms; SGS_ROUNDS=3 brings the third launch back for A/B runs)
        const bool two = !(getenv("SGS_ROUNDS") && atoi(getenv("SGS_ROUNDS")) == 3);
        if (prune) {
            SGS_DFS(true, true, D, items, count, c.counters + 1, c.shard, nsh, cap, (int)pw, count_dev, perm, c.retry[0], rc, retry_cap, valid0, 0);
            if (two) SGS_DFS(false, true, D, c.retry[0], retry_cap, rc + 4, 0, 1, cap, (int)pw, rc, nullptr, nullptr, nullptr, 0, rc + 1, 0);
            else {
                SGS_DFS(true, true, D, c.retry[0], retry_cap, rc + 4, 0, 1, cap, (int)pw, rc, nullptr, c.retry[1], rc + 2, retry_cap, rc + 1, 0);
                SGS_DFS(false, true, D, c.retry[1], retry_cap, rc + 5, 0, 1, cap, (int)pw, rc + 2, nullptr, nullptr, nullptr, 0, rc + 3, 0);
            }
        } else {
            SGS_DFS(true, false, D, items, count, c.counters + 1, c.shard, nsh, cap, (int)pw, count_dev, perm, c.retry[0], rc, retry_cap, valid0, 0);
            if (two) SGS_DFS(false, false, D, c.retry[0], retry_cap, rc + 4, 0, 1, cap, (int)pw, rc, nullptr, nullptr, nullptr, 0, rc + 1, 0);
            else {
                SGS_DFS(true, false, D, c.retry[0], retry_cap, rc + 4, 0, 1, cap, (int)pw, rc, nullptr, c.retry[1], rc + 2, retry_cap, rc + 1, 0);
                SGS_DFS(false, false, D, c.retry[1], retry_cap, rc + 5, 0, 1, cap, (int)pw, rc + 2, nullptr, nullptr, nullptr, 0, rc + 3, 0);
            }
        }
        c.launches += two ? 2 : 3;
    } else {
        if (prune) SGS_DFS(false, true, D, items, count, c.counters + 1, c.shard, nsh, cap, (int)pw, count_dev, perm, nullptr, rc, 0, valid0, 0);
        else SGS_DFS(false, false, D, items, count, c.counters + 1, c.shard, nsh, cap, (int)pw, count_dev, perm, nullptr, rc, 0, valid0, 0);
        c.launches++;
    }
#undef SGS_DFS
    SGS_CUDA_TRY(cudaGetLastError());
    return Status();
}

// The last level of the top-of-tree expansion, done by the walk kernel in expansion mode: every parent (a node of the
// level before the subtree roots) is evaluated and counted, its valid children are written to `out` as the subtree
// roots.  One alive list and one duplicate check per parent instead of a from-scratch evaluation per parent by
// k_general (121 -> ~30 us on config 4); with several shards a parent belongs to the shard of its own tuple, so its
// children land in the owner's queue and the walk needs no filter.  counters[4] = number of roots written,
// counters[5] = first slot that did not fit (0xffffffff: none), counters[6] = work-item cursor of this launch.
template <int W_>
Status SparsePlan::Impl::launch_expand(DevCtx &c, const SparseDev &D, unsigned grid, int wpb, size_t pw, int cap, const uint16_t *items,
                                       unsigned count, uint16_t *out)
{
    cudaStream_t st = c.stream;
    SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 4, 0, 3 * sizeof(unsigned), st));
    SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 5, 0xFF, sizeof(unsigned), st));
    const int mode = kDfsOwnSelf | kDfsExpand;
    if (opts.prune)
        k_dfs<W_, true, true><<<grid, wpb * 32, pw * wpb, st>>>(D, items, count, c.counters + 6, c.shard, nshards, cap, (int)pw, nullptr, nullptr, out,
                                                                c.counters + 4, qcap, nullptr, mode);
    else
        k_dfs<W_, true, false><<<grid, wpb * 32, pw * wpb, st>>>(D, items, count, c.counters + 6, c.shard, nshards, cap, (int)pw, nullptr, nullptr, out,
                                                                 c.counters + 4, qcap, nullptr, mode);
    c.launches++;
    SGS_CUDA_TRY(cudaGetLastError());
    return Status();
}

template <int W_>
Status SparsePlan::Impl::search_local(DevCtx &c)
{
    SGS_CUDA_TRY(cudaSetDevice(c.device));
    const SparseDev D = dev_view(c);
    cudaStream_t st = c.stream;
    c.launches = 0;
    SGS_CUDA_TRY(cudaMemsetAsync(c.wbest, 0, sizeof(BestRec) * c.nwbest, st));
    SGS_CUDA_TRY(cudaMemsetAsync(c.stats, 0, sizeof(SparseStats), st));
    SGS_CUDA_TRY(cudaMemsetAsync(c.counters, 0, 4 * sizeof(unsigned), st));
        if (opts.prune) SGS_CUDA_TRY(cudaMemcpyAsync(c.inc, &kNoIncumbent, sizeof(kNoIncumbent), cudaMemcpyHostToDevice, st));
    // root record: the empty group
    std::vector<uint16_t> root(S, 0);
    SGS_CUDA_TRY(cudaMemcpyAsync(c.qA, root.data(), S * 2, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaEventRecord(c.ev0, st));

    // ---- k_general launch geometry ----
    const size_t gen_pw = general_smem_bytes<W_>(std::max(T, 1));
    int gen_wpb = 4;
    while (gen_wpb > 1 && gen_pw * gen_wpb > 96 * 1024) gen_wpb >>= 1;
    if (gen_pw * gen_wpb > 200 * 1024) return Status::fail(SGS_ERR_LIMIT, "too many terms for the general kernel's shared-memory lists");
    SGS_CUDA_TRY(cudaFuncSetAttribute(k_general<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(gen_pw * gen_wpb)));
    int gen_occ = 1;
    SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gen_occ, k_general<W_>, gen_wpb * 32, gen_pw * gen_wpb));
    gen_occ = std::max(1, gen_occ);
    const unsigned gen_maxgrid = (unsigned)std::min<long long>((long long)c.sms * gen_occ, c.nwbest / gen_wpb);
    auto launch_general = [&](const uint16_t *in, unsigned nin, uint16_t *out, unsigned *out_count, unsigned out_cap,
                              int shard, int nsh, int own_only = 0) {
        unsigned grid = std::min<unsigned>((nin + gen_wpb - 1) / gen_wpb, gen_maxgrid);
        if (grid == 0) return;
        k_general<W_><<<grid, gen_wpb * 32, gen_pw * gen_wpb, st>>>(D, in, nin, out, out_count, out_cap, shard, nsh, (int)gen_pw, own_only,
                                                                    own_only ? c.counters + 3 : nullptr);
        c.launches++;
    };

    // ---- breadth-first expansion of the top of the tree into subtree roots ----
    // Expand until (a) there are enough roots to keep every resident warp of every shard busy under
    // dynamic fetch and (b) a root's list of alive terms is expected to fit the 64-entry level-B walk
    // (alive(depth d) ~ ratio_{d-1} * d / 2 for clique-like growth), as far as the queue capacity allows.
    const unsigned long long min_items = 8ull * c.sms * 24 * (unsigned long long)nshards;
    const int dmax = opts.expand_depth > 0 ? std::min(opts.expand_depth, 24) : 10;
    uint16_t *cur = c.qA, *nxt = c.qB;
    unsigned count = 1;
    double ratio = (double)std::max(T, 1);
    int depth = 0;
    SparseStats hs{};
    unsigned hcnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<unsigned> level_counts{1};
    c.had_slow = false;
    bool dfs_expand = false;
    double ratio_pred = 0.0;
    // sharded: the last level was expanded by the owner of each parent only (as the recorded launch sequence does): this
    // shard's queue then holds its own roots and nothing else, the walk needs no filter.  The level is chosen by
    // predicting that it will be the last one (clique-like growth); a prediction that turns out wrong in either
    // direction costs time, never the partition — a root belongs to the shard of its parent's tuple in every case.
    bool sharded = false;
    while (count > 0 && depth < dmax && !sharded) {
        if (opts.expand_depth <= 0) {
            const double est_alive = ratio * std::max(depth, 1) / 2.0;
            if (depth > 0 && count >= min_items && est_alive <= 64.0) break;
        }
        // Can the NEXT level be the last one?  Then the walk kernel produces it itself (launch_expand), incrementally.
        // Clique-like growth: count_{d+1} ~ count_d * ratio_d * d / (2 (d + 1)).
        // (opt-in, SGS_DFS_EXPAND=1: measured on config 4 it saves 26 us of expansion and costs 50 us of walk — the
        //  k_general levels stay the default; the mode is kept for instances whose last level is large, and tested)
        if (!no_dfs_expand && depth >= 1 && getenv("SGS_DFS_EXPAND")) {
            const double predicted = (double)count * ratio * depth / (2.0 * (depth + 1));
            const double ratio_next = predicted / (double)count;
            const bool last = opts.expand_depth > 0 ? depth == dmax - 1
                                                    : (predicted >= (double)min_items && ratio_next * (depth + 1) / 2.0 <= 64.0);
            if (last && predicted * 1.5 <= (double)qcap && ratio * depth / 2.0 <= 1500.0) { dfs_expand = true; ratio_pred = std::max(1.0, ratio_next); break; }
        }
        // worst case: the branching cannot grow (ratio is non-increasing for clique-like trees)
        const double bound = (double)count * std::min<double>((double)std::max(T, 1), ratio);
        if (bound > (double)qcap) {
            // likely case: clique-like growth, ratio_d ~ ratio_{d-1} * d / (2 (d+1)).  If that fits, measure the
            // exact size of the next level with a count-only pass before writing it.
            const double predicted = (double)count * ratio * std::max(depth, 1) / (2.0 * (depth + 1));
            if (opts.expand_depth > 0 || predicted * 1.5 > (double)qcap) break;
            SGS_CUDA_TRY(cudaMemsetAsync(c.counters, 0, sizeof(unsigned), st));
            launch_general(cur, count, nullptr, c.counters, qcap, c.shard, nshards);
            SGS_CUDA_TRY(cudaGetLastError());
            SGS_CUDA_TRY(cudaMemcpyAsync(hcnt, c.counters, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
            SGS_CUDA_TRY(cudaStreamSynchronize(st));
            if (hcnt[0] > qcap) break;
        }
        bool own = false;
        if (nshards > 1 && depth >= 1 && !getenv("SGS_NO_OWN_LEVEL")) {
            const double predicted = (double)count * ratio * depth / (2.0 * (depth + 1));
            const double ratio_next = predicted / (double)count;
            own = opts.expand_depth > 0 ? depth == dmax - 1 : (predicted >= (double)min_items && ratio_next * (depth + 1) / 2.0 <= 64.0);
        }
        SGS_CUDA_TRY(cudaMemsetAsync(c.counters, 0, sizeof(unsigned), st));
        if (own) SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 3, 0, sizeof(unsigned), st));
        launch_general(cur, count, nxt, c.counters, qcap, c.shard, nshards, own ? 1 : 0);
        SGS_CUDA_TRY(cudaGetLastError());
        SGS_CUDA_TRY(cudaMemcpyAsync(hcnt, c.counters, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaMemcpyAsync(&hs, c.stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs.error) return Status::fail(hs.error == kErrOverflow ? SGS_ERR_OVERFLOW : SGS_ERR_LIMIT,
                                          "sparse search: device limit hit while expanding the tree (queue overflow or rank/table limit)");
        sharded = own;
        ratio = std::max(1.0, (double)hcnt[0] * (own ? nshards : 1) / (double)count);
        if (D.trace) fprintf(stderr, "sparse shard %d: level %d -> %u nodes\n", c.shard, depth + 1, hcnt[0]);
        count = hcnt[0];
        level_counts.push_back(count);
        std::swap(cur, nxt);
        depth++;
    }

    // ---- the hot kernel: depth-first enumeration below every subtree root ----
    if (count > 0) {
        // list arena per warp: a root holds about `ratio` alive terms (the measured mean number of
        // children of the last expanded level) and the levels below it halve; 6x leaves room for spread
        // list arena per warp: a root holds a ~ ratio * depth / 2 alive terms (ratio = measured mean number of
        // children of the last expanded level; binomial spread), the lists of the levels below it halve:
        // 2.2 x (mean + 4 sigma), in steps of 32 entries.  A root that needs more goes to the exact kernel.
        // (after an expansion-mode launch the roots sit one level deeper: predicted ratio, depth + 1)
        const double ratio_w = dfs_expand ? ratio_pred : ratio;
        const int depth_w = dfs_expand ? depth + 1 : depth;
        const double a_est = std::max(8.0, ratio_w * std::max(depth_w, 1) / 2.0);
        int cap = (int)(2.2 * (a_est + 4.0 * std::sqrt(a_est)));
        bool cap_forced = false;
        if (const char *v = getenv("SGS_DFS_CAP")) if (atoi(v) > 0) { cap = atoi(v); cap_forced = true; }
        cap = std::min(2048, std::max(96, (cap + 31) & ~31));
        int dfs_wpb = kDfsWpb;
        while (dfs_wpb > 1 && dfs_smem_bytes<W_>(cap, opts.prune != 0) * dfs_wpb > 200 * 1024) dfs_wpb >>= 1;
        if (!cap_forced) {
            // the arena bounds the resident warps of a large instance (n = 40, T = 1000: 384 entries = 16 warps per SM):
            // give up to a fifth of the spread allowance (down to 1.9 x (mean + 3 sigma)) when that buys another block per SM
            const int cap_min = std::min(cap, std::max(96, ((int)(1.9 * (a_est + 3.0 * std::sqrt(a_est))) + 31) & ~31));
            const void *f = (const void *)k_dfs<W_, false, false>;
            SGS_CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(dfs_smem_bytes<W_>(cap, opts.prune != 0) * dfs_wpb)));
            int best_occ = 0, best_cap = cap;
            for (int cx = cap; cx >= cap_min; cx -= 32) {
                int o = 0;
                SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, f, dfs_wpb * 32, dfs_smem_bytes<W_>(cx, opts.prune != 0) * dfs_wpb));
                if (o > best_occ) { best_occ = o; best_cap = cx; }
            }
            cap = best_cap;
        }
        dfs_cap = cap;
        const size_t dfs_pw = dfs_smem_bytes<W_>(dfs_cap, opts.prune != 0);
        // expansion-mode launch: one list per parent (a ~ ratio * depth / 2 alive terms, 5 sigma of room)
        int exp_cap = 0, exp_wpb = kDfsWpb;
        size_t exp_pw = 0;
        if (dfs_expand) {
            const double a_par = std::max(8.0, ratio * depth / 2.0);
            exp_cap = std::min(2048, std::max(96, ((int)(1.25 * (a_par + 5.0 * std::sqrt(a_par))) + 31) & ~31));
            exp_pw = dfs_smem_bytes<W_>(exp_cap, opts.prune != 0);
            while (exp_wpb > 1 && exp_pw * exp_wpb > 200 * 1024) exp_wpb >>= 1;
        }
        int dfs_occ = 1 << 30, exp_occ = 1 << 30;
        {
            const void *variants[4] = {(const void *)k_dfs<W_, false, false>, (const void *)k_dfs<W_, true, false>,
                                       (const void *)k_dfs<W_, false, true>, (const void *)k_dfs<W_, true, true>};
            for (const void *f : variants) {
                int o = 1;
                SGS_CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(dfs_pw * dfs_wpb, exp_pw * exp_wpb)));
                SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, f, dfs_wpb * 32, dfs_pw * dfs_wpb));
                dfs_occ = std::min(dfs_occ, o);
                if (dfs_expand) {
                    SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, f, exp_wpb * 32, exp_pw * exp_wpb));
                    exp_occ = std::min(exp_occ, o);
                }
            }
        }
        dfs_occ = std::max(1, dfs_occ);
        const unsigned dfs_maxgrid = (unsigned)std::min<long long>((long long)c.sms * dfs_occ, c.nwbest / dfs_wpb);
        // (roots of an expansion-mode launch are counted on the device: a full grid)
        const unsigned grid = dfs_expand ? dfs_maxgrid : std::min<unsigned>(dfs_maxgrid, (count + dfs_wpb - 1) / dfs_wpb);
        unsigned exp_grid = 0;
        if (dfs_expand) {
            const unsigned mg = (unsigned)std::min<long long>((long long)c.sms * std::max(1, exp_occ), c.nwbest / exp_wpb);
            exp_grid = std::max(1u, std::min<unsigned>(mg, (count + exp_wpb - 1) / exp_wpb));
        }
        if (&c == &devs[0]) {
            rp.depth = depth; rp.counts = level_counts; rp.dfs_cap = dfs_cap; rp.dfs_wpb = dfs_wpb; rp.dfs_pw = dfs_pw;
            rp.dfs_grid = grid; rp.gen_wpb = gen_wpb; rp.gen_pw = gen_pw; rp.gen_maxgrid = gen_maxgrid;
            rp.dfs_expand = dfs_expand; rp.exp_cap = exp_cap; rp.exp_wpb = exp_wpb; rp.exp_pw = exp_pw; rp.exp_grid = exp_grid;
        }
        if (dfs_expand) {
            SGS_TRY(launch_expand<W_>(c, D, exp_grid, exp_wpb, exp_pw, exp_cap, cur, count, nxt));
            std::swap(cur, nxt);
        }
        SGS_CUDA_TRY(cudaEventRecord(c.evh0, st));
        // heavy-first order pays when the walk is short (several shards): the longest subtrees must not start last
        const bool use_perm = !getenv("SGS_NOPERM");
        const unsigned *cnt_dev = dfs_expand ? c.counters + 4 : nullptr;        // roots written by the expansion-mode launch
        const unsigned cnt_host = dfs_expand ? qcap : count;
        if (use_perm) {
            SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 8, 0, 16 * sizeof(unsigned), st));
            const unsigned cg = dfs_expand ? (unsigned)c.sms * 4 : std::min<unsigned>((count + 255) / 256, (unsigned)c.sms * 4);
            k_class_count<<<cg, 256, 0, st>>>(cur, cnt_host, cnt_dev, S, T, c.counters + 8);
            k_class_scatter<<<cg, 256, 0, st>>>(cur, cnt_host, cnt_dev, S, T, c.counters + 8, c.counters + 16, c.perm);
            c.launches += 2;
        }
        const unsigned long long rps = sharded ? (unsigned long long)count : (unsigned long long)count * (dfs_expand ? 8u : 1u) / (unsigned)nshards;
        if (&c == &devs[0]) rp.roots_per_shard = rps;
        Status sd = launch_dfs<W_>(c, D, grid, dfs_wpb, dfs_pw, dfs_cap, cur, cnt_host, (dfs_expand || sharded) ? 1 : nshards, cnt_dev,
                                   use_perm ? c.perm : nullptr, dfs_expand ? c.counters + 5 : nullptr, rps);
        if (!sd.ok()) return sd;
        SGS_CUDA_TRY(cudaGetLastError());
        SGS_CUDA_TRY(cudaEventRecord(c.evh1, st));
    } else {
        SGS_CUDA_TRY(cudaEventRecord(c.evh0, st));
        SGS_CUDA_TRY(cudaEventRecord(c.evh1, st));
    }

    // ---- drain the work stack of groups with dependent terms through the exact kernel ----
    for (;;) {
        SGS_CUDA_TRY(cudaMemcpyAsync(hcnt, c.counters, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaMemcpyAsync(&hs, c.stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs.error) return Status::fail(hs.error == kErrOverflow ? SGS_ERR_OVERFLOW : SGS_ERR_LIMIT,
                                          "sparse search: device limit hit (work-stack overflow, or a group whose sign table exceeds the limit)");
        if (dfs_expand && hcnt[4] > qcap) {
            // the predicted size of the last level was too low and its queue overflowed: redo the search with k_general levels only
            no_dfs_expand = true;
            return search_local<W_>(c);
        }
        const unsigned sc = hcnt[2];
        if (D.trace) fprintf(stderr, "sparse shard %d: work stack %u nodes%s\n", c.shard, sc, dfs_expand ? " (last level by k_dfs)" : "");
        if (sc == 0) break;
        c.had_slow = true;
        const unsigned take = std::min(sc, chunk);
        SGS_CUDA_TRY(cudaMemcpyAsync(c.work, c.slow + (size_t)(sc - take) * S, (size_t)take * S * 2, cudaMemcpyDeviceToDevice, st));
        const unsigned left = sc - take;
        SGS_CUDA_TRY(cudaMemcpyAsync(c.counters + 2, &left, sizeof(unsigned), cudaMemcpyHostToDevice, st));
        launch_general(c.work, take, c.slow, c.counters + 2, slow_cap, 0, 1);
        SGS_CUDA_TRY(cudaGetLastError());
    }

    k_reduce_pack<<<1, 1024, 0, st>>>(c.wbest, c.nwbest, c.stats, c.shardrec);
    c.launches++;
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaEventRecord(c.ev1, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(&c.hbest, &c.shardrec->best, sizeof(BestRec), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(&c.hstats, c.stats, sizeof(SparseStats), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    SGS_CUDA_TRY(cudaEventElapsedTime(&c.ms_device, c.ev0, c.ev1));
    SGS_CUDA_TRY(cudaEventElapsedTime(&c.ms_hot, c.evh0, c.evh1));
    return Status();
}

// combine the shard records (argmin + summed counters) and turn the winner into the printed result: one launch, one
// copy of the result span (FinalOut, term signs, statistics, best record)
template <int W_>
Status SparsePlan::Impl::finalize(DevCtx &c, const ShardRec *src, int nrec, sgs_result *out, SparseStats *tot, BestRec *best)
{
    SGS_CUDA_TRY(cudaSetDevice(c.device));
    const SparseDev D = dev_view(c);
    cudaStream_t st = c.stream;
    k_finalize<W_><<<1, 256, 0, st>>>(D, src, nrec, c.gbest, c.order, c.fin, c.term_signs, c.memX, c.memA, c.memT, c.memCount);
    c.launches++;
    SGS_CUDA_TRY(cudaGetLastError());
    const size_t span = (size_t)((const char *)(c.counters + 4) - (const char *)c.fin);
    std::vector<char> hostspan(span);
    SGS_CUDA_TRY(cudaMemcpyAsync(hostspan.data(), c.fin, span, cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    FinalOut fo;
    memcpy(&fo, hostspan.data(), sizeof(FinalOut));
    memcpy(tot, hostspan.data() + ((const char *)c.stats - (const char *)c.fin), sizeof(SparseStats));
    memcpy(best, hostspan.data() + ((const char *)c.gbest - (const char *)c.fin), sizeof(BestRec));
    if (!best->valid) return Status();                 // no candidate anywhere: the caller decides what that means
    if (fo.error) return Status::fail(SGS_ERR_LIMIT, "ground-state group rank exceeds the sign-table limit and its terms are linearly dependent");
    const int Wout = (2 * n + 63) / 64;
    out->energy = fo.energy;
    out->n = n; out->m = fo.m; out->W = Wout; out->T = T;
    out->gens = (uint64_t *)calloc((size_t)std::max(fo.m, 1) * Wout, 8);
    out->gen_signs = (int8_t *)calloc(std::max(fo.m, 1), 1);
    out->term_signs = (int8_t *)calloc(std::max(T, 1), 1);
    for (int i = 0; i < fo.m; i++) {
        for (int k = 0; k < Wout; k++) out->gens[(size_t)i * Wout + k] = fo.rows[i * W_ + k];
        out->gen_signs[i] = fo.signs[i];
    }
    memcpy(out->term_signs, hostspan.data() + ((const char *)c.term_signs - (const char *)c.fin), T);
    return Status();
}

void SparsePlan::Impl::fill_stats(const SparseStats &tot, sgs_result *out)
{
    out->candidates = 0;
    out->sum_2m = 0.0;
    for (int i = 0; i <= kMaxG && i <= SGS_MAX_RANK; i++) {
        out->rank_hist[i] = tot.hist[i];
        out->candidates += tot.hist[i];
        out->sum_2m += (double)tot.hist[i] * std::ldexp(1.0, i);
    }
    out->slow_nodes = tot.slow_nodes;
    for (int i = 0; i < 8; i++) out->walk_paths[i] = (uint32_t)std::min<unsigned long long>(tot.paths[i], 0xffffffffull);
}

// result of a shard (rank of a world without communicator) that owns no candidate: energy +inf, no generator
void SparsePlan::Impl::fill_empty(const SparseStats &tot, sgs_result *out)
{
    const int Wout = (2 * n + 63) / 64;
    out->energy = HUGE_VAL;
    out->n = n; out->m = 0; out->W = Wout; out->T = T;
    out->gens = (uint64_t *)calloc((size_t)Wout, 8);
    out->gen_signs = (int8_t *)calloc(1, 1);
    out->term_signs = (int8_t *)calloc(std::max(T, 1), 1);
    fill_stats(tot, out);
}

template <int W_>
Status SparsePlan::Impl::run_replay(sgs_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (auto &c : devs) {
        SGS_CUDA_TRY(cudaSetDevice(c.device));
        const SparseDev D = dev_view(c);
        cudaStream_t st = c.stream;
        c.launches = 0;
        SGS_CUDA_TRY(cudaMemsetAsync(c.wbest, 0, sizeof(BestRec) * c.nwbest, st));
        SGS_CUDA_TRY(cudaMemsetAsync(c.stats, 0, sizeof(SparseStats), st));
        SGS_CUDA_TRY(cudaMemsetAsync(c.counters, 0, 4 * sizeof(unsigned), st));
        if (opts.prune) SGS_CUDA_TRY(cudaMemcpyAsync(c.inc, &kNoIncumbent, sizeof(kNoIncumbent), cudaMemcpyHostToDevice, st));
        SGS_CUDA_TRY(cudaMemsetAsync(c.qA, 0, (size_t)S * 2, st));          // root record: the empty group
        SGS_CUDA_TRY(cudaEventRecord(c.ev0, st));
        uint16_t *cur = c.qA, *nxt = c.qB;
        for (int d = 0; d < rp.depth; d++) {
            const unsigned nin = rp.counts[d];
            const unsigned grid = std::min<unsigned>((nin + rp.gen_wpb - 1) / rp.gen_wpb, rp.gen_maxgrid);
            SGS_CUDA_TRY(cudaMemsetAsync(c.counters, 0, sizeof(unsigned), st));
            // the last level is expanded by the owner of each parent only: its children land in the owner's
            // queue and k_dfs walks all of them (no shard filter), so the expansion work shards too
            const int own_only = (nshards > 1 && d == rp.depth - 1 && !rp.dfs_expand) ? 1 : 0;
            if (own_only) SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 3, 0, sizeof(unsigned), st));
            k_general<W_><<<grid, rp.gen_wpb * 32, rp.gen_pw * rp.gen_wpb, st>>>(D, cur, nin, nxt, c.counters, qcap, c.shard, nshards,
                                                                                 (int)rp.gen_pw, own_only, own_only ? c.counters + 3 : nullptr);
            c.launches++;
            std::swap(cur, nxt);
        }
        // last level by the walk kernel in expansion mode (parents owned by the shard of their own tuple)
        if (rp.dfs_expand) {
            SGS_TRY(launch_expand<W_>(c, D, rp.exp_grid, rp.exp_wpb, rp.exp_pw, rp.exp_cap, cur, rp.counts[rp.depth], nxt));
            std::swap(cur, nxt);
        }
        const bool dev_count = rp.dfs_expand || (nshards > 1 && rp.depth > 0);      // queue length only known on the device
        const unsigned *ndev = rp.dfs_expand ? c.counters + 4 : (dev_count ? c.counters : nullptr);
        // (device-side count: the recorded number is only this plan's first shard's — the host bound is the queue capacity)
        const unsigned nroots = (rp.dfs_expand || dev_count) ? qcap : rp.counts[rp.depth];
        const bool use_perm = !getenv("SGS_NOPERM");
        if (use_perm) {
            const unsigned cnt = nroots;
            const unsigned cg = rp.dfs_expand ? (unsigned)c.sms * 4 : std::max(1u, std::min<unsigned>((cnt + 255) / 256, (unsigned)c.sms * 4));
            SGS_CUDA_TRY(cudaMemsetAsync(c.counters + 8, 0, 16 * sizeof(unsigned), st));
            k_class_count<<<cg, 256, 0, st>>>(cur, cnt, ndev, S, T, c.counters + 8);
            k_class_scatter<<<cg, 256, 0, st>>>(cur, cnt, ndev, S, T, c.counters + 8, c.counters + 16, c.perm);
            c.launches += 2;
        }
        SGS_CUDA_TRY(cudaEventRecord(c.evh0, st));
        // with a parent-sharded last level the queue length is only known on the device: counters[0]
        {
            Status sd = launch_dfs<W_>(c, D, rp.dfs_grid, rp.dfs_wpb, rp.dfs_pw, rp.dfs_cap, cur, nroots, dev_count ? 1 : nshards, ndev,
                                       use_perm ? c.perm : nullptr, rp.dfs_expand ? c.counters + 5 : nullptr, rp.roots_per_shard);
            if (!sd.ok()) return sd;
        }
        SGS_CUDA_TRY(cudaEventRecord(c.evh1, st));
        k_reduce_pack<<<1, 1024, 0, st>>>(c.wbest, c.nwbest, c.stats, c.shardrec);
        c.launches += 1;
        SGS_CUDA_TRY(cudaGetLastError());
        SGS_CUDA_TRY(cudaEventRecord(c.evA, st));
    }
    if (nccl) {
        // the one collective: every shard's (best record, counters)
        std::vector<void *> sendb, recvb;
        std::vector<cudaStream_t> sts;
        for (auto &c : devs) { sendb.push_back(c.shardrec); recvb.push_back(c.allrec); sts.push_back(c.stream); }
        Status s = nccl->all_gather_bytes(sendb, recvb, sizeof(ShardRec), sts);
        if (!s.ok()) return s;
    }
    DevCtx &c0 = devs[0];
    SGS_CUDA_TRY(cudaSetDevice(c0.device));
    SGS_CUDA_TRY(cudaEventRecord(c0.evB, c0.stream));
    {
        const SparseDev D = dev_view(c0);
        k_finalize<W_><<<1, 256, 0, c0.stream>>>(D, nccl ? c0.allrec : c0.shardrec, nccl ? nshards : 1, c0.gbest, c0.order, c0.fin, c0.term_signs,
                                                 c0.memX, c0.memA, c0.memT, c0.memCount);
        c0.launches++;
        SGS_CUDA_TRY(cudaGetLastError());
    }
    // one device-to-host copy of the result span (FinalOut, term signs, statistics, best record, counters)
    const size_t span = (size_t)((const char *)(c0.counters + 4) - (const char *)c0.fin);
    std::vector<char> hostspan(span);
    SGS_CUDA_TRY(cudaMemcpyAsync(hostspan.data(), c0.fin, span, cudaMemcpyDeviceToHost, c0.stream));
    for (auto &c : devs) { SGS_CUDA_TRY(cudaSetDevice(c.device)); SGS_CUDA_TRY(cudaEventRecord(c.ev1, c.stream)); }
    float msd = 0.f, msh = 0.f;
    unsigned launches = 0;
    for (auto &c : devs) {
        SGS_CUDA_TRY(cudaSetDevice(c.device));
        SGS_CUDA_TRY(cudaStreamSynchronize(c.stream));
        float a = 0.f, b = 0.f;
        SGS_CUDA_TRY(cudaEventElapsedTime(&a, c.ev0, c.ev1));
        SGS_CUDA_TRY(cudaEventElapsedTime(&b, c.evh0, c.evh1));
        msd = std::max(msd, a); msh = std::max(msh, b); launches += c.launches;
    }
    FinalOut fo;
    std::vector<int8_t> ts(std::max(T, 1));
    unsigned hcnt[4] = {0, 0, 0, 0};
    memcpy(&fo, hostspan.data(), sizeof(FinalOut));
    memcpy(ts.data(), hostspan.data() + ((const char *)c0.term_signs - (const char *)c0.fin), std::max(T, 1));
    memcpy(&c0.hstats, hostspan.data() + ((const char *)c0.stats - (const char *)c0.fin), sizeof(SparseStats));
    memcpy(&c0.hbest, hostspan.data() + ((const char *)c0.gbest - (const char *)c0.fin), sizeof(BestRec));
    memcpy(hcnt, hostspan.data() + ((const char *)c0.counters - (const char *)c0.fin), sizeof(hcnt));
    if (getenv("SGS_TRACE")) {
        float e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0;
        cudaEventElapsedTime(&e1, c0.ev0, c0.evh0); cudaEventElapsedTime(&e2, c0.evh0, c0.evh1); cudaEventElapsedTime(&e3, c0.evh1, c0.evA);
        cudaEventElapsedTime(&e4, c0.evA, c0.evB); cudaEventElapsedTime(&e5, c0.evB, c0.ev1);
        fprintf(stderr, "replay shard %d: expansion %.3f ms, walk %.3f ms, reduce %.3f ms, collective %.3f ms, finalise+copy %.3f ms, host total %.3f ms\n",
                c0.shard, e1, e2, e3, e4, e5, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3);
        const unsigned long long *g = c0.hstats.dbg;
        fprintf(stderr, "  walk: longest root %.1f us; roots not certified at the root %llu (%.3f warp-ms of %.3f); roots > 50 us: %llu (%.3f warp-ms)\n",
                g[0] * 1e-3, g[1], g[2] * 1e-6, g[3] * 1e-6, g[5], g[6] * 1e-6);
        fprintf(stderr, "  warp lifetimes %.3f warp-ms in total (%u warps x walk time = %.3f)\n", g[4] * 1e-6, rp.dfs_grid * rp.dfs_wpb, rp.dfs_grid * rp.dfs_wpb * e2);
        if (g[4]) {
            const unsigned nw = rp.dfs_grid * rp.dfs_wpb;
            std::vector<unsigned long long> wt(2 * (size_t)nw);
            cudaMemcpy(wt.data(), c0.wtime, wt.size() * 8, cudaMemcpyDeviceToHost);
            unsigned long long t0w = ~0ull, t1w = 0;
            for (unsigned i = 0; i < nw; i++) { t0w = std::min(t0w, wt[2 * i]); t1w = std::max(t1w, wt[2 * i + 1]); }
            std::vector<double> ends(nw);
            for (unsigned i = 0; i < nw; i++) ends[i] = (double)(wt[2 * i + 1] - t0w) * 1e-3;
            std::sort(ends.begin(), ends.end());
            fprintf(stderr, "  warp end times (us after the first start): 1%% %.0f, 10%% %.0f, 25%% %.0f, 50%% %.0f, 75%% %.0f, 90%% %.0f, 99%% %.0f, last %.0f\n",
                    ends[nw / 100], ends[nw / 10], ends[nw / 4], ends[nw / 2], ends[3 * (size_t)nw / 4], ends[9 * (size_t)nw / 10], ends[99 * (size_t)nw / 100], ends[nw - 1]);
        }
        fprintf(stderr, "  longest root: round %llu, rank %llu, alive %llu, candidates %llu, certified %d\n", (g[7] >> 21) & 7, (g[7] >> 15) & 63,
                (g[7] >> 8) & 127, (g[7] >> 1) & 127, (int)!(g[7] & 1));
    }
    if (!c0.hstats.error && hcnt[2] == 0 && !c0.hbest.valid && nshards > 1 && !nccl) {
        fill_empty(c0.hstats, out);
        out->seconds_device = msd * 1e-3; out->seconds_hot_kernel = msh * 1e-3; out->kernel_launches = launches;
        out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return Status();
    }
    if (c0.hstats.error || hcnt[2] != 0 || !c0.hbest.valid || fo.error) {
        rp.ok = false;      // should not happen (the sequence is deterministic); fall back to the adaptive path
        return Status::fail(SGS_ERR_CUDA, "sparse search: replayed launch sequence diverged from the recorded run");
    }
    const int Wout = (2 * n + 63) / 64;
    out->energy = fo.energy;
    out->n = n; out->m = fo.m; out->W = Wout; out->T = T;
    out->gens = (uint64_t *)calloc((size_t)std::max(fo.m, 1) * Wout, 8);
    out->gen_signs = (int8_t *)calloc(std::max(fo.m, 1), 1);
    out->term_signs = (int8_t *)calloc(std::max(T, 1), 1);
    for (int i = 0; i < fo.m; i++) {
        for (int k = 0; k < Wout; k++) out->gens[(size_t)i * Wout + k] = fo.rows[i * W_ + k];
        out->gen_signs[i] = fo.signs[i];
    }
    memcpy(out->term_signs, ts.data(), T);
    fill_stats(c0.hstats, out);
    out->seconds_device = msd * 1e-3;
    out->seconds_hot_kernel = msh * 1e-3;
    out->kernel_launches = launches;
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return Status();
}

Status SparsePlan::Impl::run(sgs_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    memset(out, 0, sizeof(*out));
    if (rp.ok) {
        Status s = W == 1 ? run_replay<1>(out) : (W == 2 ? run_replay<2>(out) : run_replay<4>(out));
        // A diverged replay is redone adaptively only when this process is alone: with a communicator the other
        // ranks have left their collectives behind, so a second pair here would hang or mismatch — report it.
        if (s.ok() || rp.ok || nccl) return s;
        sgs_result_free(out);
        memset(out, 0, sizeof(*out));
    }
    auto local = [&](DevCtx &c) {
        if (W == 1) c.status = search_local<1>(c);
        else if (W == 2) c.status = search_local<2>(c);
        else c.status = search_local<4>(c);
    };
    if (devs.size() == 1) local(devs[0]);
    else {
        std::vector<std::thread> th;
        for (auto &c : devs) th.emplace_back(local, std::ref(c));
        for (auto &t : th) t.join();
    }
    for (auto &c : devs) if (!c.status.ok()) return c.status;

    // ---- combine the shards: ONE collective, an all-gather of every shard's (best record, counters); then the same
    // deterministic argmin, the summed counters and the finalisation in one kernel on every rank ----
    // (timed with CUDA events on the launching stream and added to the device time: the reported
    //  seconds_device covers expansion + walk + collective + finalisation)
    DevCtx &c0 = devs[0];
    cudaSetDevice(c0.device);
    cudaEventRecord(c0.ev0, c0.stream);
    if (nccl) {
        std::vector<void *> sendb, recvb;
        std::vector<cudaStream_t> sts;
        for (auto &c : devs) { sendb.push_back(c.shardrec); recvb.push_back(c.allrec); sts.push_back(c.stream); }
        Status sg = nccl->all_gather_bytes(sendb, recvb, sizeof(ShardRec), sts);
        if (!sg.ok()) return sg;
    }
    BestRec best{};
    SparseStats tot{};
    Status s;
    const ShardRec *src = nccl ? c0.allrec : c0.shardrec;
    const int nrec = nccl ? nshards : 1;
    if (W == 1) s = finalize<1>(c0, src, nrec, out, &tot, &best);
    else if (W == 2) s = finalize<2>(c0, src, nrec, out, &tot, &best);
    else s = finalize<4>(c0, src, nrec, out, &tot, &best);
    if (!s.ok()) return s;
    for (auto &c : devs) { cudaSetDevice(c.device); SGS_CUDA_TRY(cudaStreamSynchronize(c.stream)); }
    if (!best.valid) {
        if (nshards > 1 && !nccl) {          // a shard without communicator that owns no candidate: empty result
            fill_empty(tot, out);
            out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            return Status();
        }
        return Status::fail(SGS_ERR_CUDA, "sparse search produced no candidate");
    }
    float ms_merge = 0.f;
    cudaSetDevice(c0.device);
    cudaEventRecord(c0.ev1, c0.stream);
    cudaEventSynchronize(c0.ev1);
    cudaEventElapsedTime(&ms_merge, c0.ev0, c0.ev1);

    fill_stats(tot, out);
    // replay is valid when no device needed the exact kernel below the roots and there were roots at all
    bool can = !rp.counts.empty() && rp.dfs_grid > 0 && opts.expand_depth >= 0;
    for (auto &c : devs) can = can && !c.had_slow;
    rp.ok = can;
    float msd = 0.f, msh = 0.f;
    unsigned launches = 0;
    for (auto &c : devs) { msd = std::max(msd, c.ms_device); msh = std::max(msh, c.ms_hot); launches += c.launches; }
    out->seconds_device = (msd + ms_merge) * 1e-3;
    out->seconds_hot_kernel = msh * 1e-3;
    out->kernel_launches = launches;
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return Status();
}

SparsePlan::SparsePlan() : impl_(new Impl()) {}
SparsePlan::~SparsePlan() { delete impl_; }

Status SparsePlan::create(const Hamiltonian &H, const sgs_opts &opts, std::unique_ptr<SparsePlan> &out)
{
    std::unique_ptr<SparsePlan> p(new SparsePlan());
    p->impl_->H = H;
    p->impl_->opts = opts;
    Status s = p->impl_->init();
    if (!s.ok()) return s;
    out = std::move(p);
    return Status();
}

Status SparsePlan::run(sgs_result *out) { return impl_->run(out); }

}  // namespace sgs
