// klocal_search.cu — host driver of the 1D k-local dynamic programme and of the periodic chain search.
// Mirrors StateMachine (cpp/state.hpp:100-121, cpp/state.cpp:445-550), generate_Sright_all (cpp/state.cpp:218-238),
// and the Python-only StateMachinePeriodic / generate_Sright_periodic / klocal_periodic.py driver
// (py/state.py:402-462,542-567; py/klocal_periodic.py:1-63).  The host sequences sites and moves small
// records; every group operation, table update and merge runs in the kernels of klocal_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

#include "arena.hpp"
#include "common.hpp"
#include "group_ops.hpp"
#include "klocal_kernels.cuh"
#include "klocal_search.hpp"
#include "nccl_glue.hpp"

namespace sgs {

namespace {

// Bump allocator over one cached device arena (arena.hpp); when the arena is exhausted further chunks are
// cudaMalloc'ed and freed with the machine.  Nothing is freed individually: a machine lives for one search.
struct DevPool {
    int device = 0;
    void *arena = nullptr;
    std::vector<void *> extra;
    char *cur = nullptr;
    size_t left = 0;
    ~DevPool()
    {
        if (arena) klocal_arena().release(device, arena);
        for (void *c : extra) cudaFree(c);
    }
    cudaError_t open(int dev, size_t bytes)
    {
        device = dev;
        size_t got = 0;
        cudaError_t e = klocal_arena().acquire(dev, bytes, &arena, &got);
        if (e != cudaSuccess) return e;
        cur = (char *)arena; left = got;
        return cudaSuccess;
    }
    cudaError_t take(void **out, size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        if (bytes > left) {
            const size_t sz = std::max<size_t>(bytes, (size_t)64 << 20);
            void *c = nullptr;
            cudaError_t e = cudaMalloc(&c, sz);
            if (e != cudaSuccess) return e;
            extra.push_back(c);
            cur = (char *)c; left = sz;
        }
        *out = cur; cur += bytes; left -= bytes;
        return cudaSuccess;
    }
};

template <typename T>
struct DBuf {                       // growable device buffer carved from the pool (an outgrown region is abandoned)
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t need(DevPool &pool, size_t n)
    {
        if (n <= cap) return cudaSuccess;
        const size_t want = std::max<size_t>(n, cap * 2 + 64);
        void *q = nullptr;
        cudaError_t e = pool.take(&q, want * sizeof(T));
        if (e != cudaSuccess) return e;
        p = (T *)q; cap = want;
        return cudaSuccess;
    }
};

bool group_equal(const DpGroup &a, const DpGroup &b)
{
    if (a.r != b.r) return false;
    for (int i = 0; i < a.r; i++) if (a.g[i] != b.g[i] || a.neg[i] != b.neg[i]) return false;
    return true;
}

unsigned pow2_at_least(size_t v)
{
    unsigned p = 1024;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

struct HostState { DpState s; std::vector<double> E; };

struct KlocalMachine::Impl {
    Hamiltonian H;
    sgs_opts opts{};
    int device = 0, nsm = 0;
    int n = 0, k = 0, T = 0, W = 1;
    int scale = 1;                 // capacity multiplier (x4 per retry after an overflow)
    int max_m = 0;                 // highest level index (n, or l(1+cmax)+k for the periodic driver)
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool seg_open = false;
    DevPool pool;
    DpTerms terms{};
    std::vector<int> boff;
    // right environments
    SrBufs sr{};
    SrCtl hsr{};
    DpLevelDev *d_levels = nullptr;
    std::vector<DpLevelDev> hlev;                  // host mirror of the level directory
    std::map<int, std::vector<DpGroup>> hgroups;   // host copies of level group lists (periodic driver)
    // DP
    DpBufs dp{};
    DpCtl hctl{};
    bool hctl_valid = false;
    int m_ = 0, tstride = 1, slab_stride = 1;
    int caps = 0, capc = 0, hm_cap = 0, hb_cap = 0, log_cap = 1 << 14;
    int *d_trace = nullptr;
    char *d_finish = nullptr;      // workspace of group_finish
    int trace_cap = 0;
    unsigned launches = 0;
    double dev_seconds = 0.0;
    int cl_size = 0;               // CTAs of the cluster-mode kernels (0: no cluster mode; SGS_DP_CLUSTER)
    bool blk_ok = true;            // single-CTA mode available (SGS_DP_BLOCK=0 turns it off)
    unsigned small_dp = 0, small_sr = 0;           // bit MODE: still worth trying that small mode first (cleared once it bailed)
    Status launch_persistent(const void *const fns[3], unsigned small, void **args, int *resume);
    Status pending;                // error met inside a const accessor, returned by the next call

    ~Impl();
    Status init();
    Status alloc_sr(int cap, size_t pool_bytes);
    Status alloc_dp();
    Status set_level(int m, const DpLevelDev &L);
    Status make_top_level(int m);
    Status run_sright(int m_hi, int m_lo);
    Status host_groups(int m, const std::vector<DpGroup> **out);
    Status gen_sright_all();
    Status write_ctl(int nin, bool reset_history);
    Status sync_ctl();
    Status init_states();
    Status advance(int nsteps);
    Status set_states(const std::vector<HostState> &hs, int m, bool reset_history);
    Status get_states(std::vector<HostState> &hs);
    Status ground_state(sgs_result *out);
    Status traceback(int state, int entry, std::vector<std::pair<int, int>> &out);   // (file index, sign) in history order
    Status fetch_log(std::vector<DpStepLog> &log);
    // periodic chain / stepwise access (py/klocal_periodic.py, py/state.py:309-325,402-462,542-567)
    int epoch = 0;                 // bumped whenever the device states or their history change (set_states, advance)
    std::vector<int> sr_log;       // (m, count) pairs in the order generate_Sright_* prints them
    Status gen_sright_periodic(int cmax);
    Status shift_states(std::vector<HostState> &hs, int from_m, int add);
    std::vector<uint64_t> key_of(const HostState &h, int m) const;
    void dedupe_sort(std::vector<HostState> &hs, int m, std::vector<int> *origin = nullptr) const;
    template <typename T> Status take(T **p, size_t count) { void *q = nullptr; SGS_CUDA_TRY(pool.take(&q, std::max<size_t>(count, 1) * sizeof(T))); *p = (T *)q; return Status(); }
};

KlocalMachine::Impl::~Impl()
{
    cudaSetDevice(device);
    if (st) cudaStreamSynchronize(st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
}

static Status ctl_error(int err, const char *what)
{
    const char *why = err == 1 ? "window group rank exceeds the table limit (k too large)"
                      : err == 2 ? "more than 256 terms in k+2 consecutive last-site buckets"
                      : err == 3 ? "more than 16 terms share one last site (2^t right-environment branches)"
                      : err == 4 ? "right-environment branch capacity exceeded"
                                 : "batch capacity exceeded";
    return Status::fail(err >= 4 ? SGS_ERR_OVERFLOW : SGS_ERR_LIMIT, std::string(what) + ": " + why);
}


Status KlocalMachine::Impl::init()
{
    n = H.n; k = H.k; T = H.T();
    if (k < 1 || k > kDpMaxK) return Status::fail(SGS_ERR_LIMIT, "k-local DP supports 1 <= k <= 15");
    for (int t = 0; t < T; t++)
        if (H.end[t] - H.begin[t] > k) return Status::fail(SGS_ERR_ARG, "a term spans more than k sites (Hamiltonian::check_locality)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return Status::fail(SGS_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                                   " (libsgs has no CPU search path)");
    device = opts.device;
    if (device < 0 || device >= ndev) return Status::fail(SGS_ERR_ARG, "device ordinal out of range");
    SGS_CUDA_TRY(cudaSetDevice(device));
    SGS_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    int coop = 0, occ_dp = 0, occ_sr = 0;
    SGS_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dp, k_dp_run<kGrid>, 256, 0));
    SGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sr, k_sr_run<kGrid>, 256, 0));
    if (!coop || occ_dp < 1 || occ_sr < 1) return Status::fail(SGS_ERR_CUDA, "device cannot co-schedule the persistent DP kernels");
    if (const char *v = getenv("SGS_DP_GRID")) if (atoi(v) > 0) nsm = std::min(nsm, atoi(v));     // persistent grid size (experiments)
    // cluster mode for small batches: one cluster of 16 CTAs (non-portable size; 8 if the device refuses;
    // SGS_DP_CLUSTER=8|16 picks one, 0 turns cluster mode off)
    int want = 16;
    if (const char *v = getenv("SGS_DP_CLUSTER")) want = atoi(v);
    cl_size = 0;
    for (int size : {16, 8}) {
        if (size > want || (want != 16 && want != 8)) continue;
        bool ok = true;
        const void *fns[2] = {(const void *)k_dp_run<kCluster>, (const void *)k_sr_run<kCluster>};
        for (const void *fn : fns) {
            if (size > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { ok = false; break; }
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(size); cfg.blockDim = dim3(256); cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg) != cudaSuccess || ncl < 1) { ok = false; break; }
        }
        (void)cudaGetLastError();
        if (ok) { cl_size = size; break; }
    }
    if (const char *v = getenv("SGS_DP_BLOCK")) blk_ok = atoi(v) != 0;
    small_dp = small_sr = (cl_size > 0 ? 1u << kCluster : 0u) | (blk_ok ? 1u << kBlock : 0u);
    SGS_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    SGS_CUDA_TRY(cudaEventCreate(&e0));
    SGS_CUDA_TRY(cudaEventCreate(&e1));
    W = (2 * (n + 1) + 63) / 64;
    tstride = 1 << std::min(k, kDpTab);
    slab_stride = tstride;
    max_m = std::max(max_m, n);
    // SGS_DP_CAPS / SGS_SR_CAP / SGS_SR_POOL_KB: start capacities for tests of the grow-and-redo paths
    auto env_ll = [](const char *name, long long dflt) { const char *v = getenv(name); return v && atoll(v) > 0 ? atoll(v) : dflt; };
    const long long base = env_ll("SGS_DP_CAPS", std::max<long long>(2048, (1ll << 22) / ((long long)slab_stride * 16)));   // k=4: 16384 states
    caps = (int)std::min<long long>(base * scale, 1ll << 22);
    capc = (int)std::min<long long>((long long)caps * 8, 1ll << 25);
    hm_cap = (int)std::min<long long>((long long)caps * 64, 1ll << 27);
    hb_cap = (int)std::min<long long>((long long)caps * 64 * std::min(tstride, 64), 1ll << 29);
    const int srcap = (int)std::min<long long>(env_ll("SGS_SR_CAP", 65536) * scale, 1ll << 24);
    const size_t srpool = (size_t)env_ll("SGS_SR_POOL_KB", 64 << 10) * 1024 * (size_t)std::min(scale, 64);
    {   // one cached arena for the tables, the DP batch buffers, the right-environment scratch and the level tables
        const size_t dpb = (size_t)caps * (2 * sizeof(DpState) + sizeof(DpMoved) + sizeof(Ech) + 16) + (size_t)caps * tstride * 16 +
                           (size_t)caps * 2 * slab_stride * 16 + (size_t)capc * (sizeof(DpKey) + 48 + 2 * sizeof(DdSlot)) +
                           (size_t)hm_cap * sizeof(DpHistAdd) + (size_t)hb_cap * sizeof(DpBack);
        const size_t srb = (size_t)srcap * (2 * sizeof(SrBranch) + 2 * sizeof(DpGroup) + 64 + 2 * sizeof(DdSlot)) + srpool;
        SGS_CUDA_TRY(pool.open(device, dpb + srb + ((size_t)16 << 20)));
    }
    const int Tn = std::max(T, 1);
    std::vector<uint64_t> rows((size_t)Tn, 0);
    std::vector<double> w(Tn, 0.0);
    std::vector<int> fidx(Tn, 0), tbeg(Tn, 0);
    for (int p = 0; p < T; p++) {
        const int f = H.order[p];
        if (!H.compact[f]) return Status::fail(SGS_ERR_LIMIT, "k-local DP: a term spans more than 32 sites");
        rows[p] = H.crow[f];
        tbeg[p] = H.begin[f];
        w[p] = H.weights[f];
        fidx[p] = f;
    }
    boff = H.bucket_off;            // n + 1 entries, boff[n] = T
    boff.push_back(T);              // boff[n + 1]
    uint64_t *d_rows; double *d_w; int *d_boff, *d_fidx, *d_tb;
    SGS_TRY(take(&d_rows, rows.size())); SGS_TRY(take(&d_w, w.size())); SGS_TRY(take(&d_boff, boff.size())); SGS_TRY(take(&d_fidx, fidx.size()));
    SGS_TRY(take(&d_tb, tbeg.size()));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_tb, tbeg.data(), tbeg.size() * 4, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_w, w.data(), w.size() * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_boff, boff.data(), boff.size() * 4, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_fidx, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));                // the host vectors go out of scope
    terms.n = n; terms.k = k; terms.T = T; terms.W = W;
    terms.rows = d_rows; terms.tb = d_tb; terms.w = d_w; terms.boff = d_boff; terms.file_index = d_fidx;
    hlev.assign(max_m + 2, DpLevelDev{});
    SGS_TRY(take(&d_levels, hlev.size()));
    SGS_CUDA_TRY(cudaMemsetAsync(d_levels, 0, hlev.size() * sizeof(DpLevelDev), st));
    SGS_TRY(alloc_sr(srcap, srpool));
    SGS_TRY(alloc_dp());
    return Status();
}

static Status alloc_view(KlocalMachine::Impl &K, DdView &D, size_t entries, cudaStream_t st)
{
    SGS_TRY(K.take(&D.pre, entries)); SGS_TRY(K.take(&D.owner, entries)); SGS_TRY(K.take(&D.next, entries));
    SGS_TRY(K.take(&D.leader, entries)); SGS_TRY(K.take(&D.rank, entries)); SGS_TRY(K.take(&D.lidx, entries));
    SGS_TRY(K.take(&D.lpre, entries));
    const unsigned tsz = pow2_at_least(2 * entries);
    SGS_TRY(K.take(&D.tab, tsz));
    D.mask = tsz - 1;
    SGS_CUDA_TRY(cudaMemsetAsync(D.tab, 0xFF, (size_t)tsz * sizeof(DdSlot), st));   // every word carries the tag of generation 0 = empty
    return Status();
}

Status KlocalMachine::Impl::alloc_sr(int cap, size_t pool_bytes)
{
    sr.cap = cap;
    sr.waves_only = getenv("SGS_SR_WAVES") ? 1 : 0;        // breadth-first waves only (A/B, tests)
    SGS_TRY(take(&sr.br[0], cap)); SGS_TRY(take(&sr.br[1], cap)); SGS_TRY(take(&sr.flat, cap));
    SGS_TRY(take(&sr.par, cap)); SGS_TRY(take(&sr.cnt, (size_t)cap + 1)); SGS_TRY(take(&sr.moff, (size_t)cap + 1)); SGS_TRY(take(&sr.first, cap));
    SGS_TRY(alloc_view(*this, sr.D, cap, st));
    SGS_TRY(take(&sr.ctl, 1));
    SGS_TRY(take(&sr.pool, pool_bytes));
    sr.levels = d_levels;
    hsr = SrCtl{};
    hsr.gen = 1;
    hsr.pool_cap = pool_bytes;
    SGS_CUDA_TRY(cudaMemcpyAsync(sr.ctl, &hsr, sizeof(SrCtl), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    return Status();
}

Status KlocalMachine::Impl::alloc_dp()
{
    for (int i = 0; i < 2; i++) { SGS_TRY(take(&dp.S[i], caps)); SGS_TRY(take(&dp.E[i], (size_t)caps * tstride)); }
    SGS_TRY(take(&dp.slabE, (size_t)caps * 2 * slab_stride)); SGS_TRY(take(&dp.slabB, (size_t)caps * 2 * slab_stride));
    SGS_TRY(take(&dp.moved, caps)); SGS_TRY(take(&dp.span, caps)); SGS_TRY(take(&dp.nacc, (size_t)caps + 1)); SGS_TRY(take(&dp.off, (size_t)caps + 1));
    SGS_TRY(take(&dp.keys, capc)); SGS_TRY(take(&dp.parent, capc));
    SGS_TRY(alloc_view(*this, dp.D, capc, st));
    SGS_TRY(take(&dp.hm, hm_cap)); SGS_TRY(take(&dp.hb, hb_cap));
    SGS_TRY(take(&dp.ctl, 1)); SGS_TRY(take(&dp.log, log_cap));
    trace_cap = T + 16;
    SGS_TRY(take(&d_trace, 2 * (size_t)trace_cap + 8));
    SGS_TRY(take(&d_finish, group_finish_ws_bytes(T)));
    dp.levels = d_levels;
    dp.tstride = tstride; dp.slab_stride = slab_stride;
    hctl = DpCtl{};
    hctl.gen = 1;
    hctl.caps = caps; hctl.capc = capc; hctl.hm_cap = hm_cap; hctl.hb_cap = hb_cap; hctl.log_cap = log_cap;
    SGS_CUDA_TRY(cudaMemcpyAsync(dp.ctl, &hctl, sizeof(DpCtl), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    hctl_valid = true;
    return Status();
}

Status KlocalMachine::Impl::set_level(int m, const DpLevelDev &L)
{
    hlev[m] = L;
    hgroups.erase(m);
    SGS_CUDA_TRY(cudaMemcpyAsync(d_levels + m, &hlev[m], sizeof(DpLevelDev), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    return Status();
}

// Stab::empty(H.k, H.n - H.k + 1) as the only group of level m (cpp/state.cpp:219)
Status KlocalMachine::Impl::make_top_level(int m)
{
    DpGroup empty{};
    const int zero2[2] = {0, 0};
    DpGroup *g; int *mo, *mi;
    SGS_TRY(take(&g, 1)); SGS_TRY(take(&mo, 2)); SGS_TRY(take(&mi, 1));
    SGS_CUDA_TRY(cudaMemcpyAsync(g, &empty, sizeof(DpGroup), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(mo, zero2, sizeof(zero2), cudaMemcpyHostToDevice, st));
    DpLevelDev L{g, mo, mi, 1, 0};
    return set_level(m, L);
}

// One run of a persistent kernel chain: the small-mode kernels that are still worth trying go first (one CTA,
// then one cluster), the grid-mode kernel behind them resumes wherever they stopped (nothing to do if they
// finished: an empty cooperative launch).  fns[] is indexed by mode; args[last] is the resume flag.
Status KlocalMachine::Impl::launch_persistent(const void *const fns[3], unsigned small, void **args, int *resume)
{
    *resume = 0;
    if (small & (1u << kBlock)) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.stream = st;
        SGS_CUDA_TRY(cudaLaunchKernelExC(&cfg, fns[kBlock], args));
        launches++;
        *resume = 1;
    }
    if (small & (1u << kCluster)) {
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cl_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(cl_size); cfg.blockDim = dim3(256); cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
        SGS_CUDA_TRY(cudaLaunchKernelExC(&cfg, fns[kCluster], args));
        launches++;
        *resume = 1;
    }
    SGS_CUDA_TRY(cudaLaunchCooperativeKernel(fns[kGrid], dim3(nsm), dim3(256), args, 0, st));
    launches++;
    return Status();
}

// levels m_hi-1 ... m_lo from level m_hi: one cooperative launch (behind a cluster-mode one), one synchronisation
Status KlocalMachine::Impl::run_sright(int m_hi, int m_lo)
{
    for (int attempt = 0;; attempt++) {
        int resume = 0;
        void *args[] = {&terms, &sr, &m_hi, &m_lo, &resume};
        const void *const fns[3] = {(const void *)k_sr_run<kGrid>, (const void *)k_sr_run<kCluster>, (const void *)k_sr_run<kBlock>};
        SGS_TRY(launch_persistent(fns, small_sr, args, &resume));
        SGS_CUDA_TRY(cudaMemcpyAsync(&hsr, sr.ctl, sizeof(SrCtl), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaStreamSynchronize(st));
        small_sr &= ~(unsigned)hsr.bailed;
        if (!hsr.err) break;
        if ((hsr.err != 4 && hsr.err != 8) || attempt >= 6) return ctl_error(hsr.err, "right environments");
        // scratch (4) or level pool (8) too small: take bigger ones and redo this run
        const int ncap = hsr.err == 4 ? (int)std::min<long long>((long long)sr.cap * 8, 1ll << 26) : sr.cap;
        const size_t npool = hsr.err == 8 ? (size_t)hsr.pool_cap * 4 : (size_t)hsr.pool_cap;
        if (hsr.err == 4) {
            SGS_TRY(alloc_sr(ncap, npool));                 // fresh scratch, table and control block
        } else {
            SGS_TRY(take(&sr.pool, npool));                 // levels already written stay where they are
            hsr.pool_used = 0; hsr.pool_cap = npool;
            hsr.err = hsr.err_snap = 0;
            SGS_CUDA_TRY(cudaMemcpyAsync(sr.ctl, &hsr, sizeof(SrCtl), cudaMemcpyHostToDevice, st));
        }
    }
    SGS_CUDA_TRY(cudaMemcpyAsync(&hlev[m_lo], d_levels + m_lo, sizeof(DpLevelDev) * (m_hi - m_lo), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    for (int m = m_lo; m < m_hi; m++) hgroups.erase(m);
    return Status();
}

Status KlocalMachine::Impl::host_groups(int m, const std::vector<DpGroup> **out)
{
    auto it = hgroups.find(m);
    if (it == hgroups.end()) {
        std::vector<DpGroup> g(hlev[m].count);
        if (!g.empty()) SGS_CUDA_TRY(cudaMemcpy(g.data(), hlev[m].groups, sizeof(DpGroup) * g.size(), cudaMemcpyDeviceToHost));
        it = hgroups.emplace(m, std::move(g)).first;
    }
    *out = &it->second;
    return Status();
}

Status KlocalMachine::Impl::gen_sright_all()
{
    SGS_TRY(make_top_level(n));
    return run_sright(n, 0);
}

Status KlocalMachine::Impl::write_ctl(int nin, bool reset_history)
{
    SGS_TRY(sync_ctl());
    hctl.nin = nin; hctl.total = 0; hctl.nuniq = nin; hctl.err = 0; hctl.err_snap = 0; hctl.cur = 0;
    if (reset_history) { hctl.hm_off = 0; hctl.hb_off = 0; hctl.step = 0; }
    SGS_CUDA_TRY(cudaMemcpyAsync(dp.ctl, &hctl, sizeof(DpCtl), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    hctl_valid = true;
    return Status();
}

// the one host synchronisation point of the DP: closes the timed segment and refreshes the control block
Status KlocalMachine::Impl::sync_ctl()
{
    SGS_CUDA_TRY(cudaSetDevice(device));
    if (hctl_valid && !seg_open) return Status();
    if (seg_open) SGS_CUDA_TRY(cudaEventRecord(e1, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(&hctl, dp.ctl, sizeof(DpCtl), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    if (seg_open) {
        float ms = 0.f;
        SGS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        dev_seconds += ms * 1e-3;
        seg_open = false;
    }
    hctl_valid = true;
    small_dp &= ~(unsigned)hctl.bailed;
    if (hctl.err) return ctl_error(hctl.err, "k-local DP");
    return Status();
}

Status KlocalMachine::Impl::set_states(const std::vector<HostState> &hs, int m, bool reset_history)
{
    if ((int)hs.size() > caps) return Status::fail(SGS_ERR_OVERFLOW, "k-local DP: more states than the batch capacity");
    m_ = m;
    epoch++;
    const size_t ns = std::max<size_t>(hs.size(), 1);
    std::vector<DpState> ss(ns);
    std::vector<double> ee(ns * tstride, 0.0);
    for (size_t i = 0; i < hs.size(); i++) {
        ss[i] = hs[i].s;
        for (size_t e = 0; e < hs[i].E.size() && e < (size_t)tstride; e++) ee[i * tstride + e] = hs[i].E[e];
    }
    SGS_TRY(write_ctl((int)hs.size(), reset_history));      // states go to buffer 0
    SGS_CUDA_TRY(cudaMemcpyAsync(dp.S[0], ss.data(), ns * sizeof(DpState), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(dp.E[0], ee.data(), ns * tstride * sizeof(double), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    return Status();
}

Status KlocalMachine::Impl::get_states(std::vector<HostState> &hs)
{
    SGS_TRY(sync_ctl());
    const int ns = hctl.nin;
    hs.assign(ns, HostState());
    if (!ns) return Status();
    std::vector<DpState> ss(ns);
    std::vector<double> ee((size_t)ns * tstride);
    SGS_CUDA_TRY(cudaMemcpyAsync(ss.data(), dp.S[hctl.cur], ns * sizeof(DpState), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(ee.data(), dp.E[hctl.cur], ee.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < ns; i++) {
        hs[i].s = ss[i];
        hs[i].E.assign(ee.begin() + (size_t)i * tstride, ee.begin() + (size_t)i * tstride + ((size_t)1 << ss[i].r));
    }
    return Status();
}

// StateMachine ctor (cpp/state.cpp:445-458): one state per right environment at site 0
Status KlocalMachine::Impl::init_states()
{
    std::vector<HostState> hs(hlev[0].count);
    for (size_t i = 0; i < hs.size(); i++) {
        memset(&hs[i].s, 0, sizeof(DpState));
        for (int w = 0; w < kDpVW; w++) hs[i].s.valid[w] = ~0ULL;
        hs[i].s.r = 0;
        hs[i].s.sright = (int)i;
        hs[i].E = {0.0};
    }
    return set_states(hs, 0, true);
}

// StateMachine::evolve (cpp/state.cpp:483-533) for nsteps consecutive sites: one cooperative launch, enqueued
// only — batch sizes live in the device control block, nothing is read back here.
Status KlocalMachine::Impl::advance(int nsteps)
{
    if (!pending.ok()) return pending;
    if (nsteps <= 0) return Status();
    if (m_ + nsteps > max_m || !hlev[m_].groups || !hlev[m_ + nsteps].groups)
        return Status::fail(SGS_ERR_ARG, "no right-environment level for this site");
    if (!seg_open) { SGS_CUDA_TRY(cudaEventRecord(e0, st)); seg_open = true; }
    hctl_valid = false;
    epoch++;
    int m0 = m_, resume = 0, part = 0, s_lo = 0, s_hi = 0;
    void *args[] = {&terms, &dp, &m0, &nsteps, &resume, &part, &s_lo, &s_hi};
    const void *const fns[3] = {(const void *)k_dp_run<kGrid>, (const void *)k_dp_run<kCluster>, (const void *)k_dp_run<kBlock>};
    SGS_TRY(launch_persistent(fns, small_dp, args, &resume));
    m_ += nsteps;
    return Status();
}

Status KlocalMachine::Impl::fetch_log(std::vector<DpStepLog> &log)
{
    SGS_TRY(sync_ctl());
    log.resize(std::min(hctl.step, log_cap));
    if (!log.empty()) SGS_CUDA_TRY(cudaMemcpy(log.data(), dp.log, sizeof(DpStepLog) * log.size(), cudaMemcpyDeviceToHost));
    return Status();
}

// history of (state, entry) of the current site back to the last cleared point, walked on the device
Status KlocalMachine::Impl::traceback(int state, int entry, std::vector<std::pair<int, int>> &out)
{
    SGS_TRY(sync_ctl());
    int *d_pos = d_trace, *d_sign = d_trace + trace_cap, *d_n = d_trace + 2 * trace_cap;
    k_dp_trace<<<1, 128, 0, st>>>(dp.ctl, dp.log, dp.hm, dp.hb, tstride, 0, state, entry, d_pos, d_sign, d_n, trace_cap);
    launches++;
    int nn = 0;
    SGS_CUDA_TRY(cudaMemcpyAsync(&nn, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> pos(std::max(nn, 1)), sg(std::max(nn, 1));
    if (nn) {
        SGS_CUDA_TRY(cudaMemcpy(pos.data(), d_pos, sizeof(int) * nn, cudaMemcpyDeviceToHost));
        SGS_CUDA_TRY(cudaMemcpy(sg.data(), d_sign, sizeof(int) * nn, cudaMemcpyDeviceToHost));
    }
    out.clear();
    for (int i = 0; i < nn; i++) out.push_back({H.order[pos[i]], sg[i]});
    return Status();
}

// StateMachine::generate_gs + State::get_stab_gs (cpp/state.cpp:421-442,535-550)
Status KlocalMachine::Impl::ground_state(sgs_result *out)
{
    SGS_CUDA_TRY(cudaSetDevice(device));        // (a multi-GPU driver may have left another device current)
    std::vector<HostState> hs;
    Status s = get_states(hs);
    if (!s.ok()) return s;
    if (hs.empty()) return Status::fail(SGS_ERR_DEAD, "the DP ended with no state");
    // first strict minimum over the states in key order, first minimum inside the table
    int bs = -1, be = 0;
    double bestE = 0.0;
    for (size_t i = 0; i < hs.size(); i++) {
        int arg = 0;
        for (size_t e = 1; e < hs[i].E.size(); e++) if (hs[i].E[e] < hs[i].E[arg]) arg = (int)e;
        if (bs < 0 || hs[i].E[arg] < bestE) { bestE = hs[i].E[arg]; bs = (int)i; be = arg; }
    }
    std::vector<DpStepLog> lg;
    s = fetch_log(lg);
    if (!s.ok()) return s;
    uint64_t transitions = 0;                   // evolve_single calls = states entering each site step
    for (const DpStepLog &e : lg) transitions += (uint64_t)e.nin;
    std::vector<std::pair<int, int>> hist;
    const auto tg0 = std::chrono::steady_clock::now();
    s = traceback(bs, be, hist);
    if (!s.ok()) return s;
    const auto tg1 = std::chrono::steady_clock::now();
    const int Wout = (2 * n + 63) / 64;
    const int mg = (int)hist.size();
    std::vector<uint64_t> rows((size_t)std::max(mg, 1) * Wout, 0);
    std::vector<int8_t> signs(std::max(mg, 1), 1);
    for (int i = 0; i < mg; i++) {
        H.row_words(hist[i].first, rows.data() + (size_t)i * Wout, Wout);
        signs[i] = (int8_t)hist[i].second;
    }
    int rankc = 0;
    std::vector<uint64_t> P((size_t)std::max(T, 1) * Wout, 0);
    for (int t = 0; t < T; t++) H.row_words(t, P.data() + (size_t)t * Wout, Wout);
    std::vector<int8_t> ts(std::max(T, 1), 0);
    s = group_finish(st, d_finish, n, mg, rows.data(), signs.data(), &rankc, T, P.data(), ts.data());   // K1 + K2 on the device
    if (!s.ok()) return s;
    if (getenv("SGS_DP_TRACE"))
        fprintf(stderr, "  ground state: traceback %.3f ms, canonicalise + term signs %.3f ms\n", std::chrono::duration<double>(tg1 - tg0).count() * 1e3,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tg1).count() * 1e3);
    out->energy = bestE;
    out->n = n; out->m = rankc; out->W = Wout; out->T = T;
    out->gens = (uint64_t *)calloc((size_t)std::max(rankc, 1) * Wout, 8);
    out->gen_signs = (int8_t *)calloc(std::max(rankc, 1), 1);
    out->term_signs = (int8_t *)calloc(std::max(T, 1), 1);
    memcpy(out->gens, rows.data(), (size_t)rankc * Wout * 8);
    memcpy(out->gen_signs, signs.data(), rankc);
    memcpy(out->term_signs, ts.data(), T);
    out->kernel_launches = launches + 2;
    out->transitions = transitions;
    out->seconds_device = dev_seconds;
    return Status();
}

// generate_Sright_periodic (py/state.py:542-567): backward sweeps over one period, shifted by +l, until the set
// of right environments at the period boundary is a fixed point; then one sweep over the whole unrolled chain.
// Fills sr_log with the (m, count) pairs the reference prints.
Status KlocalMachine::Impl::gen_sright_periodic(int cmax)
{
    const int l = H.l, kk = H.k;
    if (l <= 0) return Status::fail(SGS_ERR_ARG, "not a periodic Hamiltonian");
    const int end = l * (1 + cmax) + kk;
    if (end > max_m) return Status::fail(SGS_ERR_ARG, "periodic chain: cmax exceeds the unrolled length of the Hamiltonian");
    sr_log.clear();
    Status s = make_top_level(end);
    if (!s.ok()) return s;
    std::vector<DpGroup> id;
    bool have_id = false;
    for (;;) {
        s = run_sright(end, end - l);
        if (!s.ok()) return s;
        for (int m = end - 1; m >= end - l; m--) { sr_log.push_back(m); sr_log.push_back(hlev[m].count); }
        const std::vector<DpGroup> *cur = nullptr;
        s = host_groups(end - l, &cur);
        if (!s.ok()) return s;
        bool same = have_id && id.size() == cur->size();
        for (size_t i = 0; same && i < id.size(); i++) same = group_equal(id[i], (*cur)[i]);
        id = *cur;
        have_id = true;
        // stabs shifted by +l: same relative rows, now the level-`end` list
        s = set_level(end, hlev[end - l]);
        if (!s.ok()) return s;
        if (same) break;
    }
    s = run_sright(end, 0);
    if (!s.ok()) return s;
    for (int m = end - 1; m >= 0; m--) { sr_log.push_back(m); sr_log.push_back(hlev[m].count); }
    return Status();
}

// State.shift(add) (py/state.py:309-322): rows are origin-relative, so the shift is a relabelling of the level;
// only the right environment has to be looked up, by content, in the destination level.
Status KlocalMachine::Impl::shift_states(std::vector<HostState> &hs, int from_m, int add)
{
    if (from_m + add < 0 || from_m + add > max_m) return Status::fail(SGS_ERR_ARG, "shift leaves the chain");
    const std::vector<DpGroup> *A = nullptr, *B = nullptr;
    Status sg = host_groups(from_m, &A);
    if (sg.ok()) sg = host_groups(from_m + add, &B);
    if (!sg.ok()) return sg;
    for (auto &h : hs) {
        if (h.s.sright < 0 || h.s.sright >= (int)A->size()) return Status::fail(SGS_ERR_ARG, "state with an unknown right environment");
        const DpGroup &G = (*A)[h.s.sright];
        int found = -1;
        for (size_t i = 0; i < B->size(); i++) if (group_equal((*B)[i], G)) { found = (int)i; break; }
        if (found < 0) return Status::fail(SGS_ERR_ARG, "periodic shift: right environment not found at the shifted site (increase cmax)");
        h.s.sright = found;
    }
    return Status();
}

// State.__hash__ content (py/state.py:224-225; cpp/state.cpp:311-330): window group, the (in)valid indices of the
// terms whose last site is in [m+1, min(m+k-1, n)), right environment.  (m itself is added by the caller.)
std::vector<uint64_t> KlocalMachine::Impl::key_of(const HostState &h, int m) const
{
    const int klo = boff[std::min(m + 1, n + 1)] - boff[m];
    const int khi = boff[std::min(std::max(m + k - 1, m + 1), n)] - boff[m];
    std::vector<uint64_t> key;
    key.push_back((uint64_t)h.s.r);
    key.push_back((uint64_t)h.s.sright);
    for (int i = 0; i < kDpRows; i++) key.push_back(i < h.s.r ? h.s.g[i] : 0);
    uint64_t v[kDpVW] = {0, 0, 0, 0};
    for (int bb = klo; bb < khi; bb++) if ((h.s.valid[bb >> 6] >> (bb & 63)) & 1) v[bb >> 6] |= 1ULL << (bb & 63);
    for (int i = 0; i < kDpVW; i++) key.push_back(v[i]);
    return key;
}

// state_dict = {hash(state): state}: a later state with the same key replaces an earlier one; then key order.
// origin (optional) receives, for every kept state, its position in the input list.
void KlocalMachine::Impl::dedupe_sort(std::vector<HostState> &hs, int m, std::vector<int> *origin) const
{
    std::vector<HostState> outv;
    std::vector<int> org;
    for (size_t i = 0; i < hs.size(); i++) {
        bool rep = false;
        for (size_t o = 0; o < outv.size(); o++)
            if (key_of(outv[o], m) == key_of(hs[i], m)) { outv[o] = hs[i]; org[o] = (int)i; rep = true; break; }
        if (!rep) { outv.push_back(hs[i]); org.push_back((int)i); }
    }
    std::vector<int> perm(outv.size());
    for (size_t i = 0; i < perm.size(); i++) perm[i] = (int)i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key_of(outv[a], m) < key_of(outv[b], m); });
    std::vector<HostState> res;
    std::vector<int> ro;
    for (int i : perm) { res.push_back(outv[i]); ro.push_back(org[i]); }
    hs.swap(res);
    if (origin) origin->swap(ro);
}

// ------------------------------------------------------------------------------------------------ public

KlocalMachine::KlocalMachine() : impl_(new Impl()) {}
KlocalMachine::~KlocalMachine() { delete impl_; }

Status KlocalMachine::create(const Hamiltonian &H, const sgs_opts &opts, int periodic_cmax, std::unique_ptr<KlocalMachine> &out, int scale)
{
    std::unique_ptr<KlocalMachine> M(new KlocalMachine());
    M->impl_->H = H;
    M->impl_->opts = opts;
    M->impl_->scale = std::max(scale, 1);
    M->impl_->max_m = periodic_cmax >= 0 ? std::max(H.n, H.l * (1 + periodic_cmax) + H.k) : H.n;
    const auto t0 = std::chrono::steady_clock::now();
    Status s = M->impl_->init();
    if (!s.ok()) return s;
    if (periodic_cmax < 0) {
        const auto t1 = std::chrono::steady_clock::now();
        s = M->impl_->gen_sright_all();
        if (!s.ok()) return s;
        const auto t2 = std::chrono::steady_clock::now();
        s = M->impl_->init_states();
        if (!s.ok()) return s;
        if (getenv("SGS_DP_TRACE"))
            fprintf(stderr, "klocal create: init %.3f ms, right environments %.3f ms, initial states %.3f ms\n",
                    std::chrono::duration<double>(t1 - t0).count() * 1e3, std::chrono::duration<double>(t2 - t1).count() * 1e3,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count() * 1e3);
    }
    else if (H.l > 0) {
        // StateMachinePeriodic(H, generate_Sright_periodic(H, cmax)) (py/state.py:402-416,542-567)
        s = M->impl_->gen_sright_periodic(periodic_cmax);
        if (!s.ok()) return s;
        s = M->impl_->init_states();
        if (!s.ok()) return s;
    }
    out = std::move(M);
    return Status();
}

int KlocalMachine::m() const { return impl_->m_; }
int KlocalMachine::nstates() const
{
    Status s = impl_->sync_ctl();          // reads the device control block (one stream sync)
    if (!s.ok()) { impl_->pending = s; return -1; }
    return impl_->hctl.nin;
}
int KlocalMachine::sright_count(int m) const { return m < 0 || m >= (int)impl_->hlev.size() ? 0 : impl_->hlev[m].count; }
Status KlocalMachine::evolve() { return impl_->advance(1); }
Status KlocalMachine::ground_state(sgs_result *out) { return impl_->ground_state(out); }

// ------------------------------------------------------------------------------------------------ stepwise access
struct KlocalStates {
    std::vector<HostState> hs;
    std::vector<int> dev;      // device index each entry was read from (-1: none)
    int m = 0;                 // site of every state in the list
    int epoch = -1;            // machine epoch the device indices belong to
};

void klocal_states_free(KlocalStates *s) { delete s; }
int klocal_states_count(const KlocalStates *s) { return s ? (int)s->hs.size() : 0; }

// state_dict.values() of the machine's current site, in key order (the device order)
Status klocal_get_states(KlocalMachine &M, KlocalStates **out)
{
    KlocalMachine::Impl &K = *M.impl();
    std::unique_ptr<KlocalStates> S(new KlocalStates());
    Status s = K.get_states(S->hs);
    if (!s.ok()) return s;
    S->dev.resize(S->hs.size());
    for (size_t i = 0; i < S->dev.size(); i++) S->dev[i] = (int)i;
    S->m = K.m_;
    S->epoch = K.epoch;
    *out = S.release();
    return Status();
}

Status klocal_states_select(const KlocalStates *s, const int *idx, int count, KlocalStates **out)
{
    std::unique_ptr<KlocalStates> S(new KlocalStates());
    S->m = s->m; S->epoch = s->epoch;
    for (int i = 0; i < count; i++) {
        if (idx[i] < 0 || idx[i] >= (int)s->hs.size()) return Status::fail(SGS_ERR_ARG, "state index out of range");
        S->hs.push_back(s->hs[idx[i]]);
        S->dev.push_back(s->dev[idx[i]]);
    }
    *out = S.release();
    return Status();
}

// list building: dst.append(copy of src[i]); all states of a list sit at one site
Status klocal_states_append(KlocalStates *dst, const KlocalStates *src, int i)
{
    if (!src || i < 0 || i >= (int)src->hs.size()) return Status::fail(SGS_ERR_ARG, "state index out of range");
    if (dst->hs.empty()) { dst->m = src->m; dst->epoch = src->epoch; }
    else if (dst->m != src->m) return Status::fail(SGS_ERR_ARG, "states of different sites in one list");
    dst->hs.push_back(src->hs[i]);
    dst->dev.push_back(dst->epoch == src->epoch ? src->dev[i] : -1);
    return Status();
}

static Status check_index(const KlocalStates *s, int i)
{
    if (!s || i < 0 || i >= (int)s->hs.size()) return Status::fail(SGS_ERR_ARG, "state index out of range");
    return Status();
}

// State.m, rank and window of State.stab (py/state.py:150-176: the projected group sits on the k-1 sites below m)
Status klocal_state_info(const KlocalMachine &M, const KlocalStates *s, int i, int *m, int *rank, int *begin, int *end, int *sright)
{
    SGS_TRY(check_index(s, i));
    const KlocalMachine::Impl &K = *const_cast<KlocalMachine &>(M).impl();
    if (m) *m = s->m;
    if (rank) *rank = s->hs[i].s.r;
    if (begin) *begin = std::max(s->m - K.k + 1, 0);
    if (end) *end = s->m;
    if (sright) *sright = s->hs[i].s.sright;
    return Status();
}

// the generators of State.stab as absolute W-word rows (all signs '+': StabEnergies keeps them so, cpp/stab.cpp:626)
Status klocal_state_rows(const KlocalMachine &M, const KlocalStates *s, int i, uint64_t *rows)
{
    SGS_TRY(check_index(s, i));
    const KlocalMachine::Impl &K = *const_cast<KlocalMachine &>(M).impl();
    const int Wout = (2 * K.n + 63) / 64, o = s->m - K.k + 1;       // origin of the relative rows
    for (int g = 0; g < s->hs[i].s.r; g++) {
        for (int w = 0; w < Wout; w++) rows[(size_t)g * Wout + w] = 0;
        const uint64_t rel = s->hs[i].s.g[g];
        for (int j = 0; j < 32; j++) {
            const uint64_t v = (rel >> (2 * j)) & 3;
            const int site = o + j;
            if (v && site >= 0 && site < K.n) rows[(size_t)g * Wout + ((2 * site) >> 6)] |= v << ((2 * site) & 63);
        }
    }
    return Status();
}

Status klocal_state_key(const KlocalMachine &M, const KlocalStates *s, int i, uint64_t *key, int cap, int *nwords)
{
    SGS_TRY(check_index(s, i));
    const KlocalMachine::Impl &K = *const_cast<KlocalMachine &>(M).impl();
    std::vector<uint64_t> kk = K.key_of(s->hs[i], s->m);
    kk.insert(kk.begin(), (uint64_t)(int64_t)s->m);
    if ((int)kk.size() > cap) return Status::fail(SGS_ERR_ARG, "key buffer too small");
    memcpy(key, kk.data(), kk.size() * 8);
    *nwords = (int)kk.size();
    return Status();
}

// EnergyLevel.Es of the state (2^rank doubles, flat index bit (rank-1-i) set <=> generator i carries '-'); set != 0 writes
Status klocal_state_energies(const KlocalStates *s, int i, double *Es, int set)
{
    SGS_TRY(check_index(s, i));
    HostState &h = const_cast<KlocalStates *>(s)->hs[i];
    const size_t sz = (size_t)1 << h.s.r;
    if (set) h.E.assign(Es, Es + sz);
    else memcpy(Es, h.E.data(), sz * sizeof(double));
    return Status();
}

// State.shift(add, clear_history) for every state of the list (py/state.py:309-325)
Status klocal_states_shift(KlocalMachine &M, KlocalStates *s, int add, int clear_history)
{
    KlocalMachine::Impl &K = *M.impl();
    SGS_TRY(K.shift_states(s->hs, s->m, add));
    s->m += add;
    if (clear_history) {
        for (auto &h : s->hs) std::fill(h.E.begin(), h.E.end(), 0.0);     // EnergyLevel.clear (py/stab.py:282-286)
        for (auto &d : s->dev) d = -1;
    }
    return Status();
}

// StateMachinePeriodic.state_dict = {hash(state): state ...}; .m = m  (py/klocal_periodic.py:12-14): later states
// replace earlier ones with the same key, the machine holds them in key order; its history starts here.
Status klocal_set_states(KlocalMachine &M, const KlocalStates *s, int m)
{
    KlocalMachine::Impl &K = *M.impl();
    if (m != s->m) return Status::fail(SGS_ERR_ARG, "states of another site (shift them first)");
    std::vector<HostState> hs = s->hs;
    K.dedupe_sort(hs, m);
    return K.set_states(hs, m, true);
}

// EnergyLevel.get_state(index, original_paulis, return_stab=False) (py/stab.py:288-305): the signed Hamiltonian terms
// that generate the best full state reaching table entry `entry` of state i, in history order
Status klocal_state_history(KlocalMachine &M, const KlocalStates *s, int i, int entry, int *terms, int8_t *signs, int cap, int *count)
{
    SGS_TRY(check_index(s, i));
    KlocalMachine::Impl &K = *M.impl();
    if (s->dev[i] < 0) { *count = 0; return Status(); }                  // history cleared
    if (s->epoch != K.epoch) return Status::fail(SGS_ERR_ARG, "the machine has moved on: this state's history is no longer on the device");
    if (entry < 0 || entry >= (1 << s->hs[i].s.r)) return Status::fail(SGS_ERR_ARG, "table entry out of range");
    std::vector<std::pair<int, int>> h;
    SGS_TRY(K.traceback(s->dev[i], entry, h));
    if ((int)h.size() > cap) return Status::fail(SGS_ERR_ARG, "history buffer too small");
    for (size_t g = 0; g < h.size(); g++) { terms[g] = h[g].first; signs[g] = (int8_t)h[g].second; }
    *count = (int)h.size();
    return Status();
}

// relative rows (origin o: bit 2j = z, 2j+1 = x of site o + j) -> absolute Wout-word rows
static void rel_to_abs(uint64_t rel, int o, int n, uint64_t *row, int Wout)
{
    for (int w = 0; w < Wout; w++) row[w] = 0;
    for (int j = 0; j < 32; j++) {
        const uint64_t v = (rel >> (2 * j)) & 3;
        const int site = o + j;
        if (v && site >= 0 && site < n) row[(2 * site) >> 6] |= v << ((2 * site) & 63);
    }
}

// the i-th right environment of level m (std::get<0>(Sright_all.at(m)), cpp/state.cpp:218-238): a signed canonical
// group on the k sites ending at m
Status klocal_sright_group(KlocalMachine &M, int m, int i, int *rank, int *begin, int *end, uint64_t *rows, int8_t *signs)
{
    KlocalMachine::Impl &K = *M.impl();
    const std::vector<DpGroup> *G = nullptr;
    SGS_TRY(K.host_groups(m, &G));
    if (i < 0 || i >= (int)G->size()) return Status::fail(SGS_ERR_ARG, "right-environment index out of range");
    const DpGroup &g = (*G)[i];
    const int Wout = (2 * K.n + 63) / 64;
    if (rank) *rank = g.r;
    if (begin) *begin = std::max(m - K.k + 1, 0);
    if (end) *end = m + 1;                          // (the level-n group overhangs the chain by one site, cpp/state.cpp:219)
    for (int a = 0; a < g.r; a++) {
        if (rows) rel_to_abs(g.g[a], m - K.k + 1, K.n, rows + (size_t)a * Wout, Wout);
        if (signs) signs[a] = g.neg[a] ? -1 : 1;
    }
    return Status();
}

// State::get_valid_Ps(valid_range()) (cpp/state.cpp:293-309): file indices of the still-valid terms whose last site
// lies in [m + 1, min(m + k - 1, n)), bucket by bucket
Status klocal_state_valid_terms(const KlocalMachine &M, const KlocalStates *s, int i, int *terms, int cap, int *count)
{
    if (!s || i < 0 || i >= (int)s->hs.size()) return Status::fail(SGS_ERR_ARG, "state index out of range");
    const KlocalMachine::Impl &K = *const_cast<KlocalMachine &>(M).impl();
    const int m = s->m;
    const int klo = K.boff[std::min(m + 1, K.n + 1)] - K.boff[m];
    const int khi = K.boff[std::min(std::max(m + K.k - 1, m + 1), K.n)] - K.boff[m];
    int c = 0;
    for (int bb = klo; bb < khi; bb++) {
        if (!((s->hs[i].s.valid[bb >> 6] >> (bb & 63)) & 1)) continue;
        if (c >= cap) return Status::fail(SGS_ERR_ARG, "term buffer too small");
        terms[c++] = K.H.order[K.boff[m] + bb];
    }
    *count = c;
    return Status();
}

int klocal_sright_log(const KlocalMachine &M, int *pairs, int cap)
{
    const KlocalMachine::Impl &K = *const_cast<KlocalMachine &>(M).impl();
    const int np = (int)K.sr_log.size() / 2;
    for (int i = 0; i < np && i < cap; i++) { pairs[2 * i] = K.sr_log[2 * i]; pairs[2 * i + 1] = K.sr_log[2 * i + 1]; }
    return np;
}

// ------------------------------------------------------------------------------------------------ multi-GPU DP
// StateMachine::evolve (cpp/state.cpp:483-533) on several GPUs.  The reference parallelises the loop over the states
// of a site (omp parallel for, :494) and merges under a lock (:501); here the states of a site are REPLICATED on
// every GPU and the expensive half of a site step — State::move_to_next + project_to + check_Sright_validity of every
// state, the move phase of k_dp_run — is sharded: GPU g moves the states [g * per, (g + 1) * per), the results
// (moved states, their 2^rank tables with back-pointers, valid-term spans, fan-out counts, history adds) are
// exchanged by one NCCL group of in-place all-gathers over NVLink, and every GPU then runs the same deterministic
// emit / de-duplicate / order / merge on the whole batch, ending with identical state lists and histories (so the
// traceback needs no communication).  One host round trip per site: worth it from ~10^4 states per site on; below
// that one GPU's persistent kernel wins.  Errors met by one GPU in its slice travel with an
// all-reduce(max) of the control block's error word in the same group.
struct KlocalMulti {
    std::vector<std::unique_ptr<KlocalMachine>> M;      // one replica per local device
    std::shared_ptr<NcclWorld> nccl;

    Status create(const Hamiltonian &H, const sgs_opts &opts, int scale)
    {
        const int ng = std::max(1, opts.ngpus), world = std::max(1, opts.world);
        std::vector<int> ids;
        // the replicas (right environments included: "replicas only") are built concurrently, one host thread per device
        M.resize(ng);
        std::vector<Status> st(ng);
        std::vector<std::thread> th;
        for (int i = 0; i < ng; i++) {
            ids.push_back(opts.device + i);
            th.emplace_back([&, i]() {
                sgs_opts o = opts;
                o.device = opts.device + i; o.ngpus = 1;
                st[i] = KlocalMachine::create(H, o, -1, M[i], scale);
            });
        }
        for (auto &t : th) t.join();
        for (auto &x : st) if (!x.ok()) return x;
        return NcclWorld::get(ids, std::max(0, opts.rank), world, opts.nccl_id, nccl);
    }

    Status advance(int nsteps)
    {
        const int N = nccl->size();
        for (int it = 0; it < nsteps; it++) {
            for (auto &m : M) SGS_TRY(m->impl()->sync_ctl());
            KlocalMachine::Impl &K0 = *M[0]->impl();
            if (K0.hctl.err) return ctl_error(K0.hctl.err, "k-local DP");
            const int nin = K0.hctl.nin, hm_off = K0.hctl.hm_off;
            const int per = (nin + N - 1) / N;
            if ((long long)per * N > K0.caps || (long long)hm_off + (long long)per * N > K0.hm_cap)
                return Status::fail(SGS_ERR_OVERFLOW, "k-local DP: batch buffers too small for the per-site exchange");
            // move phase of this GPU's slice
            for (size_t i = 0; i < M.size(); i++) {
                KlocalMachine::Impl &K = *M[i]->impl();
                SGS_CUDA_TRY(cudaSetDevice(K.device));
                const int g = nccl->global_rank((int)i);
                int m0 = K.m_, one = 1, resume = 0, part = 1, s_lo = std::min(nin, g * per), s_hi = std::min(nin, (g + 1) * per);
                void *args[] = {&K.terms, &K.dp, &m0, &one, &resume, &part, &s_lo, &s_hi};
                if (!K.seg_open) { SGS_CUDA_TRY(cudaEventRecord(K.e0, K.st)); K.seg_open = true; }
                SGS_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_dp_run<kGrid>, dim3(K.nsm), dim3(256), args, 0, K.st));
                K.launches++;
            }
            // the per-site exchange
            if (per > 0) {
                std::vector<std::vector<void *>> base;
                std::vector<void *> errw;
                std::vector<cudaStream_t> sts;
                const size_t slab = (size_t)2 * K0.slab_stride;
                const std::vector<size_t> bytes = {(size_t)per * sizeof(DpMoved), (size_t)per * sizeof(Ech), (size_t)per * sizeof(int),
                                                   (size_t)per * slab * sizeof(double), (size_t)per * slab * sizeof(DpBack),
                                                   (size_t)per * sizeof(DpHistAdd)};
                for (auto &m : M) {
                    KlocalMachine::Impl &K = *m->impl();
                    base.push_back({K.dp.moved, K.dp.span, K.dp.nacc, K.dp.slabE, K.dp.slabB, K.dp.hm + hm_off});
                    errw.push_back(&K.dp.ctl->err);
                    sts.push_back(K.st);
                }
                SGS_TRY(nccl->gather_arrays_and_max(base, bytes, errw, sts));
            }
            auto launch_part = [&](int part, bool slice) -> Status {
                for (size_t i = 0; i < M.size(); i++) {
                    KlocalMachine::Impl &K = *M[i]->impl();
                    SGS_CUDA_TRY(cudaSetDevice(K.device));
                    int m0 = K.m_, one = 1, resume = 0, s_lo = slice ? nccl->global_rank((int)i) : 0, s_hi = slice ? N : 0;
                    void *args[] = {&K.terms, &K.dp, &m0, &one, &resume, &part, &s_lo, &s_hi};
                    SGS_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_dp_run<kGrid>, dim3(K.nsm), dim3(256), args, 0, K.st));
                    K.launches++;
                }
                return Status();
            };
            // scan, emit, de-duplicate, leaders: replicated on the whole batch
            SGS_TRY(launch_part(2, false));
            // the key-order ranking (all-pairs, the dominant phase of a large batch) is shared out by task;
            // the partial ranks are summed over the GPUs
            int total = 0, nuniq = 0;
            for (auto &m : M) {
                KlocalMachine::Impl &K = *m->impl();
                K.hctl_valid = false;
                SGS_TRY(K.sync_ctl());
                if (K.hctl.err) return ctl_error(K.hctl.err, "k-local DP");
                total = K.hctl.total; nuniq = K.hctl.nuniq;
            }
            // the compacted leader list is built with an atomic cursor: every GPU takes rank 0's order, so that task t
            // means the same (32 leaders) x (64 prefixes) everywhere
            if (nuniq > 0) {
                std::vector<std::vector<void *>> bufs;
                std::vector<cudaStream_t> sts;
                for (auto &m : M) { bufs.push_back({m->impl()->dp.D.lidx, m->impl()->dp.D.lpre}); sts.push_back(m->impl()->st); }
                SGS_TRY(nccl->broadcast_arrays(bufs, {(size_t)nuniq * sizeof(int), (size_t)nuniq * sizeof(uint64_t)}, sts));
            }
            SGS_TRY(launch_part(3, true));
            if (total > 0) {
                std::vector<void *> rk;
                std::vector<cudaStream_t> sts;
                for (auto &m : M) { rk.push_back(m->impl()->dp.D.rank); sts.push_back(m->impl()->st); }
                SGS_TRY(nccl->all_reduce_sum_i32(rk, (size_t)total, sts));
            }
            // merge and commit: replicated
            SGS_TRY(launch_part(4, false));
            for (auto &m : M) {
                KlocalMachine::Impl &K = *m->impl();
                K.m_++;
                K.hctl_valid = false;
                K.epoch++;
            }
        }
        for (auto &m : M) SGS_TRY(m->impl()->sync_ctl());
        return Status();
    }
};

Status klocal_search(const Hamiltonian &H, const sgs_opts &opts, sgs_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<KlocalMachine> M;
    std::vector<int> sr(H.n + 1), sp(H.n + 1, 0);
    std::vector<DpStepLog> dlog;
    double ms_create = 0, ms_dp = 0;
    const bool multi = opts.ngpus > 1 || (opts.world > 1 && opts.nccl_id != nullptr);
    KlocalMulti MM;
    for (int scale = 1;; scale *= 4) {
        const auto ta = std::chrono::steady_clock::now();
        M.reset();                         // hand the arena back before asking for a bigger one
        MM = KlocalMulti();
        Status s = multi ? MM.create(H, opts, scale) : KlocalMachine::create(H, opts, -1, M, scale);
        if (s.ok()) {
            if (multi) M = std::move(MM.M[0]);          // (results are read from the first local replica; it stays a member of MM below)
            const auto tb = std::chrono::steady_clock::now();
            ms_create = std::chrono::duration<double>(tb - ta).count() * 1e3;
            sp[0] = M->nstates();
            if (multi) {
                MM.M[0] = std::move(M);
                s = MM.advance(H.n);       // one exchange per site
                M = std::move(MM.M[0]);
            } else {
                s = M->impl()->advance(H.n);   // all n sites in one persistent launch
            }
            if (s.ok()) s = M->impl()->fetch_log(dlog);
            ms_dp = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count() * 1e3;
        }
        if (s.ok()) break;
        if (s.code != SGS_ERR_OVERFLOW || scale >= 1024) return s;     // a batch outgrew its buffers: redo with 4x the capacity
        if (multi) MM.M.clear();
    }
    for (int m = 0; m <= H.n; m++) sr[m] = M->sright_count(m);
    for (const DpStepLog &e : dlog) if (e.m + 1 <= H.n) sp[e.m + 1] = e.nuniq;
    const auto t2 = std::chrono::steady_clock::now();
    Status s = M->ground_state(out);
    if (!s.ok()) return s;
    if (getenv("SGS_DP_TRACE")) {
        fprintf(stderr, "klocal_search: create %.3f ms, DP %.3f ms, ground state %.3f ms\n", ms_create, ms_dp,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count() * 1e3);
        const KlocalMachine::Impl &K = *M->impl();
        fprintf(stderr, "  DP phases (us): move+fanout %.0f, scan %.0f, emit %.0f, dedup %.0f, leaders %.0f, order %.0f, merge %.0f\n",
                K.hctl.prof[0] * 1e-3, K.hctl.prof[1] * 1e-3, K.hctl.prof[2] * 1e-3, K.hctl.prof[3] * 1e-3, K.hctl.prof[4] * 1e-3,
                K.hctl.prof[5] * 1e-3, K.hctl.prof[6] * 1e-3);
        fprintf(stderr, "  right-environment phases (us): waves %.0f, project %.0f, dedup %.0f, leaders %.0f, order %.0f, "
                        "groups+links %.0f, offsets %.0f, fill %.0f\n",
                K.hsr.prof[0] * 1e-3, K.hsr.prof[1] * 1e-3, K.hsr.prof[2] * 1e-3, K.hsr.prof[3] * 1e-3, K.hsr.prof[4] * 1e-3,
                K.hsr.prof[5] * 1e-3, K.hsr.prof[6] * 1e-3, K.hsr.prof[7] * 1e-3);
    }
    out->nsites = H.n;
    out->sright_counts = (int *)malloc(sizeof(int) * (H.n + 1));
    out->states_per_site = (int *)malloc(sizeof(int) * (H.n + 1));
    memcpy(out->sright_counts, sr.data(), sizeof(int) * (H.n + 1));
    memcpy(out->states_per_site, sp.data(), sizeof(int) * (H.n + 1));
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return Status();
}

// ------------------------------------------------------------------------------------------------ periodic
// py/klocal_periodic.py:1-63 with generate_Sright_periodic (py/state.py:542-567) and StateMachinePeriodic
// (py/state.py:402-462).  Rows are origin-relative, so "shift by l sites" (Pauli.shift, State.shift
// py/state.py:309-322) is a relabelling of the level; only the right-environment index has to be looked up
// in the destination level.

static Status periodic_once(const Hamiltonian &H, int cmax, int c, const sgs_opts &opts, int scale, sgs_periodic_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    const double max_float = 1e100;                        // py/klocal_periodic.py:4
    std::unique_ptr<KlocalMachine> M;
    Status s = KlocalMachine::create(H, opts, cmax, M, scale);
    if (!s.ok()) return s;
    KlocalMachine::Impl &K = *M->impl();
    const int l = H.l, k = H.k, n = H.n;
    if (l <= 0) return Status::fail(SGS_ERR_ARG, "not a periodic Hamiltonian");
    std::vector<int> log, fixed;

    // ---- generate_Sright_periodic: done by create() ----
    log = K.sr_log;
    const int m0k = k + l - 1;
    auto same_key = [&](const HostState &a, const HostState &b) { return K.key_of(a, m0k) == K.key_of(b, m0k); };
    auto shift_states = [&](std::vector<HostState> &hs, int from_m, int add) { return K.shift_states(hs, from_m, add); };
    auto dedupe_sort = [&](std::vector<HostState> &hs) { K.dedupe_sort(hs, m0k); };
    // periodic_evolve_states (py/klocal_periodic.py:11-18)
    auto periodic_evolve = [&](std::vector<HostState> hs, bool clear_history, int cc, std::vector<HostState> &res) -> Status {
        dedupe_sort(hs);
        const int m0 = k + l - 1;
        Status st = K.set_states(hs, m0, true);
        if (!st.ok()) return st;
        st = K.advance(l * cc);
        if (!st.ok()) return st;
        st = K.get_states(res);
        if (!st.ok()) return st;
        st = shift_states(res, m0 + l * cc, -l * cc);
        if (!st.ok()) return st;
        if (clear_history) for (auto &h : res) std::fill(h.E.begin(), h.E.end(), 0.0);     // EnergyLevel.clear, py/stab.py:282-286
        return Status();
    };
    // ---- driver: py/klocal_periodic.py:21-34 (the initial states are set by create()) ----
    s = K.advance(k + 2 * l - 1);
    if (!s.ok()) return s;
    std::vector<HostState> states;
    s = K.get_states(states);
    if (!s.ok()) return s;
    s = shift_states(states, k + 2 * l - 1, -l);
    if (!s.ok()) return s;
    for (auto &h : states) std::fill(h.E.begin(), h.E.end(), 0.0);
    dedupe_sort(states);
    std::vector<HostState> prev;
    bool have_prev = false;
    for (;;) {
        std::vector<HostState> nst;
        s = periodic_evolve(states, true, 1, nst);
        if (!s.ok()) return s;
        dedupe_sort(nst);
        fixed.push_back((int)nst.size());
        bool same = have_prev && prev.size() == nst.size();
        for (size_t i = 0; same && i < nst.size(); i++) same = same_key(prev[i], nst[i]);
        states = nst;
        prev = nst;
        have_prev = true;
        if (same) break;
    }

    // ---- orbit search: py/klocal_periodic.py:35-58 ----
    struct Orb { double energy; int begin, end; std::vector<std::pair<int, int>> gens; };
    std::vector<Orb> orbits;
    for (size_t u = 0; u < states.size(); u++) {
        HostState ns = states[u];
        std::vector<HostState> all;
        s = periodic_evolve({ns}, true, c, all);
        if (!s.ok()) return s;
        bool recurs = false;
        for (auto &a : all) recurs |= same_key(a, ns);
        if (!recurs) continue;
        const size_t sz = (size_t)1 << ns.s.r;
        for (size_t index = 0; index < sz; index++) {
            ns.E.assign(sz, max_float);
            ns.E[index] = 0.0;
            s = periodic_evolve({ns}, false, c, all);
            if (!s.ok()) return s;
            for (size_t i = 0; i < all.size(); i++) {
                if (!same_key(all[i], ns)) continue;
                // EnergyLevel.get_state(index, original_paulis, return_stab=False), py/stab.py:288-305
                // `all` is in the device order of the final states (get_states keeps it), so i indexes K's history
                Orb ob;
                ob.energy = all[i].E[index];
                s = K.traceback((int)i, (int)index, ob.gens);
                if (!s.ok()) return s;
                ob.begin = n; ob.end = 0;
                for (auto &g : ob.gens) { ob.begin = std::min(ob.begin, H.begin[g.first]); ob.end = std::max(ob.end, H.end[g.first]); }
                if (ob.gens.empty()) { ob.begin = 0; ob.end = 0; }
                orbits.push_back(ob);
            }
        }
    }
    const int Wout = (2 * n + 63) / 64;
    auto gen_string = [&](const Orb &o, size_t g) {
        std::vector<uint64_t> rw(Wout);
        H.row_words(o.gens[g].first, rw.data(), Wout);
        return row_to_string(rw.data(), o.gens[g].second, o.begin, o.end);
    };
    std::sort(orbits.begin(), orbits.end(), [&](const Orb &a, const Orb &b) {
        if (a.energy != b.energy) return a.energy < b.energy;
        if (a.begin != b.begin) return a.begin < b.begin;
        if (a.end != b.end) return a.end < b.end;
        if (a.gens.size() != b.gens.size()) return a.gens.size() < b.gens.size();
        for (size_t g = 0; g < a.gens.size(); g++) {
            const std::string sa = gen_string(a, g), sb = gen_string(b, g);
            if (sa != sb) return sa < sb;
        }
        return false;
    });

    out->l = l; out->k = k; out->n = n; out->W = Wout; out->nterms = H.T();
    out->nlog = (int)log.size() / 2;
    out->sright_log = (int *)malloc(sizeof(int) * std::max<size_t>(log.size(), 1));
    memcpy(out->sright_log, log.data(), sizeof(int) * log.size());
    out->nfixed = (int)fixed.size();
    out->fixed_sizes = (int *)malloc(sizeof(int) * std::max<size_t>(fixed.size(), 1));
    memcpy(out->fixed_sizes, fixed.data(), sizeof(int) * fixed.size());
    out->norbits = (int)orbits.size();
    out->orbits = (sgs_orbit *)calloc(std::max<size_t>(orbits.size(), 1), sizeof(sgs_orbit));
    for (size_t i = 0; i < orbits.size(); i++) {
        sgs_orbit &o = out->orbits[i];
        o.energy = orbits[i].energy; o.begin = orbits[i].begin; o.end = orbits[i].end; o.m = (int)orbits[i].gens.size();
        o.gens = (uint64_t *)calloc(std::max(o.m, 1) * (size_t)Wout, 8);
        o.signs = (int8_t *)calloc(std::max(o.m, 1), 1);
        for (int g = 0; g < o.m; g++) {
            H.row_words(orbits[i].gens[g].first, o.gens + (size_t)g * Wout, Wout);
            o.signs[g] = (int8_t)orbits[i].gens[g].second;
        }
    }
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return Status();
}

Status klocal_periodic(const Hamiltonian &H, int cmax, int c, const sgs_opts &opts, sgs_periodic_result *out)
{
    for (int scale = 1;; scale *= 4) {
        Status s = periodic_once(H, cmax, c, opts, scale, out);
        if (s.code != SGS_ERR_OVERFLOW || scale >= 1024) return s;     // a batch outgrew its buffers: redo with 4x the capacity
    }
}

}  // namespace sgs
