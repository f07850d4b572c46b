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
#include <vector>

#include "arena.hpp"
#include "common.hpp"
#include "group_ops.hpp"
#include "klocal_kernels.cuh"
#include "klocal_search.hpp"

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

struct Level {                      // type_Sright (cpp/state.hpp:49-51) of one site, device resident
    int count = 0, nlinks = 0, maxfan = 0;
    DpGroup *d_groups = nullptr;
    int *d_map_off = nullptr, *d_map_idx = nullptr;
    std::vector<DpGroup> groups;    // host copy (periodic driver only)
};

bool group_equal(const DpGroup &a, const DpGroup &b)
{
    if (a.r != b.r) return false;
    for (int i = 0; i < a.r; i++) if (a.g[i] != b.g[i] || a.neg[i] != b.neg[i]) return false;
    return true;
}

}  // namespace

struct HostState { DpState s; std::vector<double> E; };

struct KlocalMachine::Impl {
    Hamiltonian H;
    sgs_opts opts{};
    int device = 0;
    int n = 0, k = 0, T = 0, W = 1;
    bool keep_host_levels = false;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool seg_open = false;
    uint64_t *d_rows = nullptr;
    double *d_w = nullptr;
    int *d_boff = nullptr, *d_fidx = nullptr, *d_info = nullptr;
    DpCtl *d_ctl = nullptr;
    DpStepLog *d_log = nullptr;
    int log_cap = 1 << 14;
    std::vector<int> boff;
    DpTerms terms{};
    DevPool pool;
    std::map<int, Level> lv;
    int m_ = 0, tstride = 1, slab_stride = 1;
    // capacities of the DP batch buffers (grown x4 and the run restarted on overflow)
    int caps = 0, capc = 0, hm_cap = 0, hb_cap = 0;
    DBuf<DpState> statesA, statesB;
    DBuf<double> tabEA, tabEB, slabE;
    DBuf<DpBack> slabB, hb;
    DBuf<DpHistAdd> hm;
    DBuf<DpMoved> moved;
    DBuf<int> nacc, off, parent, leader, rank, trace;
    DBuf<DpKey> keys;
    // right-environment scratch
    DBuf<SrBranch> srscratch;
    DBuf<DpGroup> srout, srflat, srgroups;
    DBuf<int> srcnt, srpar, sroffp, srleader, srrank, srmapoff, srmapidx, srgcnt;
    bool cur_is_A = true;
    int hist_floor = 0;
    DpCtl hctl{};                  // last host copy of the control block
    bool hctl_valid = false;
    unsigned launches = 0;
    double dev_seconds = 0.0;
    Status pending;                // error met inside a const accessor, returned by the next call

    ~Impl();
    Status init();
    Status ensure_caps(int scale);
    Status gen_level(int m, const Level &next, Level &out);
    Status gen_sright_all();
    Status write_ctl(int nin, bool reset_history);
    Status sync_ctl();
    Status init_states();
    Status evolve();
    Status set_states(const std::vector<HostState> &hs, int m, bool reset_history);
    Status get_states(std::vector<HostState> &hs);
    Status ground_state(sgs_result *out);
    Status traceback(int state, int entry, std::vector<std::pair<int, int>> &out);   // (file index, sign) in history order
    Status fetch_log(std::vector<DpStepLog> &log);
    DpState *cur_states() { return cur_is_A ? statesA.p : statesB.p; }
    DpState *nxt_states() { return cur_is_A ? statesB.p : statesA.p; }
    double *cur_E() { return cur_is_A ? tabEA.p : tabEB.p; }
    double *nxt_E() { return cur_is_A ? tabEB.p : tabEA.p; }
};

KlocalMachine::Impl::~Impl()
{
    cudaSetDevice(device);
    cudaFree(d_rows); cudaFree(d_w); cudaFree(d_boff); cudaFree(d_fidx); cudaFree(d_info); cudaFree(d_ctl); cudaFree(d_log);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
}

static Status ctl_error(int err, const char *what)
{
    const char *why = err == 1 ? "window group rank exceeds the table limit (k too large)"
                      : err == 2 ? "more than 256 terms in k+2 consecutive last-site buckets"
                      : err == 3 ? "more than 30 terms share one last site"
                      : err == 4 ? "right-environment branch capacity exceeded"
                                 : "batch capacity exceeded";
    return Status::fail(err >= 5 ? SGS_ERR_OVERFLOW : SGS_ERR_LIMIT, std::string(what) + ": " + why);
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
    SGS_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    SGS_CUDA_TRY(cudaEventCreate(&e0));
    SGS_CUDA_TRY(cudaEventCreate(&e1));
    W = (2 * (n + 1) + 63) / 64;
    const int Tn = std::max(T, 1);
    std::vector<uint64_t> rows((size_t)Tn * W, 0);
    std::vector<double> w(Tn, 0.0);
    std::vector<int> fidx(Tn, 0);
    for (int p = 0; p < T; p++) {
        const int f = H.order[p];
        for (int i = 0; i < W; i++) rows[(size_t)p * W + i] = H.terms[f].w[i];
        w[p] = H.weights[f];
        fidx[p] = f;
    }
    boff = H.bucket_off;            // n + 1 entries, boff[n] = T
    boff.push_back(T);              // boff[n + 1]
    SGS_CUDA_TRY(cudaMalloc((void **)&d_rows, rows.size() * 8));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_w, w.size() * 8));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_boff, boff.size() * 4));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_fidx, fidx.size() * 4));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_info, 64));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_ctl, sizeof(DpCtl)));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_log, sizeof(DpStepLog) * log_cap));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_w, w.data(), w.size() * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_boff, boff.data(), boff.size() * 4, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_fidx, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemsetAsync(d_info, 0, 64, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    terms.n = n; terms.k = k; terms.T = T; terms.W = W;
    terms.rows = d_rows; terms.w = d_w; terms.boff = d_boff; terms.file_index = d_fidx;
    tstride = 1 << std::min(k, kDpTab);
    slab_stride = 1 << std::min(k, kDpTab);
    {   // one cached arena for the DP batch buffers, the right-environment scratch and the level tables
        const size_t c0 = (size_t)std::max<long long>(2048, (1ll << 22) / ((long long)slab_stride * 16));
        const size_t est = c0 * (2 * sizeof(DpState) + sizeof(DpMoved) + 8 * sizeof(DpKey) + 64 * sizeof(DpHistAdd) + 64) +
                           c0 * (size_t)tstride * 16 + c0 * 2 * (size_t)slab_stride * 16 +
                           c0 * 64 * (size_t)std::min(tstride, 64) * sizeof(DpBack) + ((size_t)96 << 20);
        SGS_CUDA_TRY(pool.open(device, est));
    }
    return ensure_caps(1);
}

// batch buffers of the DP; scale = 1 on creation, x4 per retry after an overflow
Status KlocalMachine::Impl::ensure_caps(int scale)
{
    const long long base = std::max<long long>(2048, (1ll << 22) / ((long long)slab_stride * 16));   // k=4: 16384 states
    caps = (int)std::min<long long>(base * scale, 1ll << 22);
    capc = (int)std::min<long long>((long long)caps * 8, 1ll << 24);
    hm_cap = (int)std::min<long long>((long long)caps * 64, 1ll << 26);
    hb_cap = (int)std::min<long long>((long long)caps * 64 * std::min(tstride, 64), 1ll << 28);
    SGS_CUDA_TRY(statesA.need(pool, caps)); SGS_CUDA_TRY(statesB.need(pool, caps));
    SGS_CUDA_TRY(tabEA.need(pool, (size_t)caps * tstride)); SGS_CUDA_TRY(tabEB.need(pool, (size_t)caps * tstride));
    SGS_CUDA_TRY(slabE.need(pool, (size_t)caps * 2 * slab_stride)); SGS_CUDA_TRY(slabB.need(pool, (size_t)caps * 2 * slab_stride));
    SGS_CUDA_TRY(moved.need(pool, caps)); SGS_CUDA_TRY(nacc.need(pool, caps + 1)); SGS_CUDA_TRY(off.need(pool, caps + 1));
    SGS_CUDA_TRY(keys.need(pool, capc)); SGS_CUDA_TRY(parent.need(pool, capc)); SGS_CUDA_TRY(leader.need(pool, capc)); SGS_CUDA_TRY(rank.need(pool, capc));
    SGS_CUDA_TRY(hm.need(pool, hm_cap)); SGS_CUDA_TRY(hb.need(pool, hb_cap));
    SGS_CUDA_TRY(trace.need(pool, 2 * (size_t)(T + 16) + 32));
    return Status();
}

// one backward step of generate_Sright_all (cpp/state.cpp:222-236): level m from level m+1, all on the device;
// the host only reads three integers (distinct groups, links, widest fan-out)
Status KlocalMachine::Impl::gen_level(int m, const Level &next, Level &out)
{
    const int nnext = next.count;
    const int nterm = boff[m + 1] - boff[m];
    if (nterm > 16) return Status::fail(SGS_ERR_LIMIT, "more than 16 terms share one last site (2^t right-environment branches)");
    const int maxbranch = 1 << nterm;
    const size_t bound = (size_t)nnext * maxbranch;
    SGS_CUDA_TRY(srscratch.need(pool, 2 * bound));
    SGS_CUDA_TRY(srout.need(pool, bound)); SGS_CUDA_TRY(srflat.need(pool, bound)); SGS_CUDA_TRY(srgroups.need(pool, bound));
    SGS_CUDA_TRY(srcnt.need(pool, nnext + 1)); SGS_CUDA_TRY(sroffp.need(pool, nnext + 1));
    SGS_CUDA_TRY(srpar.need(pool, bound)); SGS_CUDA_TRY(srleader.need(pool, bound)); SGS_CUDA_TRY(srrank.need(pool, bound));
    SGS_CUDA_TRY(srmapoff.need(pool, bound + 1)); SGS_CUDA_TRY(srmapidx.need(pool, bound)); SGS_CUDA_TRY(srgcnt.need(pool, bound + 1));
    int *d_total = d_info + 4, *d_nuniq = d_info + 5, *d_err = d_info + 6;
    SGS_CUDA_TRY(cudaMemsetAsync(d_info, 0, 32, st));
    k_sr_evolve<<<(nnext + 63) / 64, 64, 0, st>>>(terms, m, next.d_groups, nnext, maxbranch, srout.p, srcnt.p, srscratch.p, d_err);
    k_sr_flatten<<<1, 256, 0, st>>>(srout.p, srcnt.p, nnext, maxbranch, srflat.p, srpar.p, sroffp.p, d_total);
    const int gb = (int)std::min<size_t>((bound + 63) / 64, 512);
    k_group_leader<DpGroup><<<gb, 64, 0, st>>>(srflat.p, d_total, srleader.p);
    k_group_rank<DpGroup><<<gb, 64, 0, st>>>(srflat.p, d_total, srleader.p, srrank.p, d_nuniq);
    k_sr_build<<<1, 256, 0, st>>>(srflat.p, srpar.p, sroffp.p, srleader.p, srrank.p, d_total, d_nuniq, srgroups.p, srmapoff.p,
                                  srmapidx.p, srgcnt.p, d_info);
    launches += 5;
    SGS_CUDA_TRY(cudaGetLastError());
    int info[8];
    SGS_CUDA_TRY(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    if (info[6]) return ctl_error(info[6], "right environments");
    out.count = info[0]; out.nlinks = info[1]; out.maxfan = info[2];
    SGS_CUDA_TRY(pool.take((void **)&out.d_groups, sizeof(DpGroup) * std::max(out.count, 1)));
    SGS_CUDA_TRY(pool.take((void **)&out.d_map_off, sizeof(int) * (out.count + 1)));
    SGS_CUDA_TRY(pool.take((void **)&out.d_map_idx, sizeof(int) * std::max(out.nlinks, 1)));
    SGS_CUDA_TRY(cudaMemcpyAsync(out.d_groups, srgroups.p, sizeof(DpGroup) * out.count, cudaMemcpyDeviceToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(out.d_map_off, srmapoff.p, sizeof(int) * (out.count + 1), cudaMemcpyDeviceToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(out.d_map_idx, srmapidx.p, sizeof(int) * out.nlinks, cudaMemcpyDeviceToDevice, st));
    if (keep_host_levels) {
        out.groups.resize(out.count);
        SGS_CUDA_TRY(cudaMemcpyAsync(out.groups.data(), srgroups.p, sizeof(DpGroup) * out.count, cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return Status();
}

static Status make_top_level(KlocalMachine::Impl &K, Level &top);

Status KlocalMachine::Impl::gen_sright_all()
{
    Level top;
    Status s = make_top_level(*this, top);     // Stab::empty(H.k, H.n - H.k + 1), cpp/state.cpp:219
    if (!s.ok()) return s;
    lv[n] = top;
    for (int m = n - 1; m >= 0; m--) {
        Level L;
        s = gen_level(m, lv[m + 1], L);
        if (!s.ok()) return s;
        lv[m] = L;
    }
    return Status();
}

static Status make_top_level(KlocalMachine::Impl &K, Level &top)
{
    DpGroup empty{};
    const int zero2[2] = {0, 0};
    top.count = 1; top.nlinks = 0; top.maxfan = 0;
    SGS_CUDA_TRY(K.pool.take((void **)&top.d_groups, sizeof(DpGroup)));
    SGS_CUDA_TRY(K.pool.take((void **)&top.d_map_off, sizeof(int) * 2));
    SGS_CUDA_TRY(K.pool.take((void **)&top.d_map_idx, sizeof(int)));
    SGS_CUDA_TRY(cudaMemcpy(top.d_groups, &empty, sizeof(DpGroup), cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(top.d_map_off, zero2, sizeof(zero2), cudaMemcpyHostToDevice));
    top.groups.assign(1, empty);
    return Status();
}

Status KlocalMachine::Impl::write_ctl(int nin, bool reset_history)
{
    Status s = sync_ctl();
    if (!s.ok()) return s;
    DpCtl c = hctl;
    c.nin = nin; c.total = 0; c.nuniq = 0; c.err = 0;
    if (reset_history) { c.hm_off = 0; c.hb_off = 0; c.step = 0; hist_floor = 0; }
    c.caps = caps; c.capc = capc; c.hm_cap = hm_cap; c.hb_cap = hb_cap; c.log_cap = log_cap;
    hctl = c;
    SGS_CUDA_TRY(cudaMemcpyAsync(d_ctl, &hctl, sizeof(DpCtl), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    hctl_valid = true;
    return Status();
}

// the one host synchronisation point of the DP: closes the timed segment and refreshes the control block
Status KlocalMachine::Impl::sync_ctl()
{
    if (hctl_valid && !seg_open) return Status();
    if (seg_open) SGS_CUDA_TRY(cudaEventRecord(e1, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(&hctl, d_ctl, sizeof(DpCtl), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    if (seg_open) {
        float ms = 0.f;
        SGS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        dev_seconds += ms * 1e-3;
        seg_open = false;
    }
    hctl_valid = true;
    if (hctl.err) return ctl_error(hctl.err, "k-local DP");
    return Status();
}

Status KlocalMachine::Impl::set_states(const std::vector<HostState> &hs, int m, bool reset_history)
{
    if ((int)hs.size() > caps) return Status::fail(SGS_ERR_OVERFLOW, "k-local DP: more states than the batch capacity");
    m_ = m;
    const size_t ns = std::max<size_t>(hs.size(), 1);
    std::vector<DpState> ss(ns);
    std::vector<double> ee(ns * tstride, 0.0);
    for (size_t i = 0; i < hs.size(); i++) {
        ss[i] = hs[i].s;
        for (size_t e = 0; e < hs[i].E.size() && e < (size_t)tstride; e++) ee[i * tstride + e] = hs[i].E[e];
    }
    Status s = write_ctl((int)hs.size(), reset_history);
    if (!s.ok()) return s;
    SGS_CUDA_TRY(cudaMemcpyAsync(cur_states(), ss.data(), ns * sizeof(DpState), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(cur_E(), ee.data(), ns * tstride * sizeof(double), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    return Status();
}

Status KlocalMachine::Impl::get_states(std::vector<HostState> &hs)
{
    Status s = sync_ctl();
    if (!s.ok()) return s;
    const int ns = hctl.nin;
    hs.assign(ns, HostState());
    if (!ns) return Status();
    std::vector<DpState> ss(ns);
    std::vector<double> ee((size_t)ns * tstride);
    SGS_CUDA_TRY(cudaMemcpyAsync(ss.data(), cur_states(), ns * sizeof(DpState), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(ee.data(), cur_E(), ee.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
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
    std::vector<HostState> hs(lv[0].count);
    for (size_t i = 0; i < hs.size(); i++) {
        memset(&hs[i].s, 0, sizeof(DpState));
        for (int w = 0; w < kDpVW; w++) hs[i].s.valid[w] = ~0ULL;
        hs[i].s.r = 0;
        hs[i].s.sright = (int)i;
        hs[i].E = {0.0};
    }
    cur_is_A = true;
    return set_states(hs, 0, true);
}

// StateMachine::evolve (cpp/state.cpp:483-533): every state of site m → merged states of site m+1.
// Enqueue only: batch sizes live in the device control block, nothing is read back here.
Status KlocalMachine::Impl::evolve()
{
    if (lv.find(m_) == lv.end() || lv.find(m_ + 1) == lv.end()) return Status::fail(SGS_ERR_ARG, "no right-environment level for this site");
    if (!pending.ok()) return pending;
    const Level &L = lv[m_], &Ln = lv[m_ + 1];
    if (!seg_open) { SGS_CUDA_TRY(cudaEventRecord(e0, st)); seg_open = true; }
    hctl_valid = false;
    const int tb = 64, gs = std::min((caps + tb - 1) / tb, 256), gc = std::min((capc + tb - 1) / tb, 512);
    k_dp_move<<<gs, tb, 0, st>>>(terms, m_, cur_states(), cur_E(), d_ctl, tstride, L.d_groups, moved.p, slabE.p, slabB.p, slab_stride, hm.p);
    k_dp_fanout<<<gs, tb, 0, st>>>(terms, m_, moved.p, d_ctl, L.d_map_off, L.d_map_idx, Ln.d_groups, nacc.p);
    k_dp_scan<<<1, 1024, 0, st>>>(nacc.p, d_ctl, off.p);
    k_dp_emit<<<gs, tb, 0, st>>>(terms, m_, moved.p, d_ctl, L.d_map_off, L.d_map_idx, Ln.d_groups, off.p, keys.p, parent.p);
    k_group_leader<DpKey><<<gc, tb, 0, st>>>(keys.p, &d_ctl->total, leader.p);
    k_group_rank<DpKey><<<gc, tb, 0, st>>>(keys.p, &d_ctl->total, leader.p, rank.p, &d_ctl->nuniq);
    k_dp_merge<<<1024, 64, 0, st>>>(keys.p, parent.p, leader.p, rank.p, d_ctl, slabE.p, slabB.p, slab_stride, tstride, nxt_states(), nxt_E(),
                                    hb.p, moved.p);
    k_dp_advance<<<1, 1, 0, st>>>(d_ctl, d_log, m_, tstride);
    launches += 8;
    SGS_CUDA_TRY(cudaGetLastError());
    cur_is_A = !cur_is_A;
    m_++;
    return Status();
}

Status KlocalMachine::Impl::fetch_log(std::vector<DpStepLog> &log)
{
    Status s = sync_ctl();
    if (!s.ok()) return s;
    log.resize(std::min(hctl.step, log_cap));
    if (!log.empty()) SGS_CUDA_TRY(cudaMemcpy(log.data(), d_log, sizeof(DpStepLog) * log.size(), cudaMemcpyDeviceToHost));
    return Status();
}

// history of (state, entry) of the current site back to the last cleared point, walked on the device
Status KlocalMachine::Impl::traceback(int state, int entry, std::vector<std::pair<int, int>> &out)
{
    Status s = sync_ctl();
    if (!s.ok()) return s;
    const int cap = (int)(trace.cap / 2) - 8;
    int *d_pos = trace.p, *d_sign = trace.p + cap, *d_n = trace.p + 2 * cap;
    k_dp_trace<<<1, 1, 0, st>>>(d_ctl, d_log, hm.p, hb.p, tstride, hist_floor, state, entry, d_pos, d_sign, d_n, cap);
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
    s = traceback(bs, be, hist);
    if (!s.ok()) return s;
    const int Wout = (2 * n + 63) / 64;
    const int mg = (int)hist.size();
    std::vector<uint64_t> rows((size_t)std::max(mg, 1) * Wout, 0);
    std::vector<int8_t> signs(std::max(mg, 1), 1);
    for (int i = 0; i < mg; i++) {
        for (int w = 0; w < Wout; w++) rows[(size_t)i * Wout + w] = H.terms[hist[i].first].w[w];
        signs[i] = (int8_t)hist[i].second;
    }
    int rankc = 0;
    s = group_canonicalize(device, n, mg, rows.data(), signs.data(), &rankc);     // K1 on the device
    if (!s.ok()) return s;
    std::vector<uint64_t> P((size_t)std::max(T, 1) * Wout, 0);
    for (int t = 0; t < T; t++) for (int w = 0; w < Wout; w++) P[(size_t)t * Wout + w] = H.terms[t].w[w];
    std::vector<int8_t> ts(std::max(T, 1), 0);
    s = group_energy(device, n, rankc, rows.data(), signs.data(), T, P.data(), nullptr, ts.data());   // K2
    if (!s.ok()) return s;
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

// ------------------------------------------------------------------------------------------------ public

KlocalMachine::KlocalMachine() : impl_(new Impl()) {}
KlocalMachine::~KlocalMachine() { delete impl_; }

Status KlocalMachine::create(const Hamiltonian &H, const sgs_opts &opts, int periodic_cmax, std::unique_ptr<KlocalMachine> &out)
{
    std::unique_ptr<KlocalMachine> M(new KlocalMachine());
    M->impl_->H = H;
    M->impl_->opts = opts;
    M->impl_->keep_host_levels = periodic_cmax >= 0;
    const auto t0 = std::chrono::steady_clock::now();
    Status s = M->impl_->init();
    if (!s.ok()) return s;
    if (periodic_cmax < 0) {
        const auto t1 = std::chrono::steady_clock::now();
        s = M->impl_->gen_sright_all();
        if (!s.ok()) return s;
        if (getenv("SGS_DP_TRACE"))
            fprintf(stderr, "klocal create: init %.3f ms, right environments %.3f ms\n", std::chrono::duration<double>(t1 - t0).count() * 1e3,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count() * 1e3);
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
int KlocalMachine::sright_count(int m) const
{
    auto it = impl_->lv.find(m);
    return it == impl_->lv.end() ? 0 : it->second.count;
}
Status KlocalMachine::evolve() { return impl_->evolve(); }
Status KlocalMachine::ground_state(sgs_result *out) { return impl_->ground_state(out); }

Status klocal_search(const Hamiltonian &H, const sgs_opts &opts, sgs_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<KlocalMachine> M;
    Status s = KlocalMachine::create(H, opts, -1, M);
    if (!s.ok()) return s;
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<int> sr(H.n + 1), sp(H.n + 1);
    for (int m = 0; m <= H.n; m++) sr[m] = M->sright_count(m);
    KlocalMachine::Impl &K = *M->impl();
    std::vector<DpStepLog> dlog;
    for (int scale = 1;; scale *= 4) {
        sp[0] = M->nstates();
        while (M->m() < H.n) {             // n sites enqueued back to back, no host round trip in between
            s = M->evolve();
            if (!s.ok()) return s;
        }
        s = K.fetch_log(dlog);
        if (s.ok()) break;
        if (s.code != SGS_ERR_OVERFLOW || scale >= 4096) return s;
        s = K.ensure_caps(scale * 4);      // a batch outgrew its buffers: grow and redo the DP
        if (!s.ok()) return s;
        s = K.init_states();
        if (!s.ok()) return s;
    }
    for (const DpStepLog &e : dlog) if (e.m + 1 <= H.n) sp[e.m + 1] = e.nuniq;
    const auto t2 = std::chrono::steady_clock::now();
    s = M->ground_state(out);
    if (!s.ok()) return s;
    if (getenv("SGS_DP_TRACE")) {
        const auto t3 = std::chrono::steady_clock::now();
        fprintf(stderr, "klocal_search: right environments %.3f ms, DP %.3f ms, ground state %.3f ms\n",
                std::chrono::duration<double>(t1 - t0).count() * 1e3, std::chrono::duration<double>(t2 - t1).count() * 1e3,
                std::chrono::duration<double>(t3 - t2).count() * 1e3);
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

Status klocal_periodic(const Hamiltonian &H, int cmax, int c, const sgs_opts &opts, sgs_periodic_result *out)
{
    const auto t0 = std::chrono::steady_clock::now();
    const double max_float = 1e100;                        // py/klocal_periodic.py:4
    std::unique_ptr<KlocalMachine> M;
    Status s = KlocalMachine::create(H, opts, cmax, M);
    if (!s.ok()) return s;
    KlocalMachine::Impl &K = *M->impl();
    const int l = H.l, k = H.k, n = H.n;
    if (l <= 0) return Status::fail(SGS_ERR_ARG, "not a periodic Hamiltonian");
    std::vector<int> log, fixed;

    // ---- generate_Sright_periodic ----
    const int end = l * (1 + cmax) + k;
    Level cur;
    s = make_top_level(K, cur);
    if (!s.ok()) return s;
    std::vector<DpGroup> id;
    bool have_id = false;
    for (;;) {
        for (int m = end - 1; m >= end - l; m--) {
            Level nx;
            s = K.gen_level(m, cur, nx);
            if (!s.ok()) return s;
            cur = nx;
            log.push_back(m); log.push_back(cur.count);
        }
        bool same = have_id && id.size() == cur.groups.size();
        for (size_t i = 0; same && i < id.size(); i++) same = group_equal(id[i], cur.groups[i]);
        id = cur.groups;
        have_id = true;
        // stabs shifted by +l: same relative rows, now the level-`end` list
        if (same) break;
    }
    K.lv[end] = cur;
    for (int m = end - 1; m >= 0; m--) {
        Level L;
        s = K.gen_level(m, K.lv[m + 1], L);
        if (!s.ok()) return s;
        K.lv[m] = L;
        log.push_back(m); log.push_back(L.count);
    }

    // State.shift(add): relabel the level; the right environment is looked up by content
    auto shift_states = [&](std::vector<HostState> &hs, int from_m, int add) -> Status {
        const Level &A = K.lv[from_m], &B = K.lv[from_m + add];
        for (auto &h : hs) {
            const DpGroup &G = A.groups[h.s.sright];
            int found = -1;
            for (size_t i = 0; i < B.groups.size(); i++) if (group_equal(B.groups[i], G)) { found = (int)i; break; }
            if (found < 0) return Status::fail(SGS_ERR_ARG, "periodic shift: right environment not found at the shifted site (increase cmax)");
            h.s.sright = found;
        }
        return Status();
    };
    // State.__hash__ content (py/state.py:224-225): m, window group, (in)valid indices for last site in
    // [m+1, min(m+k-1, n)), right environment.  All lists compared here sit at m0 = k + l - 1.
    const int m0k = k + l - 1;
    const int klo = K.boff[std::min(m0k + 1, n + 1)] - K.boff[m0k];
    const int khi = K.boff[std::min(std::max(m0k + k - 1, m0k + 1), n)] - K.boff[m0k];
    auto key_of = [&](const HostState &h) {
        std::vector<uint64_t> key;
        key.push_back((uint64_t)h.s.r);
        key.push_back((uint64_t)h.s.sright);
        for (int i = 0; i < kDpRows; i++) key.push_back(i < h.s.r ? h.s.g[i] : 0);
        uint64_t v[kDpVW] = {0, 0, 0, 0};
        for (int bb = klo; bb < khi; bb++) if ((h.s.valid[bb >> 6] >> (bb & 63)) & 1) v[bb >> 6] |= 1ULL << (bb & 63);
        for (int i = 0; i < kDpVW; i++) key.push_back(v[i]);
        return key;
    };
    auto same_key = [&](const HostState &a, const HostState &b) { return key_of(a) == key_of(b); };
    // state_dict = {hash(state): state}: a later state with the same key replaces an earlier one; key order
    auto dedupe_sort = [&](std::vector<HostState> &hs) {
        std::vector<HostState> outv;
        for (auto &h : hs) {
            bool rep = false;
            for (auto &o : outv) if (same_key(o, h)) { o = h; rep = true; break; }
            if (!rep) outv.push_back(h);
        }
        std::stable_sort(outv.begin(), outv.end(), [&](const HostState &a, const HostState &b) { return key_of(a) < key_of(b); });
        hs.swap(outv);
    };
    // periodic_evolve_states (py/klocal_periodic.py:11-18)
    auto periodic_evolve = [&](std::vector<HostState> hs, bool clear_history, int cc, std::vector<HostState> &res) -> Status {
        dedupe_sort(hs);
        const int m0 = k + l - 1;
        Status st = K.set_states(hs, m0, true);
        if (!st.ok()) return st;
        for (int i = 0; i < l * cc; i++) { st = K.evolve(); if (!st.ok()) return st; }
        st = K.get_states(res);
        if (!st.ok()) return st;
        st = shift_states(res, m0 + l * cc, -l * cc);
        if (!st.ok()) return st;
        if (clear_history) for (auto &h : res) std::fill(h.E.begin(), h.E.end(), 0.0);     // EnergyLevel.clear, py/stab.py:282-286
        return Status();
    };
    // ---- driver: py/klocal_periodic.py:21-34 ----
    s = K.init_states();
    if (!s.ok()) return s;
    for (int i = 0; i < k + 2 * l - 1; i++) { s = K.evolve(); if (!s.ok()) return s; }
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
        return row_to_string(H.terms[o.gens[g].first].w, o.gens[g].second, o.begin, o.end);
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
            for (int w = 0; w < Wout; w++) o.gens[(size_t)g * Wout + w] = H.terms[orbits[i].gens[g].first].w[w];
            o.signs[g] = (int8_t)orbits[i].gens[g].second;
        }
    }
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return Status();
}

}  // namespace sgs
