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

#include "common.hpp"
#include "group_ops.hpp"
#include "klocal_kernels.cuh"
#include "klocal_search.hpp"

namespace sgs {

namespace {

template <typename T>
struct DBuf {                       // growable device buffer
    T *p = nullptr;
    size_t cap = 0;
    ~DBuf() { if (p) cudaFree(p); }
    cudaError_t need(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        cap = std::max<size_t>(n, cap * 2 + 64);
        return cudaMalloc((void **)&p, cap * sizeof(T));
    }
};

struct Level {                      // type_Sright (cpp/state.hpp:49-51) of one site
    std::vector<DpGroup> groups;    // sorted by content
    std::vector<int> map_off, map_idx;   // per group: indices into the next level's list
    int maxfan = 0;
    DpGroup *d_groups = nullptr;
    int *d_map_off = nullptr, *d_map_idx = nullptr;
    void free_dev() { cudaFree(d_groups); cudaFree(d_map_off); cudaFree(d_map_idx); d_groups = nullptr; d_map_off = d_map_idx = nullptr; }
};

struct Step {                       // history of one site step (what the reference keeps per table entry)
    std::vector<int> nadd;          // per parent state
    std::vector<int> added;         // per parent state, kDpMaxAdd term positions
    std::vector<DpBack> back;       // per new state * tstride
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
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint64_t *d_rows = nullptr;
    double *d_w = nullptr;
    int *d_boff = nullptr, *d_fidx = nullptr, *d_err = nullptr, *d_cnt = nullptr;
    std::vector<int> boff;
    DpTerms terms{};
    std::map<int, Level> lv;
    int m_ = 0, nstates_ = 0, tstride = 1, slab_stride = 1;
    DBuf<DpState> states, nstates_buf;
    DBuf<double> tabE, ntabE, slabE;
    DBuf<DpBack> tabB, ntabB, slabB;
    DBuf<DpMoved> moved;
    DBuf<unsigned char> accept;
    DBuf<int> nacc, off, parent, leader, rank;
    DBuf<DpKey> keys;
    DBuf<SrBranch> srscratch;
    DBuf<DpGroup> srout;
    std::vector<Step> steps;
    int hist_floor = 0;
    uint64_t transitions = 0;
    unsigned launches = 0;
    double dev_seconds = 0.0;

    ~Impl();
    Status init(int periodic_cmax);
    Status upload_level(Level &L);
    Status gen_level(int m, const Level &next, Level &out);
    Status gen_sright_all();
    Status init_states();
    Status evolve();
    Status set_states(const std::vector<HostState> &hs, int m);
    Status get_states(std::vector<HostState> &hs);
    Status ground_state(sgs_result *out);
    std::vector<std::pair<int, int>> traceback(int state, int entry) const;   // (file index, sign) in history order
    Status check_err(const char *what);
};

KlocalMachine::Impl::~Impl()
{
    cudaSetDevice(device);
    for (auto &kv : lv) kv.second.free_dev();
    cudaFree(d_rows); cudaFree(d_w); cudaFree(d_boff); cudaFree(d_fidx); cudaFree(d_err); cudaFree(d_cnt);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
}

Status KlocalMachine::Impl::check_err(const char *what)
{
    int err = 0;
    SGS_CUDA_TRY(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    if (err == 0) return Status();
    const char *why = err == 1 ? "window group rank exceeds the table limit (k too large)"
                      : err == 2 ? "more than 256 terms in k+2 consecutive last-site buckets"
                      : err == 3 ? "more than 30 terms share one last site"
                                 : "right-environment branch capacity exceeded";
    return Status::fail(SGS_ERR_LIMIT, std::string(what) + ": " + why);
}

Status KlocalMachine::Impl::init(int periodic_cmax)
{
    (void)periodic_cmax;
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
    SGS_CUDA_TRY(cudaMalloc((void **)&d_err, 4));
    SGS_CUDA_TRY(cudaMalloc((void **)&d_cnt, 16));
    SGS_CUDA_TRY(cudaMemcpy(d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(d_w, w.data(), w.size() * 8, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(d_boff, boff.data(), boff.size() * 4, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(d_fidx, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemset(d_err, 0, 4));
    terms.n = n; terms.k = k; terms.T = T; terms.W = W;
    terms.rows = d_rows; terms.w = d_w; terms.boff = d_boff; terms.file_index = d_fidx;
    tstride = 1 << std::min(k, kDpTab);
    slab_stride = 1 << std::min(k, kDpTab);
    return Status();
}

Status KlocalMachine::Impl::upload_level(Level &L)
{
    L.free_dev();
    const size_t ng = std::max<size_t>(L.groups.size(), 1);
    SGS_CUDA_TRY(cudaMalloc((void **)&L.d_groups, ng * sizeof(DpGroup)));
    SGS_CUDA_TRY(cudaMalloc((void **)&L.d_map_off, (L.groups.size() + 1) * sizeof(int)));
    SGS_CUDA_TRY(cudaMalloc((void **)&L.d_map_idx, std::max<size_t>(L.map_idx.size(), 1) * sizeof(int)));
    if (!L.groups.empty()) SGS_CUDA_TRY(cudaMemcpy(L.d_groups, L.groups.data(), L.groups.size() * sizeof(DpGroup), cudaMemcpyHostToDevice));
    if (L.map_off.size() != L.groups.size() + 1) L.map_off.assign(L.groups.size() + 1, 0);
    SGS_CUDA_TRY(cudaMemcpy(L.d_map_off, L.map_off.data(), L.map_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!L.map_idx.empty()) SGS_CUDA_TRY(cudaMemcpy(L.d_map_idx, L.map_idx.data(), L.map_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    L.maxfan = 0;
    for (size_t i = 0; i + 1 < L.map_off.size(); i++) L.maxfan = std::max(L.maxfan, L.map_off[i + 1] - L.map_off[i]);
    return Status();
}

// one backward step of generate_Sright_all (cpp/state.cpp:222-236): level m from level m+1
Status KlocalMachine::Impl::gen_level(int m, const Level &next, Level &out)
{
    const int nnext = (int)next.groups.size();
    const int nterm = boff[m + 1] - boff[m];
    int maxbranch = 1;
    for (int i = 0; i < nterm && maxbranch < (1 << 20); i++) maxbranch *= 2;
    if (nterm > 16) return Status::fail(SGS_ERR_LIMIT, "more than 16 terms share one last site (2^t right-environment branches)");
    SGS_CUDA_TRY(srscratch.need((size_t)nnext * 2 * maxbranch));
    SGS_CUDA_TRY(srout.need((size_t)nnext * maxbranch));
    SGS_CUDA_TRY(nacc.need(nnext + 1));
    k_sr_evolve<<<(nnext + 63) / 64, 64, 0, st>>>(terms, m, next.d_groups, nnext, maxbranch, srout.p, nacc.p, srscratch.p, d_err);
    launches++;
    SGS_CUDA_TRY(cudaGetLastError());
    std::vector<int> cnt(nnext);
    SGS_CUDA_TRY(cudaMemcpyAsync(cnt.data(), nacc.p, nnext * sizeof(int), cudaMemcpyDeviceToHost, st));
    Status s = check_err("right environments");
    if (!s.ok()) return s;
    // gather the children contiguously (device-to-device), then group by content on the device
    int total = 0;
    for (int c : cnt) total += c;
    DBuf<DpGroup> flat;
    SGS_CUDA_TRY(flat.need(std::max(total, 1)));
    std::vector<int> par;
    par.reserve(total);
    int at = 0;
    for (int p = 0; p < nnext; p++) {
        if (cnt[p]) SGS_CUDA_TRY(cudaMemcpyAsync(flat.p + at, srout.p + (size_t)p * maxbranch, cnt[p] * sizeof(DpGroup), cudaMemcpyDeviceToDevice, st));
        for (int j = 0; j < cnt[p]; j++) par.push_back(p);
        at += cnt[p];
    }
    SGS_CUDA_TRY(leader.need(std::max(total, 1)));
    SGS_CUDA_TRY(rank.need(std::max(total, 1)));
    SGS_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 16, st));
    if (total) {
        k_group_leader<DpGroup><<<(total + 63) / 64, 64, 0, st>>>(flat.p, total, leader.p);
        k_group_rank<DpGroup><<<(total + 63) / 64, 64, 0, st>>>(flat.p, total, leader.p, rank.p, d_cnt);
        launches += 2;
    }
    std::vector<int> hl(total), hr(total);
    std::vector<DpGroup> hg(total);
    if (total) {
        SGS_CUDA_TRY(cudaMemcpyAsync(hl.data(), leader.p, total * sizeof(int), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaMemcpyAsync(hr.data(), rank.p, total * sizeof(int), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaMemcpyAsync(hg.data(), flat.p, total * sizeof(DpGroup), cudaMemcpyDeviceToHost, st));
    }
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    int nuniq = 0;
    for (int c = 0; c < total; c++) if (hl[c] == c) nuniq++;
    out.groups.assign(nuniq, DpGroup());
    std::vector<std::vector<int>> parents(nuniq);
    for (int c = 0; c < total; c++) {
        const int g = hr[hl[c]];
        if (hl[c] == c) out.groups[g] = hg[c];
        parents[g].push_back(par[c]);
    }
    out.map_off.assign(nuniq + 1, 0);
    out.map_idx.clear();
    for (int g = 0; g < nuniq; g++) {        // duplicates removed, ascending (py/state.py:523 uses a set)
        auto &v = parents[g];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        out.map_idx.insert(out.map_idx.end(), v.begin(), v.end());
        out.map_off[g + 1] = (int)out.map_idx.size();
    }
    return upload_level(out);
}

Status KlocalMachine::Impl::gen_sright_all()
{
    Level top;
    DpGroup empty{};
    top.groups.push_back(empty);                 // Stab::empty(H.k, H.n - H.k + 1), cpp/state.cpp:219
    top.map_off = {0, 0};
    lv[n] = top;
    Status s = upload_level(lv[n]);
    if (!s.ok()) return s;
    for (int m = n - 1; m >= 0; m--) {
        Level L;
        s = gen_level(m, lv[m + 1], L);
        if (!s.ok()) return s;
        lv[m] = L;
        L.d_groups = nullptr; L.d_map_off = L.d_map_idx = nullptr;     // ownership moved into the map
    }
    return Status();
}

Status KlocalMachine::Impl::set_states(const std::vector<HostState> &hs, int m)
{
    nstates_ = (int)hs.size();
    m_ = m;
    const size_t ns = std::max<size_t>(hs.size(), 1);
    SGS_CUDA_TRY(states.need(ns));
    SGS_CUDA_TRY(tabE.need(ns * tstride));
    SGS_CUDA_TRY(tabB.need(ns * tstride));
    std::vector<DpState> ss(ns);
    std::vector<double> ee(ns * tstride, 0.0);
    for (size_t i = 0; i < hs.size(); i++) {
        ss[i] = hs[i].s;
        for (size_t e = 0; e < hs[i].E.size() && e < (size_t)tstride; e++) ee[i * tstride + e] = hs[i].E[e];
    }
    SGS_CUDA_TRY(cudaMemcpyAsync(states.p, ss.data(), ns * sizeof(DpState), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(tabE.p, ee.data(), ns * tstride * sizeof(double), cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    return Status();
}

Status KlocalMachine::Impl::get_states(std::vector<HostState> &hs)
{
    hs.assign(nstates_, HostState());
    if (!nstates_) return Status();
    std::vector<DpState> ss(nstates_);
    std::vector<double> ee((size_t)nstates_ * tstride);
    SGS_CUDA_TRY(cudaMemcpyAsync(ss.data(), states.p, nstates_ * sizeof(DpState), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(ee.data(), tabE.p, ee.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < nstates_; i++) {
        hs[i].s = ss[i];
        hs[i].E.assign(ee.begin() + (size_t)i * tstride, ee.begin() + (size_t)i * tstride + ((size_t)1 << ss[i].r));
    }
    return Status();
}

// StateMachine ctor (cpp/state.cpp:445-458): one state per right environment at site 0
Status KlocalMachine::Impl::init_states()
{
    std::vector<HostState> hs(lv[0].groups.size());
    for (size_t i = 0; i < hs.size(); i++) {
        memset(&hs[i].s, 0, sizeof(DpState));
        for (int w = 0; w < kDpVW; w++) hs[i].s.valid[w] = ~0ULL;
        hs[i].s.r = 0;
        hs[i].s.sright = (int)i;
        hs[i].E = {0.0};
    }
    steps.clear();
    hist_floor = 0;
    return set_states(hs, 0);
}

// StateMachine::evolve (cpp/state.cpp:483-533): every state of site m → merged states of site m+1
Status KlocalMachine::Impl::evolve()
{
    if (lv.find(m_) == lv.end() || lv.find(m_ + 1) == lv.end()) return Status::fail(SGS_ERR_ARG, "no right-environment level for this site");
    const Level &L = lv[m_], &Ln = lv[m_ + 1];
    const int nin = nstates_;
    Step step;
    if (nin == 0) { steps.push_back(step); m_++; return Status(); }
    const int maxfan = std::max(L.maxfan, 1);
    SGS_CUDA_TRY(moved.need(nin));
    SGS_CUDA_TRY(slabE.need((size_t)nin * 2 * slab_stride));
    SGS_CUDA_TRY(slabB.need((size_t)nin * 2 * slab_stride));
    SGS_CUDA_TRY(accept.need((size_t)nin * maxfan));
    SGS_CUDA_TRY(nacc.need(nin + 1));
    SGS_CUDA_TRY(off.need(nin + 1));
    SGS_CUDA_TRY(cudaEventRecord(e0, st));
    const int tb = 64, gb = (nin + tb - 1) / tb;
    k_dp_move<<<gb, tb, 0, st>>>(terms, m_, states.p, tabE.p, nin, tstride, L.d_groups, moved.p, slabE.p, slabB.p, slab_stride, d_err);
    k_dp_fanout<<<gb, tb, 0, st>>>(terms, m_, moved.p, nin, L.d_map_off, L.d_map_idx, Ln.d_groups, maxfan, accept.p, nacc.p);
    k_dp_scan<<<1, 1, 0, st>>>(nacc.p, nin, off.p, d_cnt);
    launches += 3;
    SGS_CUDA_TRY(cudaGetLastError());
    int total = 0;
    SGS_CUDA_TRY(cudaMemcpyAsync(&total, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    Status s = check_err("k-local DP step");
    if (!s.ok()) return s;
    transitions += (uint64_t)nin;
    // history of this step: what each parent added
    std::vector<DpMoved> hm(nin);
    SGS_CUDA_TRY(cudaMemcpyAsync(hm.data(), moved.p, nin * sizeof(DpMoved), cudaMemcpyDeviceToHost, st));
    int nuniq = 0;
    if (total > 0) {
        SGS_CUDA_TRY(keys.need(total));
        SGS_CUDA_TRY(parent.need(total));
        SGS_CUDA_TRY(leader.need(total));
        SGS_CUDA_TRY(rank.need(total));
        SGS_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 16, st));
        k_dp_emit<<<gb, tb, 0, st>>>(terms, m_, moved.p, nin, L.d_map_off, L.d_map_idx, maxfan, accept.p, off.p, keys.p, parent.p);
        k_group_leader<DpKey><<<(total + 63) / 64, 64, 0, st>>>(keys.p, total, leader.p);
        k_group_rank<DpKey><<<(total + 63) / 64, 64, 0, st>>>(keys.p, total, leader.p, rank.p, d_cnt);
        launches += 3;
        SGS_CUDA_TRY(cudaMemcpyAsync(&nuniq, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        SGS_CUDA_TRY(cudaStreamSynchronize(st));
        SGS_CUDA_TRY(nstates_buf.need(nuniq));
        SGS_CUDA_TRY(ntabE.need((size_t)nuniq * tstride));
        SGS_CUDA_TRY(ntabB.need((size_t)nuniq * tstride));
        k_dp_merge<<<total, 64, 0, st>>>(keys.p, parent.p, leader.p, rank.p, total, slabE.p, slabB.p, slab_stride, tstride,
                                         nstates_buf.p, ntabE.p, ntabB.p, moved.p);
        launches++;
        SGS_CUDA_TRY(cudaGetLastError());
        step.back.resize((size_t)nuniq * tstride);
        SGS_CUDA_TRY(cudaMemcpyAsync(step.back.data(), ntabB.p, step.back.size() * sizeof(DpBack), cudaMemcpyDeviceToHost, st));
    }
    SGS_CUDA_TRY(cudaEventRecord(e1, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    SGS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    dev_seconds += ms * 1e-3;
    step.nadd.resize(nin);
    step.added.resize((size_t)nin * kDpMaxAdd);
    for (int i = 0; i < nin; i++) {
        step.nadd[i] = hm[i].dead ? 0 : hm[i].nadd;
        for (int a = 0; a < kDpMaxAdd; a++) step.added[(size_t)i * kDpMaxAdd + a] = hm[i].added[a];
    }
    steps.push_back(std::move(step));
    std::swap(states.p, nstates_buf.p); std::swap(states.cap, nstates_buf.cap);
    std::swap(tabE.p, ntabE.p); std::swap(tabE.cap, ntabE.cap);
    std::swap(tabB.p, ntabB.p); std::swap(tabB.cap, ntabB.cap);
    nstates_ = nuniq;
    m_++;
    return Status();
}

// history of (state, entry) of the current site back to the last cleared point: (file index, sign) pairs in
// the order the reference appends them (Ps_indices / Ps_signs, cpp/stab.cpp:752-759)
std::vector<std::pair<int, int>> KlocalMachine::Impl::traceback(int state, int entry) const
{
    std::vector<std::vector<std::pair<int, int>>> chunks;
    int s = state, e = entry;
    for (int t = (int)steps.size() - 1; t >= hist_floor; t--) {
        const Step &S = steps[t];
        const DpBack &b = S.back[(size_t)s * tstride + e];
        if (b.state < 0) break;
        std::vector<std::pair<int, int>> ch;
        for (int a = 0; a < S.nadd[b.state]; a++)
            ch.push_back({H.order[S.added[(size_t)b.state * kDpMaxAdd + a]], ((b.signs >> a) & 1) ? -1 : 1});
        chunks.push_back(ch);
        s = b.state; e = b.entry;
    }
    std::vector<std::pair<int, int>> out;
    for (int i = (int)chunks.size() - 1; i >= 0; i--) out.insert(out.end(), chunks[i].begin(), chunks[i].end());
    return out;
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
    auto hist = traceback(bs, be);
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
    out->transitions = transitions;
    out->kernel_launches = launches + 2;
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
    Status s = M->impl_->init(periodic_cmax);
    if (!s.ok()) return s;
    if (periodic_cmax < 0) {
        s = M->impl_->gen_sright_all();
        if (!s.ok()) return s;
        s = M->impl_->init_states();
        if (!s.ok()) return s;
    }
    out = std::move(M);
    return Status();
}

int KlocalMachine::m() const { return impl_->m_; }
int KlocalMachine::nstates() const { return impl_->nstates_; }
int KlocalMachine::sright_count(int m) const
{
    auto it = impl_->lv.find(m);
    return it == impl_->lv.end() ? 0 : (int)it->second.groups.size();
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
    sp[0] = M->nstates();
    while (M->m() < H.n) {
        s = M->evolve();
        if (!s.ok()) return s;
        sp[M->m()] = M->nstates();
    }
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
    DpGroup empty{};
    cur.groups.push_back(empty);
    cur.map_off = {0, 0};
    s = K.upload_level(cur);
    if (!s.ok()) return s;
    std::vector<DpGroup> id;
    bool have_id = false;
    for (;;) {
        for (int m = end - 1; m >= end - l; m--) {
            Level nx;
            s = K.gen_level(m, cur, nx);
            if (!s.ok()) return s;
            cur.free_dev();
            cur = nx;
            log.push_back(m); log.push_back((int)cur.groups.size());
        }
        bool same = have_id && id.size() == cur.groups.size();
        for (size_t i = 0; same && i < id.size(); i++) same = group_equal(id[i], cur.groups[i]);
        id = cur.groups;
        have_id = true;
        // stabs shifted by +l: same relative rows, now the level-`end` list
        if (same) break;
    }
    K.lv[end] = cur;
    cur.d_groups = nullptr; cur.d_map_off = cur.d_map_idx = nullptr;
    for (int m = end - 1; m >= 0; m--) {
        Level L;
        s = K.gen_level(m, K.lv[m + 1], L);
        if (!s.ok()) return s;
        K.lv[m] = L;
        L.d_groups = nullptr; L.d_map_off = L.d_map_idx = nullptr;
        log.push_back(m); log.push_back((int)K.lv[m].groups.size());
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
        Status st = K.set_states(hs, m0);
        if (!st.ok()) return st;
        K.steps.clear();
        K.hist_floor = 0;
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
                ob.gens = K.traceback((int)i, (int)index);
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
