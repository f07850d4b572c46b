// klocal_kernels.cuh — sm_100a kernels of the 1D k-local dynamic programme.
//
// What they replace (reference): StabEnergies::{add, transform, flip_sign, add_energy,
// project_to_next, merge_gs} (cpp/stab.cpp:667-843), Stab::{canonicalize, decompose, include, commute, project}
// (cpp/stab.cpp:437-577), State::{add, move_to_next, check_Sright_validity, hash_value} (cpp/state.cpp:311-413),
// StateMachine::{evolve_single, evolve} (cpp/state.cpp:460-533) and evolve_stab / get_map_from_stabs
// (cpp/state.cpp:129-216).
//
// Design (not a port): every object a DP step touches lives on at most 2k sites, so all rows are ONE u64
// relative to the step's origin o_m = m - k + 1 (site s at bits 2(s-o), 2(s-o)+1).  One thread owns one
// state / one right-environment group; the 2^rank sign tables and their back-pointers live in a per-state
// global scratch slab.  Instead of the reference's per-entry history arrays (2n ints per table entry,
// cpp/stab.cpp:752-759) every table entry keeps one back-pointer (parent state, parent entry, signs of the
// generators added at this site); the generator list is recovered by a traceback at the end.
// The reference's keyed merge (std::map on a boost hash + omp critical, cpp/state.cpp:494-517) becomes a
// deterministic group-by-full-key with "later in (parent, child) order wins exact ties" = its strict '<'.
#pragma once
#include <cstdint>

#include "packed.cuh"

namespace sgs {

constexpr int kDpMaxK = 15;        // rows must fit 2k+2 sites in one u64; a group on k+1 sites has rank <= 16
constexpr int kDpRows = 16;        // max rank of a window group
constexpr int kDpTab = 10;         // max rank with a materialised 2^rank table
constexpr int kDpVW = 4;           // valid-term bitset words (terms in k+2 consecutive buckets <= 256)
constexpr int kDpMaxAdd = 16;      // generators added at one site
constexpr int kDpMaxSr = 2 * kDpRows;

struct DpTerms {                   // Hamiltonian in paulis_list order
    int n, k, T, W;
    const uint64_t *rows;          // [T * W] absolute rows
    const double *w;               // [T]
    const int *boff;               // [n + 2] bucket offsets by last site (boff[n+1] = T)
    const int *file_index;         // [T] position -> file index
};

struct DpGroup {                   // a right-environment group (level list entry), origin o_m
    uint64_t g[kDpRows];
    uint8_t neg[kDpRows];          // 1 if the generator's sign is '-'
    int r;
};

struct DpState {
    uint64_t g[kDpRows];           // canonical all-'+' generators, origin o_m
    uint64_t valid[kDpVW];         // bit i: term position boff[m] + i still valid
    int r;
    int sright;                    // index into level m's group list
};

struct DpKey {                     // State::hash_value content (cpp/state.cpp:311-330)
    uint64_t g[kDpRows];
    uint64_t v[kDpVW];
    int r, sright;
};

struct DpBack {                    // back-pointer of one table entry
    int state;                     // parent state index at the previous site (-1: history starts here)
    uint16_t entry;                // parent table entry
    uint16_t signs;                // bit a: the a-th generator added at this site carries '-'
};

// ---- 64-bit window extraction from an absolute W-word row: bits [2*o, 2*o + 64) (o may be negative)
SGS_D uint64_t rel_row(const uint64_t *row, int W, int o)
{
    const int sh = 2 * o;
    uint64_t out = 0;
    for (int w = 0; w < W; w++) {
        const int lo = w * 64 - sh;              // position of word w's bit 0 in the output
        if (lo <= -64 || lo >= 64) continue;
        out |= lo >= 0 ? (row[w] << lo) : (row[w] >> (-lo));
    }
    return out;
}

SGS_D int q_plus(uint64_t r) { return (4 - (popc64(r & (r >> 1) & kEven) & 3)) & 3; }
SGS_D int q_sign(uint64_t r, int q) { return ((q + popc64(r & (r >> 1) & kEven)) & 3) == 0 ? 1 : -1; }
SGS_D bool anti(uint64_t a, uint64_t b) { return popc64(a & swapzx(b)) & 1; }
SGS_D void mul1(uint64_t &a, int &qa, uint64_t b, int qb)
{
    qa = (qa + qb + 2 * popc64((a >> 1) & b & kEven)) & 3;
    a ^= b;
}

// membership of v in the span of RREF rows g[0..r) (bits only)
SGS_D bool in_rref(const uint64_t *g, int r, uint64_t v)
{
    for (int i = 0; i < r; i++) if ((v >> ctz64(g[i])) & 1) v ^= g[i];
    return v == 0;
}

// echelon basis (arbitrary order, each row clear at earlier pivots) — bits only
struct Ech { uint64_t b[kDpMaxSr]; int n; };
SGS_D uint64_t ech_reduce(const Ech &e, uint64_t v)
{
    for (int i = 0; i < e.n; i++) if ((v >> ctz64(e.b[i])) & 1) v ^= e.b[i];
    return v;
}
SGS_D void ech_add(Ech &e, uint64_t v)
{
    v = ech_reduce(e, v);
    if (v && e.n < kDpMaxSr) e.b[e.n++] = v;
}

// multiword right shift of the valid bitset by s bits, shifting in zeros
SGS_D void vshift_right(uint64_t *v, int s)
{
    const int ws = s >> 6, bs = s & 63;
    for (int i = 0; i < kDpVW; i++) {
        uint64_t lo = i + ws < kDpVW ? v[i + ws] : 0, hi = i + ws + 1 < kDpVW ? v[i + ws + 1] : 0;
        v[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
}
SGS_D bool vget(const uint64_t *v, int i) { return (v[i >> 6] >> (i & 63)) & 1; }
SGS_D void vclr(uint64_t *v, int i) { v[i >> 6] &= ~(1ULL << (i & 63)); }
SGS_D void vset_range(uint64_t *v, int a, int b) { for (int i = a; i < b; i++) v[i >> 6] |= 1ULL << (i & 63); }

// ------------------------------------------------------------------------------------------------
// canonical insertion of one '+' generator into a canonical all-'+' group with its sign table
// (StabEnergies::add, cpp/stab.cpp:747-761 = concat_last + canonicalize :693-697 = transform :699-733 +
// canonicalize_signs :685-691).  E/bk are the table and back-pointers (2^r entries in, 2^(r+1) out), tmpE/
// tmpB scratch of the same size; `abit` is the sign bit this addition occupies in DpBack::signs.
SGS_D void group_add(uint64_t *g, int &r, uint64_t P, double *E, DpBack *bk, double *tmpE, DpBack *tmpB, int abit)
{
    // 1. new generator enters as the LAST axis: new = 2*old + b, b = 1 <=> '-'
    const int sz = 1 << r;
    for (int o = sz - 1; o >= 0; o--) {
        const double e = E[o];
        DpBack b = bk[o];
        tmpE[2 * o] = e; tmpE[2 * o + 1] = e;
        tmpB[2 * o] = b;
        b.signs = (uint16_t)(b.signs | (1u << abit));
        tmpB[2 * o + 1] = b;
    }
    // 2. RREF with the new row; U[i] = which of (old generators 0..r-1, new = bit r) multiply into row i
    uint32_t U[kDpRows + 1];
    int q[kDpRows + 1];
    for (int i = 0; i < r; i++) { U[i] = 1u << i; q[i] = q_plus(g[i]); }
    uint64_t v = P;
    int qv = q_plus(P);
    uint32_t uv = 1u << r;
    for (int i = 0; i < r; i++)
        if ((v >> ctz64(g[i])) & 1) { mul1(v, qv, g[i], q[i]); uv |= 1u << i; }
    const int pv = ctz64(v);
    for (int i = 0; i < r; i++)
        if ((g[i] >> pv) & 1) { mul1(g[i], q[i], v, qv); U[i] ^= uv; }
    int pos = r;
    while (pos > 0 && ctz64(g[pos - 1]) > pv) { g[pos] = g[pos - 1]; U[pos] = U[pos - 1]; q[pos] = q[pos - 1]; pos--; }
    g[pos] = v; U[pos] = uv; q[pos] = qv;
    r++;
    // 3. the pattern with sign vector s moves to s' = U s, then axes of generators whose product came out
    //    '-' are reversed so that every stored generator is '+'
    uint32_t flip = 0;
    for (int i = 0; i < r; i++) if (q_sign(g[i], q[i]) < 0) flip |= 1u << (r - 1 - i);
    const int nsz = 1 << r;
    for (int s = 0; s < nsz; s++) {
        uint32_t smask = 0;
        for (int j = 0; j < r; j++) if ((s >> (r - 1 - j)) & 1) smask |= 1u << j;
        uint32_t ns = 0;
        for (int i = 0; i < r; i++) if (__popc(U[i] & smask) & 1) ns |= 1u << (r - 1 - i);
        ns ^= flip;
        E[ns] = tmpE[s];
        bk[ns] = tmpB[s];
    }
}

// StabEnergies::add_energy (cpp/stab.cpp:819-825) for a member term of a canonical all-'+' group
SGS_D void group_add_energy(const uint64_t *g, int r, uint64_t P, double w, double *E)
{
    uint64_t acc = P;
    int qa = q_plus(P);
    uint32_t xi = 0;
    for (int i = 0; i < r; i++)
        if ((acc >> ctz64(g[i])) & 1) { mul1(acc, qa, g[i], q_plus(g[i])); xi |= 1u << (r - 1 - i); }
    const double ws = (qa & 2) ? -w : w;            // w * sign
    const int sz = 1 << r;
    for (int o = 0; o < sz; o++) E[o] += (__popc(o & xi) & 1) ? -ws : ws;
}

// StabEnergies::project_to_next (cpp/stab.cpp:782-801) for relative site 0
SGS_D void group_project_site0(uint64_t *g, int &r, double *E, DpBack *bk)
{
    while (r > 0) {
        bool any = false;
        for (int i = 0; i < r; i++) any |= (g[i] & 3ULL) != 0;
        if (!any) break;
        const int half = 1 << (r - 1);
        for (int o = 0; o < half; o++)
            if (!(E[o] < E[o + half])) { E[o] = E[o + half]; bk[o] = bk[o + half]; }   // strict '<': a tie keeps '-'
        for (int i = 0; i + 1 < r; i++) g[i] = g[i + 1];
        r--;
    }
}

// ------------------------------------------------------------------------------------------------
// Device-resident control block of the DP: the host enqueues every site step back to back without
// reading anything; the kernels take the batch sizes from here (grid-stride loops under fixed grids).
struct DpCtl {
    int nin;                 // states entering the current step
    int total;               // children emitted by the current step
    int nuniq;               // merged states = nin of the next step
    int err;                 // 0 ok, 1 rank/table limit, 2 valid-set width, 5 state capacity, 6 child capacity, 7 history capacity
    int hm_off, hb_off;      // bump offsets into the history pools
    int step;                // steps logged so far
    int caps, capc, hm_cap, hb_cap, log_cap;
};
struct DpStepLog { int m, nin, total, nuniq, hm_off, hb_off; };
struct DpHistAdd { int nadd; uint16_t added[kDpMaxAdd]; };      // what one parent state added at one site

struct DpMoved {
    uint64_t g[kDpRows];
    uint64_t valid[kDpVW];
    int r, sright, dead, nadd;
};

// State::move_to_next + project_to(k-1) (cpp/state.cpp:382-413, :460-466) for one state.
SGS_D void dp_move_one(const DpTerms &H, int m, int s, const DpState &st, const double *inE, int tstride,
                       const DpGroup *level, DpMoved &mv, double *E, double *tE, DpBack *B, DpBack *tB, DpHistAdd *hist, int *err)
{
    const int k = H.k, o = m - k + 1;
    uint64_t g[kDpRows + 1];
    int r = st.r;
    for (int i = 0; i < r; i++) g[i] = st.g[i];
    uint64_t valid[kDpVW];
    for (int i = 0; i < kDpVW; i++) valid[i] = st.valid[i];
    for (int e = 0; e < (1 << r); e++) { E[e] = inE[(size_t)s * tstride + e]; B[e].state = s; B[e].entry = (uint16_t)e; B[e].signs = 0; }
    const DpGroup &sr = level[st.sright];
    mv.sright = st.sright; mv.dead = 0; mv.nadd = 0; mv.r = 0;
    if (hist) hist->nadd = 0;

    // stab_tmp = <stab, Sright> (cpp/state.cpp:384-390), bits only
    Ech tmp; tmp.n = 0;
    for (int i = 0; i < r; i++) ech_add(tmp, g[i]);
    for (int i = 0; i < sr.r; i++) ech_add(tmp, sr.g[i]);
    const int b0 = H.boff[m], b1 = H.boff[m + 1];
    int nadd = 0;
    for (int pos = b0; pos < b1; pos++) {
        const uint64_t P = rel_row(H.rows + (size_t)pos * H.W, H.W, o);
        if (ech_reduce(tmp, P) != 0) continue;
        if (!in_rref(sr.g, sr.r, P)) { mv.dead = 1; return; }          // cpp/state.cpp:397-399
        if (in_rref(g, r, P)) continue;
        if (r + 1 > kDpTab || r + 1 > kDpRows || nadd >= kDpMaxAdd) { atomicMax(err, 1); mv.dead = 1; return; }
        group_add(g, r, P, E, B, tE, tB, nadd);                        // State::add, cpp/state.cpp:367-380
        if (hist) hist->added[nadd] = (uint16_t)pos;
        nadd++;
        const int qhi = min(m + k, H.n);
        for (int t = b0; t < H.boff[qhi]; t++) {
            if (!vget(valid, t - b0)) continue;
            const uint64_t Q = rel_row(H.rows + (size_t)t * H.W, H.W, o);
            bool com = true;
            for (int i = 0; i < r; i++) com &= !anti(g[i], Q);
            if (!com) vclr(valid, t - b0);
        }
    }
    mv.nadd = nadd;
    if (hist) hist->nadd = nadd;
    for (int pos = b0; pos < b1; pos++) {                               // cpp/state.cpp:405-410
        const uint64_t P = rel_row(H.rows + (size_t)pos * H.W, H.W, o);
        if (in_rref(g, r, P)) group_add_energy(g, r, P, H.w[pos], E);
    }
    group_project_site0(g, r, E, B);                                    // project_to(k-1), cpp/stab.cpp:803-809 (site o_m; no-op while o_m < 0)
    for (int i = 0; i < r; i++) mv.g[i] = g[i] >> 2;                    // origin o_{m+1}
    for (int i = r; i < kDpRows; i++) mv.g[i] = 0;
    mv.r = r;
    // valid bits relative to boff[m+1]; bucket m+k+1 (never filtered so far) enters as all-valid
    vshift_right(valid, b1 - b0);
    const int lo = H.boff[min(m + k + 1, H.n + 1)] - b1, hi = H.boff[min(m + k + 2, H.n + 1)] - b1;
    if (hi > 64 * kDpVW) { atomicMax(err, 2); mv.dead = 1; return; }
    vset_range(valid, lo, hi);
    for (int i = 0; i < kDpVW; i++) mv.valid[i] = valid[i];
}

// k_dp_move: one thread per state (grid-stride over ctl->nin)
__global__ void k_dp_move(DpTerms H, int m, const DpState *__restrict__ in, const double *__restrict__ inE, DpCtl *ctl,
                          int tstride, const DpGroup *__restrict__ level, DpMoved *__restrict__ out, double *slabE,
                          DpBack *slabB, int slab_stride, DpHistAdd *hm)
{
    const int nin = ctl->nin;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nin; s += gridDim.x * blockDim.x) {
        double *E = slabE + (size_t)s * 2 * slab_stride;
        DpBack *B = slabB + (size_t)s * 2 * slab_stride;
        DpHistAdd *h = nullptr;
        if (ctl->hm_off + s < ctl->hm_cap) h = hm + ctl->hm_off + s; else atomicMax(&ctl->err, 7);
        dp_move_one(H, m, s, in[s], inE, tstride, level, out[s], E, E + slab_stride, B, B + slab_stride, h, &ctl->err);
    }
}

// span of the valid terms with last site in [m+1, min(m+k+1, n)) truncated to [max(m-k+2,0), m+2):
// State::check_Sright_validity (cpp/state.cpp:347-365)
SGS_D void dp_valid_span(const DpTerms &H, int m, const DpMoved &mv, Ech &e)
{
    const int k = H.k, o1 = m - k + 2;                     // origin of level m+1
    const uint64_t wmask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    e.n = 0;
    const int b1 = H.boff[m + 1], t1 = H.boff[min(m + k + 1, H.n)];
    for (int t = b1; t < t1; t++) {
        if (!vget(mv.valid, t - b1)) continue;
        ech_add(e, rel_row(H.rows + (size_t)t * H.W, H.W, o1) & wmask);
    }
}
SGS_D bool dp_accepts(const Ech &e, const DpGroup &G)
{
    for (int i = 0; i < G.r; i++) if (ech_reduce(e, G.g[i]) != 0) return false;
    return true;
}

// k_dp_fanout: evolve_single's loop over compatible next right environments (cpp/state.cpp:467-479):
// nacc[s] = number of accepted candidates of state s
__global__ void k_dp_fanout(DpTerms H, int m, const DpMoved *__restrict__ mv, const DpCtl *ctl, const int *__restrict__ map_off,
                            const int *__restrict__ map_idx, const DpGroup *__restrict__ next_level, int *__restrict__ nacc)
{
    const int nin = ctl->nin;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nin; s += gridDim.x * blockDim.x) {
        int cnt = 0;
        if (!mv[s].dead) {
            Ech e;
            dp_valid_span(H, m, mv[s], e);
            const int sr = mv[s].sright;
            for (int j = map_off[sr]; j < map_off[sr + 1]; j++) cnt += dp_accepts(e, next_level[map_idx[j]]);
        }
        nacc[s] = cnt;
    }
}

// exclusive scan of nacc → child offsets; one block
__global__ void k_dp_scan(const int *__restrict__ nacc, DpCtl *ctl, int *__restrict__ off)
{
    __shared__ int part[1024];
    const int nin = ctl->nin, tid = threadIdx.x, nt = blockDim.x;
    const int per = (nin + nt - 1) / nt, lo = min(tid * per, nin), hi = min(lo + per, nin);
    int acc = 0;
    for (int i = lo; i < hi; i++) acc += nacc[i];
    part[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < nt; i++) { const int v = part[i]; part[i] = run; run += v; }
        ctl->total = run;
        if (run > ctl->capc) { ctl->err = max(ctl->err, 6); ctl->total = 0; }
        off[nin] = run;
    }
    __syncthreads();
    acc = part[tid];
    for (int i = lo; i < hi; i++) { off[i] = acc; acc += nacc[i]; }
}

// children in deterministic (parent, map order): key + reference to the moved state
__global__ void k_dp_emit(DpTerms H, int m, const DpMoved *__restrict__ mv, const DpCtl *ctl, const int *__restrict__ map_off,
                          const int *__restrict__ map_idx, const DpGroup *__restrict__ next_level, const int *__restrict__ off,
                          DpKey *__restrict__ keys, int *__restrict__ parent)
{
    const int nin = ctl->nin;
    if (ctl->total == 0) return;
    const int k = H.k, m1 = m + 1;
    // key: valid bits for last site in [m1+1, min(m1+k-1, n)), positions relative to boff[m1]
    const int lo = H.boff[min(m1 + 1, H.n + 1)] - H.boff[m1];
    const int hi = H.boff[min(max(m1 + k - 1, m1 + 1), H.n)] - H.boff[m1];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nin; s += gridDim.x * blockDim.x) {
        if (mv[s].dead) continue;
        Ech e;
        dp_valid_span(H, m, mv[s], e);
        int c = off[s];
        const int sr = mv[s].sright;
        for (int j = map_off[sr]; j < map_off[sr + 1]; j++) {
            if (!dp_accepts(e, next_level[map_idx[j]])) continue;
            DpKey &K = keys[c];
            for (int i = 0; i < kDpRows; i++) K.g[i] = mv[s].g[i];
            for (int i = 0; i < kDpVW; i++) K.v[i] = 0;
            for (int b = lo; b < hi; b++) if (vget(mv[s].valid, b)) K.v[b >> 6] |= 1ULL << (b & 63);
            K.r = mv[s].r;
            K.sright = map_idx[j];
            parent[c] = s;
            c++;
        }
    }
}

SGS_D int key_cmp(const DpKey &a, const DpKey &b)
{
    if (a.r != b.r) return a.r < b.r ? -1 : 1;
    if (a.sright != b.sright) return a.sright < b.sright ? -1 : 1;
    for (int i = 0; i < kDpRows; i++) if (a.g[i] != b.g[i]) return a.g[i] < b.g[i] ? -1 : 1;
    for (int i = 0; i < kDpVW; i++) if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
    return 0;
}

SGS_D int key_cmp(const DpGroup &a, const DpGroup &b)
{
    if (a.r != b.r) return a.r < b.r ? -1 : 1;
    for (int i = 0; i < kDpRows; i++) if (a.g[i] != b.g[i]) return a.g[i] < b.g[i] ? -1 : 1;
    for (int i = 0; i < kDpRows; i++) if (a.neg[i] != b.neg[i]) return a.neg[i] < b.neg[i] ? -1 : 1;
    return 0;
}

// leader[c] = first entry with the same key (the lists are a few hundred to a few thousand entries:
// all-pairs compare); the list length is read from the device
template <typename K>
__global__ void k_group_leader(const K *__restrict__ keys, const int *nc_dev, int *__restrict__ leader)
{
    const int nc = *nc_dev;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        int lead = c;
        for (int j = 0; j < c; j++) if (key_cmp(keys[j], keys[c]) == 0) { lead = j; break; }
        leader[c] = lead;
    }
}

// rank[c] (leaders only, else -1) = number of distinct smaller keys: the deterministic total order that
// replaces the reference's std::map-on-a-boost-hash iteration order
template <typename K>
__global__ void k_group_rank(const K *__restrict__ keys, const int *nc_dev, const int *__restrict__ leader, int *__restrict__ rank, int *nuniq)
{
    const int nc = *nc_dev;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        if (leader[c] != c) { rank[c] = -1; continue; }
        int rk = 0;
        for (int j = 0; j < nc; j++) if (leader[j] == j && key_cmp(keys[j], keys[c]) < 0) rk++;
        rank[c] = rk;
        atomicAdd(nuniq, 1);
    }
}

// merged states of site m+1, in key order: element-wise minimum over the children of a key, exact ties
// won by the later child (merge_gs keeps the newcomer unless the incumbent is strictly lower,
// cpp/stab.cpp:832).  One block per leader (grid-stride), one thread per table entry.
__global__ void k_dp_merge(const DpKey *__restrict__ keys, const int *__restrict__ parent, const int *__restrict__ leader,
                           const int *__restrict__ rank, DpCtl *ctl, const double *__restrict__ slabE, const DpBack *__restrict__ slabB,
                           int slab_stride, int tstride, DpState *__restrict__ outS, double *__restrict__ outE,
                           DpBack *__restrict__ hb, const DpMoved *__restrict__ mv)
{
    const int nc = ctl->total;
    if (ctl->nuniq > ctl->caps) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicMax(&ctl->err, 5); return; }
    const bool hist_ok = (long long)ctl->hb_off + (long long)ctl->nuniq * tstride <= (long long)ctl->hb_cap;
    if (!hist_ok && threadIdx.x == 0 && blockIdx.x == 0) atomicMax(&ctl->err, 7);
    for (int c = blockIdx.x; c < nc; c += gridDim.x) {
        if (leader[c] != c) continue;
        const int dst = rank[c];
        const int r = keys[c].r;
        for (int e = threadIdx.x; e < (1 << r); e += blockDim.x) {
            double best = 0.0;
            DpBack bb;
            bb.state = -1; bb.entry = 0; bb.signs = 0;
            bool have = false;
            for (int j = c; j < nc; j++) {
                if (leader[j] != c) continue;
                const int p = parent[j];
                const double v = slabE[(size_t)p * 2 * slab_stride + e];
                if (!have || !(best < v)) { best = v; bb = slabB[(size_t)p * 2 * slab_stride + e]; have = true; }
            }
            outE[(size_t)dst * tstride + e] = best;
            if (hist_ok) hb[(size_t)ctl->hb_off + (size_t)dst * tstride + e] = bb;
        }
        if (threadIdx.x == 0) {
            DpState &S = outS[dst];
            const int p = parent[c];
            for (int i = 0; i < kDpRows; i++) S.g[i] = keys[c].g[i];
            for (int i = 0; i < kDpVW; i++) S.valid[i] = mv[p].valid[i];
            S.r = r;
            S.sright = keys[c].sright;
        }
    }
}

// end of a site step: log it, advance the history offsets, the merged states become the next input
__global__ void k_dp_advance(DpCtl *ctl, DpStepLog *log, int m, int tstride)
{
    if (threadIdx.x || blockIdx.x) return;
    if (ctl->step < ctl->log_cap) {
        DpStepLog &L = log[ctl->step];
        L.m = m; L.nin = ctl->nin; L.total = ctl->total; L.nuniq = ctl->nuniq; L.hm_off = ctl->hm_off; L.hb_off = ctl->hb_off;
    } else ctl->err = max(ctl->err, 7);
    if (ctl->err) ctl->nuniq = 0;          // a failed step leaves no state: later steps become no-ops, the host sees err
    ctl->step++;
    ctl->hm_off += ctl->nin;
    ctl->hb_off += ctl->nuniq * tstride;
    ctl->nin = ctl->nuniq;
    ctl->total = 0;
    ctl->nuniq = 0;
}

// history of (state, entry) of the current site back to step `floor`: (term position, sign) pairs in the order
// the reference appends them (Ps_indices / Ps_signs, cpp/stab.cpp:752-759).  One thread.
__global__ void k_dp_trace(const DpCtl *ctl, const DpStepLog *log, const DpHistAdd *hm, const DpBack *hb, int tstride, int floor_step,
                           int state, int entry, int *out_pos, int *out_sign, int *out_n, int cap)
{
    if (threadIdx.x || blockIdx.x) return;
    int n = 0, s = state, e = entry;
    for (int t = ctl->step - 1; t >= floor_step; t--) {
        const DpStepLog &L = log[t];
        const DpBack b = hb[(size_t)L.hb_off + (size_t)s * tstride + e];
        if (b.state < 0) break;
        const DpHistAdd &h = hm[L.hm_off + b.state];
        for (int a = h.nadd - 1; a >= 0 && n < cap; a--) { out_pos[n] = h.added[a]; out_sign[n] = ((b.signs >> a) & 1) ? -1 : 1; n++; }
        s = b.state; e = b.entry;
    }
    for (int i = 0; i < n / 2; i++) {       // collected newest-first: reverse to chronological order
        int t = out_pos[i]; out_pos[i] = out_pos[n - 1 - i]; out_pos[n - 1 - i] = t;
        t = out_sign[i]; out_sign[i] = out_sign[n - 1 - i]; out_sign[n - 1 - i] = t;
    }
    *out_n = n;
}

// valid bits outside the key range may differ between the children of one key (the reference keeps the
// first-inserted state's: State::merge_gs copies state1, cpp/state.cpp:415-419); take the FIRST child's,
// as the sequential merge does.

// ------------------------------------------------------------------------------------------------
// Right environments.  k_sr_evolve: evolve_stab (cpp/state.cpp:147-185) for one parent group per thread:
// include/exclude branching over the terms whose last site is m, widen one site to the left, project onto
// the k sites ending at m (Stab::project, cpp/stab.cpp:524-577).  Children are written to out[parent *
// maxbranch + b] relative to origin o_m.
struct SrBranch { uint64_t g[kDpRows]; int q[kDpRows]; int r; uint32_t excl; };

SGS_D void signed_add(uint64_t *g, int *q, int &r, uint64_t P, int qP)     // Stab::add: RREF with signs
{
    uint64_t v = P;
    int qv = qP;
    for (int i = 0; i < r; i++) if ((v >> ctz64(g[i])) & 1) mul1(v, qv, g[i], q[i]);
    if (v == 0) return;
    const int pv = ctz64(v);
    for (int i = 0; i < r; i++) if ((g[i] >> pv) & 1) mul1(g[i], q[i], v, qv);
    int pos = r;
    while (pos > 0 && ctz64(g[pos - 1]) > pv) { g[pos] = g[pos - 1]; q[pos] = q[pos - 1]; pos--; }
    g[pos] = v; q[pos] = qv;
    r++;
}

__global__ void k_sr_evolve(DpTerms H, int m, const DpGroup *__restrict__ next, int nnext, int maxbranch,
                            DpGroup *__restrict__ out, int *__restrict__ nout, SrBranch *scratch, int *err)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnext) return;
    const int k = H.k, o = m - k + 1;                       // everything is done at the child origin o_m = o_{m+1} - 1
    SrBranch *br = scratch + (size_t)p * 2 * maxbranch, *nb = br + maxbranch;
    int nbr = 1;
    br[0].r = next[p].r; br[0].excl = 0;
    for (int i = 0; i < next[p].r; i++) {
        br[0].g[i] = next[p].g[i] << 2;
        br[0].q[i] = (q_plus(next[p].g[i]) + (next[p].neg[i] ? 2 : 0)) & 3;
    }
    const int b0 = H.boff[m], b1 = H.boff[m + 1];
    if (b1 - b0 > 30) { atomicMax(err, 3); nout[p] = 0; return; }
    for (int pos = b0; pos < b1; pos++) {
        const uint64_t P = rel_row(H.rows + (size_t)pos * H.W, H.W, o);
        int nn = 0;
        for (int b = 0; b < nbr; b++) {
            SrBranch &Bc = br[b];
            bool com = true;
            for (int i = 0; i < Bc.r; i++) com &= !anti(Bc.g[i], P);
            if (com && !in_rref(Bc.g, Bc.r, P)) {
                if (nn + 2 > maxbranch) { atomicMax(err, 4); nout[p] = 0; return; }
                SrBranch inc = Bc;                                      // branch_include (cpp/state.cpp:129-138)
                signed_add(inc.g, inc.q, inc.r, P, q_plus(P));
                bool ok = inc.r <= kDpRows;
                for (int t = b0; t < pos && ok; t++)
                    if ((Bc.excl >> (t - b0)) & 1)
                        ok = !in_rref(inc.g, inc.r, rel_row(H.rows + (size_t)t * H.W, H.W, o));
                if (ok) nb[nn++] = inc;
                Bc.excl |= 1u << (pos - b0);                            // branch_exclude (:140-145)
                nb[nn++] = Bc;
            } else {
                if (nn + 1 > maxbranch) { atomicMax(err, 4); nout[p] = 0; return; }
                nb[nn++] = Bc;
            }
        }
        SrBranch *t = br; br = nb; nb = t;
        nbr = nn;
    }
    // widen to [max(o,0), m+2), project onto [max(o,0), m+1): keep the subgroup with no support on site m+1
    // (relative to o_m site m+1 is position k).
    const int outpos = k;
    for (int b = 0; b < nbr; b++) {
        SrBranch &Bc = br[b];
        // eliminate on the two columns of the outside site, keep rows that are clear there
        uint64_t g[kDpRows]; int q[kDpRows]; int r = Bc.r;
        for (int i = 0; i < r; i++) { g[i] = Bc.g[i]; q[i] = Bc.q[i]; }
        int used = 0;
        for (int bit = 2 * outpos; bit < 2 * outpos + 2; bit++) {
            int pr = -1;
            for (int i = used; i < r; i++) if ((g[i] >> bit) & 1) { pr = i; break; }
            if (pr < 0) continue;
            uint64_t tg = g[pr]; g[pr] = g[used]; g[used] = tg;
            int tq = q[pr]; q[pr] = q[used]; q[used] = tq;
            for (int i = 0; i < r; i++) if (i != used && ((g[i] >> bit) & 1)) mul1(g[i], q[i], g[used], q[used]);
            used++;
        }
        DpGroup &G = out[(size_t)p * maxbranch + b];
        uint64_t cg[kDpRows]; int cq[kDpRows]; int cr = 0;
        for (int i = used; i < r; i++) signed_add(cg, cq, cr, g[i], q[i]);   // canonicalise the subgroup
        G.r = cr;
        for (int i = 0; i < kDpRows; i++) { G.g[i] = 0; G.neg[i] = 0; }
        for (int i = 0; i < cr; i++) {
            G.neg[i] = q_sign(cg[i], cq[i]) < 0;
            G.g[i] = cg[i];
        }
    }
    nout[p] = nbr;
}

// children of all parents, compacted in parent order: flat[], par[] and the per-parent offsets; one block
__global__ void k_sr_flatten(const DpGroup *__restrict__ strided, const int *__restrict__ nout, int nnext, int maxbranch,
                             DpGroup *__restrict__ flat, int *__restrict__ par, int *__restrict__ offp, int *total)
{
    if (threadIdx.x == 0) {
        int run = 0;
        for (int p = 0; p < nnext; p++) { offp[p] = run; run += nout[p]; }
        offp[nnext] = run;
        *total = run;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nnext; p += blockDim.x) {
        const int o = offp[p], c = offp[p + 1] - o;
        for (int j = 0; j < c; j++) { flat[o + j] = strided[(size_t)p * maxbranch + j]; par[o + j] = p; }
    }
}

// get_map_from_stabs (cpp/state.cpp:187-216) on the device: distinct groups in rank order and, per group, the
// ascending list of distinct parents (py/state.py:523 uses a set).  info = {nuniq, nlinks, maxfan}.  One block.
__global__ void k_sr_build(const DpGroup *__restrict__ flat, const int *__restrict__ par, const int *__restrict__ offp,
                           const int *__restrict__ leader, const int *__restrict__ rank, const int *total_dev, const int *nuniq_dev,
                           DpGroup *__restrict__ groups, int *__restrict__ map_off, int *__restrict__ map_idx, int *__restrict__ cnt, int *info)
{
    const int total = *total_dev, nuniq = *nuniq_dev;
    for (int g = threadIdx.x; g <= nuniq; g += blockDim.x) cnt[g] = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
        const int g = rank[leader[c]];
        if (leader[c] == c) groups[g] = flat[c];
        bool first = true;                       // first child of this parent in this group?
        for (int j = offp[par[c]]; j < c; j++) if (leader[j] == leader[c]) { first = false; break; }
        if (first) atomicAdd(&cnt[g], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, mf = 0;
        for (int g = 0; g < nuniq; g++) { map_off[g] = run; run += cnt[g]; mf = max(mf, cnt[g]); }
        map_off[nuniq] = run;
        info[0] = nuniq; info[1] = run; info[2] = mf;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < nuniq; g += blockDim.x) {
        int k = map_off[g];
        for (int c = 0; c < total; c++) {
            if (rank[leader[c]] != g) continue;
            bool first = true;
            for (int j = offp[par[c]]; j < c; j++) if (leader[j] == leader[c]) { first = false; break; }
            if (first) map_idx[k++] = par[c];
        }
    }
}

}  // namespace sgs
