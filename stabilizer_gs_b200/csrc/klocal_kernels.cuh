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
// k_dp_move: State::move_to_next + project_to(k-1) (cpp/state.cpp:382-413, :460-466), one thread per state.
// Output: the moved state (rows relative to o_{m+1}, valid bits relative to boff[m+1]), its table and
// back-pointers in slab `s`, the terms it added, dead flag.
struct DpMoved {
    uint64_t g[kDpRows];
    uint64_t valid[kDpVW];
    int r, sright, dead, nadd;
    int added[kDpMaxAdd];          // term positions added at this site, in order
};

__global__ void k_dp_move(DpTerms H, int m, const DpState *__restrict__ in, const double *__restrict__ inE, int nin,
                          int tstride, const DpGroup *__restrict__ level, DpMoved *__restrict__ out, double *slabE,
                          DpBack *slabB, int slab_stride, int *err)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nin) return;
    const int k = H.k, o = m - k + 1;
    const DpState &st = in[s];
    DpMoved &mv = out[s];
    double *E = slabE + (size_t)s * 2 * slab_stride, *tE = E + slab_stride;
    DpBack *B = slabB + (size_t)s * 2 * slab_stride, *tB = B + slab_stride;
    uint64_t g[kDpRows + 1];
    int r = st.r;
    for (int i = 0; i < r; i++) g[i] = st.g[i];
    uint64_t valid[kDpVW];
    for (int i = 0; i < kDpVW; i++) valid[i] = st.valid[i];
    for (int e = 0; e < (1 << r); e++) { E[e] = inE[(size_t)s * tstride + e]; B[e].state = s; B[e].entry = (uint16_t)e; B[e].signs = 0; }
    const DpGroup &sr = level[st.sright];
    mv.sright = st.sright; mv.dead = 0; mv.nadd = 0;

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
        mv.added[nadd++] = pos;
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

// ------------------------------------------------------------------------------------------------
// k_dp_fanout: evolve_single's loop over compatible next right environments (cpp/state.cpp:467-479) with
// State::check_Sright_validity (:347-365).  One thread per moved state; accept[s * maxfan + j] = 1 if the
// j-th candidate of its map list is valid.
__global__ void k_dp_fanout(DpTerms H, int m, const DpMoved *__restrict__ mv, int nin, const int *__restrict__ map_off,
                            const int *__restrict__ map_idx, const DpGroup *__restrict__ next_level, int maxfan,
                            unsigned char *__restrict__ accept, int *__restrict__ nacc)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nin) return;
    nacc[s] = 0;
    if (mv[s].dead) return;
    const int k = H.k, o1 = m - k + 2;                     // origin of level m+1
    // span of the valid terms with last site in [m+1, min(m+k+1, n)) truncated to [max(m-k+2,0), m+2)
    const uint64_t wmask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    Ech e; e.n = 0;
    const int b1 = H.boff[m + 1], t1 = H.boff[min(m + k + 1, H.n)];
    for (int t = b1; t < t1; t++) {
        if (!vget(mv[s].valid, t - b1)) continue;
        ech_add(e, rel_row(H.rows + (size_t)t * H.W, H.W, o1) & wmask);
    }
    int cnt = 0;
    const int sr = mv[s].sright;
    for (int j = map_off[sr]; j < map_off[sr + 1]; j++) {
        const DpGroup &G = next_level[map_idx[j]];
        bool ok = true;
        for (int i = 0; i < G.r && ok; i++) ok = ech_reduce(e, G.g[i]) == 0;
        accept[(size_t)s * maxfan + (j - map_off[sr])] = ok ? 1 : 0;
        cnt += ok;
    }
    nacc[s] = cnt;
}

// exclusive scan of nacc → child offsets; one thread (the lists are a few hundred entries)
__global__ void k_dp_scan(const int *__restrict__ nacc, int nin, int *__restrict__ off, int *total)
{
    if (threadIdx.x || blockIdx.x) return;
    int acc = 0;
    for (int i = 0; i < nin; i++) { off[i] = acc; acc += nacc[i]; }
    off[nin] = acc;
    *total = acc;
}

// children in deterministic (parent, map order): key + reference to the moved state
__global__ void k_dp_emit(DpTerms H, int m, const DpMoved *__restrict__ mv, int nin, const int *__restrict__ map_off,
                          const int *__restrict__ map_idx, int maxfan, const unsigned char *__restrict__ accept,
                          const int *__restrict__ off, DpKey *__restrict__ keys, int *__restrict__ parent)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nin || mv[s].dead) return;
    const int k = H.k, m1 = m + 1;
    // key: valid bits for last site in [m1+1, min(m1+k-1, n)), positions relative to boff[m1]
    const int lo = H.boff[min(m1 + 1, H.n + 1)] - H.boff[m1];
    const int hi = H.boff[min(max(m1 + k - 1, m1 + 1), H.n)] - H.boff[m1];
    int c = off[s];
    const int sr = mv[s].sright;
    for (int j = map_off[sr]; j < map_off[sr + 1]; j++) {
        if (!accept[(size_t)s * maxfan + (j - map_off[sr])]) continue;
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

// leader[c] = first entry with the same key (the lists are a few hundred entries: all-pairs compare)
template <typename K>
__global__ void k_group_leader(const K *__restrict__ keys, int nc, int *__restrict__ leader)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    int lead = c;
    for (int j = 0; j < c; j++) if (key_cmp(keys[j], keys[c]) == 0) { lead = j; break; }
    leader[c] = lead;
}

// rank[c] (leaders only, else -1) = number of distinct smaller keys: the deterministic total order that
// replaces the reference's std::map-on-a-boost-hash iteration order
template <typename K>
__global__ void k_group_rank(const K *__restrict__ keys, int nc, const int *__restrict__ leader, int *__restrict__ rank, int *nuniq)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    if (leader[c] != c) { rank[c] = -1; return; }
    int rk = 0;
    for (int j = 0; j < nc; j++) if (leader[j] == j && key_cmp(keys[j], keys[c]) < 0) rk++;
    rank[c] = rk;
    atomicAdd(nuniq, 1);
}

// merged states of site m+1, in key order: element-wise minimum over the children of a key, exact ties
// won by the later child (merge_gs keeps the newcomer unless the incumbent is strictly lower,
// cpp/stab.cpp:832).  One thread per (leader, table entry).
__global__ void k_dp_merge(const DpKey *__restrict__ keys, const int *__restrict__ parent, const int *__restrict__ leader,
                           const int *__restrict__ rank, int nc, const double *__restrict__ slabE, const DpBack *__restrict__ slabB,
                           int slab_stride, int tstride, DpState *__restrict__ outS, double *__restrict__ outE,
                           DpBack *__restrict__ outB, const DpMoved *__restrict__ mv)
{
    const int c = blockIdx.x;
    if (c >= nc || leader[c] != c) return;
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
        outB[(size_t)dst * tstride + e] = bb;
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

}  // namespace sgs
