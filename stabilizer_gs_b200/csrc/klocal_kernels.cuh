// klocal_kernels.cuh — sm_100a kernels of the 1D k-local dynamic programme.
//
// What they replace (reference): StabEnergies::{add, transform, flip_sign, add_energy,
// project_to_next, merge_gs} (cpp/stab.cpp:667-843), Stab::{canonicalize, decompose, include, commute, project}
// (cpp/stab.cpp:437-577), State::{add, move_to_next, check_Sright_validity, hash_value} (cpp/state.cpp:311-413),
// StateMachine::{evolve_single, evolve} (cpp/state.cpp:460-533) and evolve_stab / get_map_from_stabs
// (cpp/state.cpp:129-216).
//
// Design (not a port): every object a DP step touches lives on at most 2k sites, so all rows are ONE u64
// relative to the step's origin o_m = m - k + 1 (site s at bits 2(s-o), 2(s-o)+1).  The DP is latency-bound, so
// the whole of it runs in two persistent kernels (k_sr_run: right environments of every site, k_dp_run: any
// number of consecutive site steps) whose phases are separated by barriers and whose batch sizes live in a
// device control block — no host round trip and no kernel boundary inside the site loop.  Each kernel has
// three launch modes (StepSync: one CTA / one thread-block cluster / cooperative grid) chained on the stream,
// the small ones handing over on the device when a batch outgrows them.
// One warp owns one state (generator rows replicated per lane, the 2^rank sign table and its back-pointers
// split across the lanes, in shared memory or a per-state global slab); one thread owns one right-environment
// branch (one root-to-leaf path of the include/exclude tree).
// Instead of the reference's per-entry history arrays (2n ints per table entry, cpp/stab.cpp:752-759) every
// table entry keeps one back-pointer (parent state, parent entry, signs of the generators added at this site)
// in a history pool; the generator list is recovered by a device traceback at the end.
// The reference's keyed merge (std::map on a boost hash + omp critical, cpp/state.cpp:494-517) becomes a
// generation-tagged open-addressing table (claim by CAS, lowest entry per key by atomicMin, children of a key on a
// linked list), keys ranked in full-key order, and "later in (parent, child) order wins exact ties" = its strict '<'.
#pragma once
#include <cooperative_groups.h>

#include <cstdint>

#include "packed.cuh"

namespace sgs {

constexpr int kDpMaxK = 15;        // rows must fit 2k+2 sites in one u64; a group on k+1 sites has rank <= 16
constexpr int kDpRows = 16;        // max rank of a window group
constexpr int kDpTab = 10;         // max rank with a materialised 2^rank table
constexpr int kDpVW = 4;           // valid-term bitset words (terms in k+2 consecutive buckets <= 256)
constexpr int kDpMaxAdd = 16;      // generators added at one site
constexpr int kDpMaxSr = 2 * kDpRows;
static_assert(kDpMaxSr <= 32, "one lane per echelon row");

struct DpTerms {                   // Hamiltonian in paulis_list order
    int n, k, T, W;
    const uint64_t *rows;          // [T] compact rows: bit 2j = z, 2j+1 = x of site tb[t] + j (a term spans <= k <= 15 sites)
    const int *tb;                 // [T] first site of each term's window
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

// ---- a term as a 64-bit row relative to origin o: bit 2j = z, 2j+1 = x of site o + j (bits outside are dropped).
// Every object of a site step lives on <= 2k sites, so the chain length n never enters: the DP runs at any n
// (cpp/state.cpp:411, cpp/stab.cpp:782-809 carry a k-site window the same way).
SGS_D uint64_t term_rel(const DpTerms &H, int pos, int o)
{
    const int sh = 2 * (H.tb[pos] - o);
    const uint64_t r = H.rows[pos];
    if (sh >= 64 || sh <= -64) return 0;
    return sh >= 0 ? (r << sh) : (r >> (-sh));
}

SGS_D int q_plus(uint64_t r) { return (4 - (popc64(r & (r >> 1) & kEven) & 3)) & 3; }
SGS_D int q_sign(uint64_t r, int q) { return ((q + popc64(r & (r >> 1) & kEven)) & 3) == 0 ? 1 : -1; }
SGS_D bool anti(uint64_t a, uint64_t b) { return parity64(a & swapzx(b)); }
SGS_D void mul1(uint64_t &a, int &qa, uint64_t b, int qb)
{
    qa = (qa + qb + 2 * parity64((a >> 1) & b & kEven)) & 3;
    a ^= b;
}

// membership of v in the span of RREF rows g[0..r) (bits only)
SGS_D bool in_rref(const uint64_t *g, int r, uint64_t v)
{
    for (int i = 0; i < r; i++) if (v & (g[i] & (0 - g[i]))) v ^= g[i];      // pivot = lowest set bit
    return v == 0;
}

// echelon basis (arbitrary order, each row clear at earlier pivots) — bits only
struct Ech { uint64_t b[kDpMaxSr]; uint64_t p[kDpMaxSr]; int n; };     // p[i] = lowest set bit of b[i] (its pivot)
SGS_D uint64_t ech_reduce(const Ech &e, uint64_t v)
{
    for (int i = 0; i < e.n; i++) if (v & e.p[i]) v ^= e.b[i];
    return v;
}
SGS_D void ech_add(Ech &e, uint64_t v)
{
    v = ech_reduce(e, v);
    if (v && e.n < kDpMaxSr) { e.b[e.n] = v; e.p[e.n] = v & (0 - v); e.n++; }
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
// Warp-cooperative group operations.  One warp owns one state: every lane holds identical copies of the
// generator rows (uniform control flow), the 2^rank sign table and its back-pointers are split across the
// lanes.
//
// canonical insertion of one '+' generator into a canonical all-'+' group with its sign table
// (StabEnergies::add, cpp/stab.cpp:747-761 = concat_last + canonicalize :693-697 = transform :699-733 +
// canonicalize_signs :685-691).  E/bk are the table and back-pointers (2^r entries in, 2^(r+1) out), tmpE/
// tmpB scratch of the same size; `abit` is the sign bit this addition occupies in DpBack::signs.
SGS_D void group_add(uint64_t *g, int &r, uint64_t P, double *E, DpBack *bk, double *tmpE, DpBack *tmpB, int abit, int lane)
{
    // 1. new generator enters as the LAST axis: new = 2*old + b, b = 1 <=> '-'
    const int sz = 1 << r;
    for (int o = lane; o < sz; o += 32) {
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
    __syncwarp();
    const int nsz = 1 << r;
    for (int sv = lane; sv < nsz; sv += 32) {
        const uint32_t smask = __brev((uint32_t)sv) >> (32 - r);        // bit j of smask = bit r-1-j of sv
        uint32_t ns = 0;
        for (int i = 0; i < r; i++) if (__popc(U[i] & smask) & 1) ns |= 1u << (r - 1 - i);
        ns ^= flip;
        E[ns] = tmpE[sv];
        bk[ns] = tmpB[sv];
    }
    __syncwarp();
}

// StabEnergies::add_energy (cpp/stab.cpp:819-825) for a member term of a canonical all-'+' group
SGS_D void group_add_energy(const uint64_t *g, int r, uint64_t P, double w, double *E, int lane)
{
    uint64_t acc = P;
    int qa = q_plus(P);
    uint32_t xi = 0;
    for (int i = 0; i < r; i++)
        if ((acc >> ctz64(g[i])) & 1) { mul1(acc, qa, g[i], q_plus(g[i])); xi |= 1u << (r - 1 - i); }
    const double ws = (qa & 2) ? -w : w;            // w * sign
    const int sz = 1 << r;
    for (int o = lane; o < sz; o += 32) E[o] += (__popc(o & xi) & 1) ? -ws : ws;
}

// StabEnergies::project_to_next (cpp/stab.cpp:782-801) for relative site 0
SGS_D void group_project_site0(uint64_t *g, int &r, double *E, DpBack *bk, int lane)
{
    __syncwarp();
    while (r > 0) {
        bool any = false;
        for (int i = 0; i < r; i++) any |= (g[i] & 3ULL) != 0;
        if (!any) break;
        const int half = 1 << (r - 1);
        for (int o = lane; o < half; o += 32)
            if (!(E[o] < E[o + half])) { E[o] = E[o + half]; bk[o] = bk[o + half]; }   // strict '<': a tie keeps '-'
        for (int i = 0; i + 1 < r; i++) g[i] = g[i + 1];
        r--;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Device-resident control block of the DP.  The whole DP (all sites) is ONE persistent cooperative kernel:
// batch sizes, history offsets and errors live here, the host reads it once at the end.
struct DpCtl {
    int nin;                 // states entering the next step
    int total;               // children emitted by the step in flight
    int nuniq;               // merged states of the step in flight
    int err;                 // 0 ok, 1 rank/table limit, 2 valid-set width, 5 state capacity, 6 child capacity, 7 history capacity
    int err_snap;            // err as published by block 0 between two grid barriers (uniform early exit)
    int hm_off, hb_off;      // bump offsets into the history pools
    int step;                // steps logged so far
    int cur;                 // which of the two state buffers holds the current states
    unsigned gen;            // generation of the dedup table (monotone over the machine's life)
    int caps, capc, hm_cap, hb_cap, log_cap;
    int adv_done;            // sites of the advance in flight that are done (hand-over from the cluster-mode kernel)
    int bailed;              // bit MODE: a small-mode kernel met a batch too large for it (sticky; the host then stops trying that mode)
    unsigned long long prof[8];   // ns spent per phase (thread 0's view), for SGS_DP_TRACE
};
struct DpStepLog { int m, nin, total, nuniq, hm_off, hb_off; };
struct DpHistAdd { int nadd; uint16_t added[kDpMaxAdd]; };      // what one parent state added at one site

struct DpLevelDev {          // type_Sright (cpp/state.hpp:49-51) of one site: groups + compatible next groups
    const DpGroup *groups;
    const int *map_off;      // [count + 1]
    const int *map_idx;      // [nlinks] indices into the level of site m+1
    int count, nlinks;
};

struct DpMoved {
    uint64_t g[kDpRows];
    uint64_t valid[kDpVW];
    uint64_t accm;           // bit i: the i-th compatible next environment (map order) accepts the state (if there are <= 64)
    int r, sright, dead, nadd;
};

// per-warp shared-memory staging of the move phase (tables of up to kMoveTab entries, i.e. k <= 4)
constexpr int kMoveTab = 16;
struct MoveStage { double E[2 * kMoveTab]; DpBack B[2 * kMoveTab]; DpState st; DpGroup sr; };
static_assert(sizeof(DpState) % 8 == 0 && sizeof(DpGroup) % 8 == 0 && sizeof(DpState) / 8 <= 32 && sizeof(DpGroup) / 8 <= 32, "one lane per word");

// State::move_to_next + project_to(k-1) (cpp/state.cpp:382-413, :460-466) for one state, by one warp.
// Returns the rank of the moved state (-1: dead) and its valid set to every lane; st / sr may live in shared
// memory, E/tE/B/tB (the 2^rank table, its back-pointers and scratch of the same size) too.
SGS_D int dp_move_warp(const DpTerms &H, int m, int s, const DpState &st, const double *inE, int tstride,
                       const DpGroup &sr, DpMoved &mv, double *E, double *tE, DpBack *B, DpBack *tB, DpHistAdd *hist, int *err, int lane,
                       const uint64_t *relw /* shared: rows of the terms from position boff[m] on, relative to o_m */, uint64_t *valid)
{
    const int k = H.k, o = m - k + 1;
    uint64_t g[kDpRows + 1];
    int r = st.r;
    for (int i = 0; i < r; i++) g[i] = st.g[i];
    for (int i = 0; i < kDpVW; i++) valid[i] = st.valid[i];
    for (int e = lane; e < (1 << r); e += 32) {
        E[e] = inE[(size_t)s * tstride + e];
        DpBack b; b.state = s; b.entry = (uint16_t)e; b.signs = 0;
        B[e] = b;
    }
    __syncwarp();
    if (lane == 0) { mv.sright = st.sright; mv.dead = 0; mv.nadd = 0; mv.r = 0; hist->nadd = 0; }

    // stab_tmp = <stab, Sright> (cpp/state.cpp:384-390), bits only
    Ech tmp; tmp.n = 0;
    for (int i = 0; i < r; i++) ech_add(tmp, g[i]);
    for (int i = 0; i < sr.r; i++) ech_add(tmp, sr.g[i]);
    const int b0 = H.boff[m], b1 = H.boff[m + 1];
    int nadd = 0;
    for (int pos = b0; pos < b1; pos++) {
        const uint64_t P = relw[pos - b0];
        if (ech_reduce(tmp, P) != 0) continue;
        if (!in_rref(sr.g, sr.r, P)) { if (lane == 0) mv.dead = 1; return -1; }       // cpp/state.cpp:397-399
        if (in_rref(g, r, P)) continue;
        if (r + 1 > kDpTab || r + 1 > kDpRows || nadd >= kDpMaxAdd) { if (lane == 0) { atomicMax(err, 1); mv.dead = 1; } return -1; }
        group_add(g, r, P, E, B, tE, tB, nadd, lane);                  // State::add, cpp/state.cpp:367-380
        if (lane == 0) hist->added[nadd] = (uint16_t)pos;
        nadd++;
        // terms that stopped commuting with the group leave the valid set: lanes stride over the terms
        const int nt = H.boff[min(m + k, H.n)] - b0;
        for (int base = 0; base < nt; base += 32) {
            const int t = base + lane;
            bool kill = false;
            if (t < nt && vget(valid, t)) {
                const uint64_t Q = relw[t];
                for (int i = 0; i < r; i++) kill |= anti(g[i], Q);
            }
            const unsigned km = __ballot_sync(0xffffffffu, kill);
            valid[base >> 6] &= ~((uint64_t)km << (base & 63));
        }
    }
    if (lane == 0) { mv.nadd = nadd; hist->nadd = nadd; }
    for (int pos = b0; pos < b1; pos++) {                               // cpp/state.cpp:405-410
        const uint64_t P = relw[pos - b0];
        if (in_rref(g, r, P)) group_add_energy(g, r, P, H.w[pos], E, lane);
    }
    group_project_site0(g, r, E, B, lane);                              // project_to(k-1), cpp/stab.cpp:803-809 (site o_m; no-op while o_m < 0)
    // valid bits relative to boff[m+1]; bucket m+k+1 (never filtered so far) enters as all-valid
    vshift_right(valid, b1 - b0);
    const int lo = H.boff[min(m + k + 1, H.n + 1)] - b1, hi = H.boff[min(m + k + 2, H.n + 1)] - b1;
    if (hi > 64 * kDpVW) { if (lane == 0) { atomicMax(err, 2); mv.dead = 1; } return -1; }
    vset_range(valid, lo, hi);
    if (lane == 0) {
        for (int i = 0; i < r; i++) mv.g[i] = g[i] >> 2;                // origin o_{m+1}
        for (int i = r; i < kDpRows; i++) mv.g[i] = 0;
        mv.r = r;
        for (int i = 0; i < kDpVW; i++) mv.valid[i] = valid[i];
    }
    return r;
}

// span of the valid terms with last site in [m+1, min(m+k+1, n)) truncated to [max(m-k+2,0), m+2):
// State::check_Sright_validity (cpp/state.cpp:347-365)
SGS_D void dp_valid_span(const DpTerms &H, int m, const uint64_t *valid, Ech &e, const uint64_t *relw)
{
    const int k = H.k, b0 = H.boff[m];
    const uint64_t wmask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    e.n = 0;
    const int b1 = H.boff[m + 1], t1 = H.boff[min(m + k + 1, H.n)];
    for (int t = b1; t < t1; t++) {
        if (!vget(valid, t - b1)) continue;
        ech_add(e, (relw[t - b0] >> 2) & wmask);             // origin o_{m+1} = o_m + 1; no window term reaches site o_m + 32
    }
}
SGS_D bool dp_accepts(const Ech &e, const DpGroup &G)
{
    for (int i = 0; i < G.r; i++) if (ech_reduce(e, G.g[i]) != 0) return false;
    return true;
}

SGS_D int key_cmp(const DpKey &a, const DpKey &b)
{
    if (a.r != b.r) return a.r < b.r ? -1 : 1;
    if (a.sright != b.sright) return a.sright < b.sright ? -1 : 1;
    for (int i = 0; i < a.r; i++) if (a.g[i] != b.g[i]) return a.g[i] < b.g[i] ? -1 : 1;     // rows >= r are zero in every key
    for (int i = 0; i < kDpVW; i++) if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
    return 0;
}

SGS_D int key_cmp(const DpGroup &a, const DpGroup &b)
{
    if (a.r != b.r) return a.r < b.r ? -1 : 1;
    for (int i = 0; i < a.r; i++) if (a.g[i] != b.g[i]) return a.g[i] < b.g[i] ? -1 : 1;     // rows and signs >= r are zero in every group
    for (int i = 0; i < a.r; i++) if (a.neg[i] != b.neg[i]) return a.neg[i] < b.neg[i] ? -1 : 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Right environments: evolve_stab (cpp/state.cpp:147-185) = include/exclude branching over the terms whose last
// site is m, widen one site to the left, project onto the k sites ending at m (Stab::project,
// cpp/stab.cpp:524-577).  A branch carries the group at origin o_m, the exclusion mask and its parent group.
struct SrBranch { uint64_t g[kDpRows]; int q[kDpRows]; int r; uint32_t excl; int par; };

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

// evolve_stab (cpp/state.cpp:147-185) as breadth-first waves over ALL branches of a level at once, one
// thread per branch.  Everything is done at the child origin o_m = o_{m+1} - 1.
// wave 0: the groups of level m+1 become the initial branches
SGS_D void sr_seed(const DpGroup &G, int p, SrBranch &b)
{
    b.r = G.r; b.excl = 0; b.par = p;
    for (int i = 0; i < G.r; i++) {
        b.g[i] = G.g[i] << 2;
        b.q[i] = (q_plus(G.g[i]) + (G.neg[i] ? 2 : 0)) & 3;
    }
}
// one term (position pos, bit `tbit` of the exclusion mask): branch_include / branch_exclude
// (cpp/state.cpp:129-145).  Returns the number of branches written to out[0..2).
SGS_D int sr_branch(const DpTerms &H, int o, int b0, int pos, const SrBranch &Bc, SrBranch *out)
{
    const uint64_t P = term_rel(H, pos, o);
    bool com = true;
    for (int i = 0; i < Bc.r; i++) com &= !anti(Bc.g[i], P);
    if (!com || in_rref(Bc.g, Bc.r, P)) { out[0] = Bc; return 1; }
    int nn = 0;
    SrBranch inc = Bc;
    signed_add(inc.g, inc.q, inc.r, P, q_plus(P));
    bool ok = inc.r <= kDpRows;
    for (int t = b0; t < pos && ok; t++)
        if ((Bc.excl >> (t - b0)) & 1) ok = !in_rref(inc.g, inc.r, term_rel(H, t, o));
    if (ok) out[nn++] = inc;
    out[nn] = Bc;
    out[nn].excl |= 1u << (pos - b0);
    return nn + 1;
}
// The same branching seen from ONE root-to-leaf path of the include/exclude tree (thread-per-path mode of
// k_sr_run): `bit` = 1 asks for the include branch of term j, 0 for the exclude branch; a term that does not
// branch (anticommutes or is already a member) continues on the 0 path only.  Returns false if the path
// does not exist.  srel[] = the site's term rows relative to o.
SGS_D bool sr_step(const uint64_t *srel, int j, int bit, SrBranch &b)
{
    const uint64_t P = srel[j];
    bool com = true;
    for (int i = 0; i < b.r; i++) com &= !anti(b.g[i], P);
    if (!com || in_rref(b.g, b.r, P)) return bit == 0;
    if (!bit) { b.excl |= 1u << j; return true; }
    if (b.r >= kDpRows) return false;
    signed_add(b.g, b.q, b.r, P, q_plus(P));
    for (int t = 0; t < j; t++)
        if (((b.excl >> t) & 1) && in_rref(b.g, b.r, srel[t])) return false;
    return true;
}
// last wave: widen to [max(o,0), m+2), project onto [max(o,0), m+1) — keep the subgroup with no support on
// site m+1 (position k relative to o_m): Stab::project (cpp/stab.cpp:524-577), canonical with signs
SGS_D void sr_project(const SrBranch &Bc, int k, DpGroup &G)
{
    uint64_t g[kDpRows]; int q[kDpRows]; int r = Bc.r;
    for (int i = 0; i < r; i++) { g[i] = Bc.g[i]; q[i] = Bc.q[i]; }
    int used = 0;
    for (int bit = 2 * k; bit < 2 * k + 2; bit++) {
        int pr = -1;
        for (int i = used; i < r; i++) if ((g[i] >> bit) & 1) { pr = i; break; }
        if (pr < 0) continue;
        uint64_t tg = g[pr]; g[pr] = g[used]; g[used] = tg;
        int tq = q[pr]; q[pr] = q[used]; q[used] = tq;
        for (int i = 0; i < r; i++) if (i != used && ((g[i] >> bit) & 1)) mul1(g[i], q[i], g[used], q[used]);
        used++;
    }
    uint64_t cg[kDpRows]; int cq[kDpRows]; int cr = 0;
    for (int i = used; i < r; i++) signed_add(cg, cq, cr, g[i], q[i]);   // canonicalise the subgroup
    G.r = cr;
    for (int i = 0; i < kDpRows; i++) { G.g[i] = 0; G.neg[i] = 0; }
    for (int i = 0; i < cr; i++) {
        G.neg[i] = q_sign(cg[i], cq[i]) < 0;
        G.g[i] = cg[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Keyed de-duplication shared by the DP merge (keys = DpKey) and the right-environment levels (keys =
// DpGroup).  An open-addressing table of generation-tagged 64-bit words replaces the reference's std::map
// on a boost hash (cpp/state.cpp:187-216, :494-517): no clearing between uses (a word whose tag is not the
// current generation is empty), every entry finds the slot of its key, the slot's `lead` becomes the lowest
// entry index with that key (deterministic) and `head` a linked list of all entries with the key.
struct DdSlot { unsigned long long claim, lead, head; };

struct DdView {
    uint64_t *pre;               // per entry: order prefix
    int *owner, *next, *leader;  // per entry: slot, next entry with the same key, lowest entry with the same key
    int *rank;                   // per leader entry: number of distinct smaller keys
    int *lidx;                   // compacted leader entries (arbitrary order)
    uint64_t *lpre;              // their prefixes
    DdSlot *tab;
    unsigned mask;               // table size - 1
};

SGS_D unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SGS_PROF(ctl, i) do { if (tid == 0) { const unsigned long long t__ = gtime(); (ctl)->prof[i] += t__ - tprof; tprof = t__; } } while (0)

SGS_D uint64_t mix64(uint64_t h, uint64_t v)
{
    h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ULL;
    return h ^ (h >> 31);
}
SGS_D uint64_t key_hash(const DpKey &k)
{
    uint64_t h = mix64((uint64_t)k.r << 32 | (uint32_t)k.sright, 0x1234567ULL);
    for (int i = 0; i < k.r; i++) h = mix64(h, k.g[i]);
    for (int i = 0; i < kDpVW; i++) h = mix64(h, k.v[i]);
    return h;
}
SGS_D uint64_t key_hash(const DpGroup &k)
{
    uint64_t h = mix64((uint64_t)k.r, 0x7654321ULL);
    for (int i = 0; i < k.r; i++) h = mix64(h, k.g[i] ^ ((uint64_t)k.neg[i] << 63));
    return h;
}
// order prefixes: pre(a) < pre(b) implies key_cmp(a, b) < 0; equal prefixes decide nothing.  Never all-ones.
// The prefix packs the leading fields of the key's lexicographic order into 64 bits.  A field of `width` bits
// holds min(value, 2^width - 1); a field that reaches its maximum closes the prefix (every later bit is zero),
// so the packing is monotone for ANY choice of widths - the widths in PrefixSpec only decide how many ties are
// left to the full-key comparison.
struct PrefixSpec {
    int srw;                     // bits for the right-environment index
    int wb;                      // bits per generator row
    int vlo, vw;                 // valid-term field: (v[0] >> vlo), vw bits (bits of v[0] below vlo are zero)
};
struct PrefixPacker {
    uint64_t acc;
    int left;
    bool open;
    SGS_D void add(uint64_t value, int width)
    {
        if (!open) return;
        width = min(width, left);
        if (width <= 0) { open = false; return; }
        const uint64_t maxv = width >= 64 ? ~0ULL : (1ULL << width) - 1;
        const uint64_t f = value >= maxv ? maxv : value;
        acc = width >= 64 ? f : (acc << width) | f;
        left -= width;
        if (f == maxv) open = false;
    }
    SGS_D uint64_t done() const { return left > 0 ? acc << left : acc; }
};
SGS_D uint64_t key_prefix(const DpKey &k, const PrefixSpec &ps)
{
    PrefixPacker P{(uint64_t)k.r, 59, true};                 // r <= kDpRows = 16 < 31: the top field never saturates
    P.add((uint64_t)(uint32_t)k.sright, ps.srw);
    for (int i = 0; i < k.r; i++) P.add(k.g[i], ps.wb);     // rows i >= r are zero in every key of this rank
    P.add(k.v[0] >> ps.vlo, ps.vw);
    return P.done();
}
SGS_D uint64_t key_prefix(const DpGroup &k, const PrefixSpec &ps)
{
    PrefixPacker P{(uint64_t)k.r, 59, true};
    for (int i = 0; i < k.r; i++) P.add(k.g[i], ps.wb);
    uint64_t sg = 0;                                         // neg[0] is the most significant sign
    for (int i = 0; i < k.r; i++) sg = (sg << 1) | (k.neg[i] ? 1 : 0);
    P.add(sg, k.r);
    return P.done();
}

// entry c looks up / claims the slot of its key, then joins the slot's minimum and list
template <typename K>
SGS_D void dd_insert(const DdView &D, unsigned gen, const K *keys, int c, const PrefixSpec &ps)
{
    const uint64_t h = key_hash(keys[c]);
    D.pre[c] = key_prefix(keys[c], ps);
    const unsigned long long tw = 0xFFFFFFFFu - gen, tagv = tw << 32;
    unsigned i = (unsigned)(h ^ (h >> 32)) & D.mask;
    for (;;) {
        unsigned long long cur = *(volatile unsigned long long *)&D.tab[i].claim;
        bool mine = false;
        while ((cur >> 32) != tw) {
            const unsigned long long prev = atomicCAS(&D.tab[i].claim, cur, tagv | (unsigned)c);
            if (prev == cur) { mine = true; break; }
            cur = prev;
        }
        if (mine) break;
        if (key_cmp(keys[(int)(unsigned)cur], keys[c]) == 0) break;
        i = (i + 1) & D.mask;
    }
    D.owner[c] = (int)i;
    atomicMin(&D.tab[i].lead, tagv | (unsigned)c);
    const unsigned long long old = atomicExch(&D.tab[i].head, tagv | (unsigned)c);
    D.next[c] = ((old >> 32) == tw) ? (int)(unsigned)old : -1;
}

// set of integer pairs on the same table (a fresh generation): true for exactly one entry per distinct
// (a[c], b[c]); which one is not significant
SGS_D bool dd_pair_first(const DdView &D, unsigned gen, const int *a, const int *b, int c)
{
    const unsigned long long tw = 0xFFFFFFFFu - gen, tagv = tw << 32;
    const uint64_t h = mix64(((uint64_t)(uint32_t)a[c] << 32) | (uint32_t)b[c], 0xABCDEFULL);
    unsigned i = (unsigned)(h ^ (h >> 32)) & D.mask;
    for (;;) {
        unsigned long long cur = *(volatile unsigned long long *)&D.tab[i].claim;
        while ((cur >> 32) != tw) {
            const unsigned long long prev = atomicCAS(&D.tab[i].claim, cur, tagv | (unsigned)c);
            if (prev == cur) return true;
            cur = prev;
        }
        const int j = (int)(unsigned)cur;
        if (a[j] == a[c] && b[j] == b[c]) return false;
        i = (i + 1) & D.mask;
    }
}

// leader of every entry; leaders are compacted (order irrelevant) for the all-pairs ranking
SGS_D void dd_lead(const DdView &D, int c, int *nuniq)
{
    const int ld = (int)(unsigned)D.tab[D.owner[c]].lead;
    D.leader[c] = ld;
    if (ld == c) {
        D.rank[c] = 0;
        const int p = atomicAdd(nuniq, 1);
        D.lidx[p] = c;
        D.lpre[p] = D.pre[c];
    }
}

// rank of every compacted leader = number of distinct smaller keys.  This is the deterministic total order that
// replaces the reference's iteration order of a std::map on a boost hash.  All-pairs, tiled: a warp
// task is (32 leaders, one per lane) x (a chunk of kRankChunk prefixes); the chunk is loaded coalesced, 32
// prefixes at a time, and broadcast lane by lane with shuffles, so every prefix is read from L2 once per 32
// leaders instead of once per leader.  Partial counts are added to rank[] (zeroed by dd_lead).  A padding
// prefix is all-ones, which key_prefix never produces.
constexpr int kRankChunk = 64;
template <typename K>
__device__ __noinline__ int dd_rank_ties(const DdView &D, const K *keys, int c, int qb, unsigned ties)
{
    int cnt = 0;
    for (; ties; ties &= ties - 1) cnt += key_cmp(keys[D.lidx[qb + __ffs(ties) - 1]], keys[c]) < 0;
    return cnt;
}
template <typename K>
SGS_D void dd_rank_tiles(const DdView &D, const K *keys, int nuniq, int wid, int nw, int lane, int slice = 0, int nslices = 1)
{
    const int ntl = (nuniq + 31) >> 5, nqc = (nuniq + kRankChunk - 1) / kRankChunk;
    for (long long task = (long long)wid * nslices + slice; task < (long long)ntl * nqc; task += (long long)nw * nslices) {
        const int p = (int)(task % ntl) * 32 + lane, q0 = (int)(task / ntl) * kRankChunk, q1 = min(q0 + kRankChunk, nuniq);
        const bool valid = p < nuniq;
        const uint64_t pc = valid ? D.lpre[p] : 0;                  // an invalid lane counts nothing that is used
        int cnt = 0;
        for (int qb = q0; qb < q1; qb += 32) {
            const uint64_t mine = qb + lane < q1 ? D.lpre[qb + lane] : ~0ULL;
            unsigned ties = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint64_t pq = __shfl_sync(0xffffffffu, mine, j);
                cnt += pq < pc;
                ties |= (unsigned)(pq == pc) << j;
            }
            if (qb == (p & ~31)) ties &= ~(1u << (p & 31));         // the leader itself
            if (valid && ties) cnt += dd_rank_ties(D, keys, D.lidx[p], qb, ties);
        }
        if (valid && cnt) atomicAdd(&D.rank[D.lidx[p]], cnt);
    }
}

// exclusive scan of v[0..n) into off[0..n] by one block; returns the total to every thread of the block
SGS_D int block_scan(const int *v, int *off, int n, int *part /* shared, blockDim.x ints */)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt, lo = min(tid * per, n), hi = min(lo + per, n);
    int acc = 0;
    for (int i = lo; i < hi; i++) acc += v[i];
    part[tid] = acc;
    __syncthreads();
    if (tid < 32) {                                          // warp 0: exclusive scan of the partial sums
        const int np = per > 0 ? (n + per - 1) / per : 0;    // threads that own elements (the others hold 0 and use nothing)
        int run = 0;
        for (int base = 0; base < np; base += 32) {
            const int t = base + tid < nt ? part[base + tid] : 0;
            int inc = t;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += y; }
            if (base + tid < nt) part[base + tid] = run + inc - t;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) off[n] = run;
    }
    __syncthreads();
    acc = part[tid];
    for (int i = lo; i < hi; i++) { const int t = v[i]; off[i] = acc; acc += t; }
    __syncthreads();
    return off[n];
}

// ------------------------------------------------------------------------------------------------
// Barrier between the phases of the persistent kernels, by launch mode:
//   kGrid    a cooperative launch of one 256-thread CTA per SM, cooperative_groups grid.sync() (1.5-2 us on B200
//            whatever the grid size) - any batch size;
//   kCluster the whole kernel is ONE thread-block cluster, phases separated by the hardware cluster barrier;
//   kBlock   the whole kernel is ONE CTA of 1024 threads, phases separated by __syncthreads(), and every table
//            of the site stays in that SM's L1.
// The DP is latency-bound: while a site's batch is small a phase costs a few dependent L2 round
// trips plus the barrier, so the small modes pay.  A small-mode kernel stops in front of the first site (level)
// whose batch is too large for its few warps and leaves the rest of the run to the next kernel enqueued right
// behind it (`resume`): block -> cluster -> grid, no host round trip in between.
enum { kGrid = 0, kCluster = 1, kBlock = 2 };
template <int MODE>
struct StepSync {
    static constexpr int kThreads = MODE == kBlock ? 1024 : 256;
    static constexpr int kBailStates = MODE == kBlock ? 1 : 2;     // bail above this many states per warp (the move phase is one warp per state) ...
    static constexpr int kBailChildren = MODE == kBlock ? 4 : 2;   // ... or children (branches) per thread
    SGS_D void sync() const
    {
        if constexpr (MODE == kCluster) cooperative_groups::this_cluster().sync();
        else if constexpr (MODE == kBlock) __syncthreads();
        else cooperative_groups::this_grid().sync();
    }
};

// ------------------------------------------------------------------------------------------------
// k_dp_run: StateMachine::evolve (cpp/state.cpp:483-533) for `nsteps` consecutive sites in one persistent
// cooperative launch.  Phases of one site, separated by grid barriers:
//   P0 move every state to the next site (one thread per state) and count its accepted next environments
//   P1 block 0: exclusive scan -> child offsets, capacity checks, error snapshot
//   P2 emit the children (key + parent) in deterministic (parent, map) order
//   P3 key de-duplication   P4 leaders   P5 key order of the leaders (tiled all-pairs)
//   P6 merge: element-wise table minimum over the children of a key (warp per key), back-pointers to the
//      history pool, merged states in key order
struct DpBufs {
    DpState *S[2];
    double *E[2];
    double *slabE;
    DpBack *slabB;
    DpMoved *moved;
    Ech *span;                   // per moved state: span of its valid terms (check_Sright_validity)
    int *nacc, *off;
    DpKey *keys;
    int *parent;
    DdView D;
    DpHistAdd *hm;
    DpBack *hb;
    DpCtl *ctl;
    DpStepLog *log;
    const DpLevelDev *levels;
    int tstride, slab_stride;
};

// phase_part (multi-GPU DP, grid mode, one site per launch): 0 = a complete site step;
//   1 = the move phase (P0) of the states [s_lo, s_hi) only — nothing is committed, the outputs (moved states, their
//       tables and back-pointers, valid-term spans, fan-out counts, history adds) are then all-gathered across the GPUs;
//   2 = scan, emit, de-duplicate, leaders (P1-P4) on the whole batch, replicated;
//   3 = the key-order ranking (P5, all-pairs: the dominant phase of a large batch) for the tasks t = s_lo mod s_hi only
//       (s_lo = this GPU's rank, s_hi = number of GPUs); the partial ranks are then all-reduced (sum);
//   4 = merge (P6) and commit, replicated.
template <int MODE>
__global__ void __launch_bounds__(StepSync<MODE>::kThreads, 1) k_dp_run(DpTerms H, DpBufs B, int m0, int nsteps, int resume, int phase_part,
                                                                        int s_lo, int s_hi)
{
    const StepSync<MODE> grid;
    constexpr bool CL = MODE != kGrid;
    constexpr int kBailS = StepSync<MODE>::kBailStates, kBailC = StepSync<MODE>::kBailChildren;
    __shared__ int part[StepSync<MODE>::kThreads];
    __shared__ MoveStage stage[StepSync<MODE>::kThreads / 32];
    __shared__ uint64_t relw[64 * kDpVW];                    // window terms of the current site, relative to o_m
    // block mode: the de-duplication table of a small site lives in shared memory (see k_sr_run)
    constexpr int kSmemSlots = MODE == kBlock ? 512 : 1;
    __shared__ DdSlot stab[kSmemSlots];
    if (MODE == kBlock) for (int i = threadIdx.x; i < kSmemSlots; i += blockDim.x) stab[i] = DdSlot{~0ULL, ~0ULL, ~0ULL};   // as the global table: all-ones = no generation
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wid = tid >> 5, nw = nth >> 5;
    DpCtl *ctl = B.ctl;
    // private running copies: every thread advances them identically, thread 0 publishes them at the end
    int nin = ctl->nin, cur = ctl->cur, hm_off = ctl->hm_off, hb_off = ctl->hb_off, step = ctl->step, err = ctl->err;
    unsigned gen = ctl->gen;
    const int caps = ctl->caps, capc = ctl->capc, hm_cap = ctl->hm_cap, hb_cap = ctl->hb_cap, log_cap = ctl->log_cap;
    const int tstride = B.tstride, slab_stride = B.slab_stride, k = H.k;
    int it = resume ? ctl->adv_done : 0;                     // sites of this advance already done by the kernel in front
    bool bail = false;
    grid.sync();                                             // everyone has read ctl before anyone writes it
    unsigned long long tprof = gtime();
    for (; it < nsteps && !err; it++) {
        if (CL && nin > kBailS * nw) { bail = true; break; }   // too many states for this mode: the next kernel takes over
        const int m = m0 + it, m1 = m + 1;
        const DpLevelDev L = B.levels[m], Ln = B.levels[m1];
        const DpState *inS = B.S[cur];
        const double *inE = B.E[cur];
        DpState *outS = B.S[cur ^ 1];
        double *outE = B.E[cur ^ 1];
        if ((long long)hm_off + nin > hm_cap || step >= log_cap) { err = 7; break; }
        // ---- P0: one warp per state.  Every block first extracts the rows of the window's terms (last site in
        //      [m, m+k+2)) relative to o_m into shared memory: the serial loops below would otherwise wait for a
        //      global load per term.
        if (tid == 0 && phase_part <= 2) ctl->nuniq = 0;      // (parts 3 and 4 of a multi-GPU step read what part 2 counted)
        {
            const int b0 = H.boff[m], nwin = min(H.boff[min(m + k + 2, H.n + 1)] - b0, 64 * kDpVW);
            __syncthreads();
            for (int i = threadIdx.x; i < nwin; i += blockDim.x) relw[i] = term_rel(H, b0 + i, m - k + 1);
            __syncthreads();
        }
        for (int s = (phase_part == 1 ? s_lo : 0) + wid; s < (phase_part == 0 ? nin : (phase_part == 1 ? min(s_hi, nin) : 0)); s += nw) {
            // the state, its right environment and (while the tables are small) the sign table with its scratch
            // are staged in shared memory: the serial group algebra below never waits for L2
            MoveStage &ms = stage[threadIdx.x >> 5];
            __syncwarp();
            if (lane < (int)(sizeof(DpState) / 8)) ((uint64_t *)&ms.st)[lane] = ((const uint64_t *)&inS[s])[lane];
            __syncwarp();
            if (lane < (int)(sizeof(DpGroup) / 8)) ((uint64_t *)&ms.sr)[lane] = ((const uint64_t *)&L.groups[ms.st.sright])[lane];
            __syncwarp();
            const bool staged = tstride <= kMoveTab;
            double *gE = B.slabE + (size_t)s * 2 * slab_stride;
            DpBack *gB = B.slabB + (size_t)s * 2 * slab_stride;
            double *E = staged ? ms.E : gE, *tE = staged ? ms.E + kMoveTab : gE + slab_stride;
            DpBack *Bk = staged ? ms.B : gB, *tBk = staged ? ms.B + kMoveTab : gB + slab_stride;
            DpMoved &mv = B.moved[s];
            uint64_t valid[kDpVW];
            const int mr = dp_move_warp(H, m, s, ms.st, inE, tstride, ms.sr, mv, E, tE, Bk, tBk, B.hm + hm_off + s, &ctl->err, lane, relw, valid);
            __syncwarp();
            int cnt = 0;
            if (mr >= 0) {
                if (staged) for (int e = lane; e < (1 << mr); e += 32) { gE[e] = E[e]; gB[e] = Bk[e]; }
                Ech e;
                dp_valid_span(H, m, valid, e, relw);
                if (lane < e.n) { B.span[s].b[lane] = e.b[lane]; B.span[s].p[lane] = e.p[lane]; }   // e is the same in every lane; kDpMaxSr = 32
                if (lane == 0) B.span[s].n = e.n;
                const int sright = ms.st.sright;
                const int j0 = L.map_off[sright], j1 = L.map_off[sright + 1];
                uint64_t accm = 0;
                for (int base = j0; base < j1; base += 32) {
                    const int j = base + lane;
                    const unsigned am = __ballot_sync(0xffffffffu, j < j1 && dp_accepts(e, Ln.groups[L.map_idx[j]]));
                    cnt += __popc(am);
                    if (base - j0 < 64) accm |= (uint64_t)am << (base - j0);
                }
                if (lane == 0) mv.accm = accm;                         // P2 re-evaluates only beyond 64 candidates
            }
            if (lane == 0) B.nacc[s] = cnt;
        }
        if (phase_part == 1) break;                            // (uniform) the other GPUs' slices arrive by all-gather
        grid.sync();
        SGS_PROF(ctl, 0);
        // ---- P1
        if (blockIdx.x == 0 && phase_part <= 2) {
            const int tot = block_scan(B.nacc, B.off, nin, part);
            if (threadIdx.x == 0) {
                if (tot > capc) atomicMax(&ctl->err, 6);
                ctl->total = tot;
                ctl->err_snap = ctl->err;
            }
        }
        grid.sync();
        SGS_PROF(ctl, 1);
        const int total = ctl->total;
        err = ctl->err_snap;
        if (err) break;
        if (CL && total > kBailC * nth) { bail = true; break; }    // nothing of this site is committed yet: the next kernel redoes it
        DdView D = B.D;
        if (MODE == kBlock && 2 * total <= kSmemSlots) { D.tab = stab; D.mask = kSmemSlots - 1; }
        // ---- P2: key = valid bits for last site in [m1+1, min(m1+k-1, n)), positions relative to boff[m1]
        if (phase_part <= 2) {
            const int lo = H.boff[min(m1 + 1, H.n + 1)] - H.boff[m1];
            const int hi = H.boff[min(max(m1 + k - 1, m1 + 1), H.n)] - H.boff[m1];
            for (int s = wid; s < nin; s += nw) {
                const DpMoved &mv = B.moved[s];
                if (mv.dead) continue;
                const Ech &e = B.span[s];
                uint64_t kv[kDpVW];
                for (int i = 0; i < kDpVW; i++) kv[i] = 0;
                for (int i = 0; i < kDpVW; i++) {                      // the valid bits [lo, hi)
                    const int a = max(lo - 64 * i, 0), z = min(hi - 64 * i, 64);
                    if (a < z) kv[i] = mv.valid[i] & ((z >= 64 ? ~0ULL : (1ULL << z) - 1ULL) & ~((1ULL << a) - 1ULL));
                }
                int c = B.off[s];
                const int j0 = L.map_off[mv.sright], j1 = L.map_off[mv.sright + 1];
                for (int base = j0; base < j1; base += 32) {          // children in (parent, map) order
                    const int j = base + lane;
                    const int nx = j < j1 ? L.map_idx[j] : 0;
                    const bool acc = j < j1 && (j - j0 < 64 ? (bool)((mv.accm >> (j - j0)) & 1) : dp_accepts(e, Ln.groups[nx]));
                    const unsigned am = __ballot_sync(0xffffffffu, acc);
                    if (acc) {
                        const int cc = c + __popc(am & ((1u << lane) - 1u));
                        DpKey &K = B.keys[cc];
                        for (int i = 0; i < kDpRows; i++) K.g[i] = mv.g[i];
                        for (int i = 0; i < kDpVW; i++) K.v[i] = kv[i];
                        K.r = mv.r;
                        K.sright = nx;
                        B.parent[cc] = s;
                    }
                    c += __popc(am);
                }
            }
        }
        grid.sync();
        SGS_PROF(ctl, 2);
        // ---- P3
        if (phase_part <= 2) {
            const int lo = H.boff[min(m1 + 1, H.n + 1)] - H.boff[m1];
            const int hi = H.boff[min(max(m1 + k - 1, m1 + 1), H.n)] - H.boff[m1];
            const int vlo = min(lo, 63);
            const PrefixSpec ps{32 - __clz(max(Ln.count, 1)), 2 * (k + 1), vlo, max(min(hi, 64) - vlo, 0)};
            for (int c = tid; c < total; c += nth) dd_insert(D, gen, B.keys, c, ps);
        }
        grid.sync();
        SGS_PROF(ctl, 3);
        // ---- P4
        if (phase_part <= 2) for (int c = tid; c < total; c += nth) dd_lead(D, c, &ctl->nuniq);
        grid.sync();
        SGS_PROF(ctl, 4);
        if (phase_part == 2) break;                            // (uniform) the ranking is shared out next
        const int nuniq = ctl->nuniq;
        if (nuniq > caps) { err = 5; break; }
        if ((long long)hb_off + (long long)nuniq * tstride > hb_cap) { err = 7; break; }
        // ---- P5
        if (phase_part == 0) dd_rank_tiles(D, B.keys, nuniq, wid, nw, lane);
        else if (phase_part == 3) dd_rank_tiles(D, B.keys, nuniq, wid, nw, lane, s_lo, s_hi);
        grid.sync();
        SGS_PROF(ctl, 5);
        if (phase_part == 3) break;                            // (uniform) partial ranks: summed over the GPUs next
        // ---- P6: exact ties are won by the later child (merge_gs keeps the newcomer unless the incumbent is
        //      strictly lower, cpp/stab.cpp:832); valid bits outside the key come from the first child
        //      (State::merge_gs copies state1, cpp/state.cpp:415-419)
        //      A key is merged by a group of 8, 16 or 32 lanes (the table has at most tstride entries), so a warp
        //      merges up to four keys at once; the merged state is written one word per lane.
        {
            const int gl = tstride >= 32 ? 32 : (tstride <= 8 ? 8 : 16), per = 32 / gl, sub = lane / gl, sl = lane % gl;
            for (int p = wid * per + sub; p < nuniq; p += nw * per) {
                const int c = D.lidx[p], dst = D.rank[c], r = B.keys[c].r;
                const int head = (int)(unsigned)D.tab[D.owner[c]].head;
                for (int e = sl; e < (1 << r); e += gl) {
                    double best = 0.0;
                    int bj = -1, bp = 0;
                    for (int j = head; j >= 0; j = D.next[j]) {
                        const int pj = B.parent[j];
                        const double v = B.slabE[(size_t)pj * 2 * slab_stride + e];
                        if (bj < 0 || v < best || (v == best && j > bj)) { best = v; bj = j; bp = pj; }
                    }
                    outE[(size_t)dst * tstride + e] = best;
                    B.hb[(size_t)hb_off + (size_t)dst * tstride + e] = B.slabB[(size_t)bp * 2 * slab_stride + e];
                }
                DpState &S = outS[dst];
                const int pp = B.parent[c];
                for (int i = sl; i < kDpRows + kDpVW + 1; i += gl) {
                    if (i < kDpRows) S.g[i] = B.keys[c].g[i];
                    else if (i < kDpRows + kDpVW) S.valid[i - kDpRows] = B.moved[pp].valid[i - kDpRows];
                    else { S.r = r; S.sright = B.keys[c].sright; }
                }
            }
        }
        if (tid == 0) {
            DpStepLog &Lg = B.log[step];
            Lg.m = m; Lg.nin = nin; Lg.total = total; Lg.nuniq = nuniq; Lg.hm_off = hm_off; Lg.hb_off = hb_off;
        }
        grid.sync();
        SGS_PROF(ctl, 6);
        hm_off += nin; hb_off += nuniq * tstride; step++; nin = nuniq; cur ^= 1; gen++;
    }
    grid.sync();                                             // every exit above is uniform; nobody reads ctl any more
    if (phase_part >= 1 && phase_part <= 3) {                 // a partial step: nothing is committed; errors stay in ctl->err for the next part
        if (tid == 0 && err) atomicMax(&ctl->err, err);
        return;
    }
    if (tid == 0) {
        if (err) { atomicMax(&ctl->err, err); nin = 0; }
        ctl->nin = nin; ctl->cur = cur; ctl->hm_off = hm_off; ctl->hb_off = hb_off; ctl->step = step;
        ctl->gen = gen + 1;                  // an aborted step may have left claims of generation gen in the table
        ctl->total = 0; ctl->nuniq = nin;
        ctl->adv_done = it;
        if (bail) ctl->bailed |= 1 << MODE;
    }
}

// history of (state, entry) of the current site back to step `floor`: (term position, sign) pairs in the order
// the reference appends them (Ps_indices / Ps_signs, cpp/stab.cpp:752-759).  One thread.
__global__ void k_dp_trace(const DpCtl *ctl, const DpStepLog *log, const DpHistAdd *hm, const DpBack *hb, int tstride, int floor_step,
                           int state, int entry, int *out_pos, int *out_sign, int *out_n, int cap)
{
    // the step log of the newest kTraceStage steps is staged by the whole block: the walk itself is a pointer
    // chase (back-pointer -> what its parent added), one thread
    constexpr int kTraceStage = 256;
    __shared__ DpStepLog slog[kTraceStage];
    const int nst = ctl->step, lo = max(floor_step, nst - kTraceStage);
    for (int t = lo + (int)threadIdx.x; t < nst; t += blockDim.x) slog[t - lo] = log[t];
    __syncthreads();
    if (threadIdx.x || blockIdx.x) return;
    int n = 0, s = state, e = entry;
    for (int t = nst - 1; t >= floor_step; t--) {
        const DpStepLog L = t >= lo ? slog[t - lo] : log[t];
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

// ------------------------------------------------------------------------------------------------
// k_sr_run: generate_Sright_all (cpp/state.cpp:218-238) for the levels m_hi-1 ... m_lo in one persistent
// cooperative launch: evolve_stab for every group of level m+1 (one thread each), get_map_from_stabs
// (cpp/state.cpp:187-216) as keyed de-duplication + key order, level tables bump-allocated on the device.
// The list of compatible next groups of a group is a set (py/state.py:523); its order is not significant.
struct SrCtl {
    int err, err_snap, nuniq, done_m;
    int bailed;                  // bit MODE: a small-mode kernel met a level too large for it (sticky)
    int nbr[3];                  // rotating branch counters of the waves
    unsigned gen;
    unsigned long long pool_used, pool_cap;
    unsigned long long prof[10];  // ns spent per phase (thread 0's view), for SGS_DP_TRACE
};
struct SrBufs {
    SrBranch *br[2];             // [cap] branch lists (ping-pong)
    DpGroup *flat;               // [cap] projected children
    int *par, *cnt, *moff;       // per child / per group
    uint8_t *first;
    DdView D;
    SrCtl *ctl;
    DpLevelDev *levels;
    char *pool;
    int cap;
    int waves_only;              // 1: never use the thread-per-path mode (tests, SGS_SR_WAVES=1)
};
constexpr int kSrPathBits = 16;             // thread-per-path mode: at most this many terms on the site ...
constexpr long long kSrPathMax = 1 << 17;   // ... and this many paths (groups << terms: a few rounds of the grid); else breadth-first waves

SGS_D void *sr_take(SrCtl *ctl, char *pool, size_t bytes)       // one thread
{
    const unsigned long long at = (ctl->pool_used + 255ULL) & ~255ULL;
    if (at + bytes > ctl->pool_cap) { atomicMax(&ctl->err, 8); return nullptr; }
    ctl->pool_used = at + bytes;
    return pool + at;
}

// The link lists of a level are filled through atomic cursors: sorting every list makes the level table (and with it
// the child order and the accept masks of the DP) the same on every run and on every GPU of a multi-GPU DP.  One thread
// per group; the lists are short.  Runs for level `lv` while the next level's first phase is under way (nothing there
// touches level lv's links), so it costs no barrier.
SGS_D void sr_sort_links(const DpLevelDev &Lv, int tid, int nth)
{
    int *idx = const_cast<int *>(Lv.map_idx);
    for (int g = tid; g < Lv.count; g += nth) {
        const int a = Lv.map_off[g], z = Lv.map_off[g + 1];
        for (int i = a + 1; i < z; i++) {
            const int v = idx[i];
            int j = i - 1;
            while (j >= a && idx[j] > v) { idx[j + 1] = idx[j]; j--; }
            idx[j + 1] = v;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(StepSync<MODE>::kThreads, 1) k_sr_run(DpTerms H, SrBufs B, int m_hi, int m_lo, int resume)
{
    const StepSync<MODE> grid;
    constexpr bool CL = MODE != kGrid;
    constexpr int kBailC = StepSync<MODE>::kBailChildren;
    __shared__ int part[StepSync<MODE>::kThreads];
    __shared__ uint64_t srel[32];
    // block mode: the de-duplication table of a small level lives in shared memory (generic-address atomics on
    // it stay inside the SM instead of three dependent L2 round trips per entry); same generation-tag scheme
    constexpr int kSmemSlots = MODE == kBlock ? 1024 : 1;
    __shared__ DdSlot stab[kSmemSlots];
    if (MODE == kBlock) for (int i = threadIdx.x; i < kSmemSlots; i += blockDim.x) stab[i] = DdSlot{~0ULL, ~0ULL, ~0ULL};   // as the global table: all-ones = no generation
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wid = tid >> 5, nw = nth >> 5;
    SrCtl *ctl = B.ctl;
    unsigned gen = ctl->gen;
    int err = ctl->err;
    const int k = H.k;
    int m = (resume ? ctl->done_m : m_hi) - 1;               // levels above were built by the kernel in front
    bool bail = false;
    if (tid == 0) { ctl->nbr[0] = 0; ctl->nbr[1] = 0; ctl->nbr[2] = 0; ctl->nuniq = 0; }
    grid.sync();
    unsigned long long tprof = gtime();
    int unsorted = -1;                                        // level whose link lists are filled but not yet sorted
    for (; m >= m_lo && !err; m--) {
        const DpLevelDev Ln = B.levels[m + 1];
        if (unsorted >= 0) { sr_sort_links(Ln, tid, nth); unsorted = -1; }      // (unsorted == m + 1)
        const int o = m - k + 1, b0 = H.boff[m], b1 = H.boff[m + 1];
        if (b1 - b0 > 30) { err = 3; break; }
        if (Ln.count > B.cap) { err = 4; break; }
        // ---- branches.  Thread-per-path mode (the usual case: few terms per site): every root-to-leaf path of
        //      every group's include/exclude tree is tried by one thread, which projects the branch it reaches -
        //      no barrier between the terms and none before the projection.  ctl->nbr[0] / nuniq are zero here
        //      (cleared before the loop and, for the next level, after their last read below).
        const int nt = b1 - b0;
        const bool paths = !B.waves_only && nt <= kSrPathBits && ((long long)Ln.count << nt) <= kSrPathMax;
        int nbr = Ln.count, cur = 0, w = 0;
        if (CL && (!paths || (Ln.count << nt) > kBailC * nth)) { bail = true; break; }    // the next kernel takes over from this level
        if (paths) {
            __syncthreads();
            if ((int)threadIdx.x < nt) srel[threadIdx.x] = term_rel(H, b0 + threadIdx.x, o);
            __syncthreads();
            const int npath = Ln.count << nt;
            for (int bw = tid - lane; bw < npath; bw += nth) {        // warp-uniform: one allocation per warp
                const int id = bw + lane;
                bool live = id < npath;
                SrBranch b;
                if (live) {
                    sr_seed(Ln.groups[id >> nt], id >> nt, b);
                    for (int j = 0; j < nt && live; j++) live = sr_step(srel, j, (id >> j) & 1, b);
                }
                const unsigned lm = __ballot_sync(0xffffffffu, live);
                int base = 0;
                if (lane == 0 && lm) base = atomicAdd(&ctl->nbr[0], __popc(lm));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int at = base + __popc(lm & ((1u << lane) - 1u));
                if (live && base + __popc(lm) <= B.cap) { sr_project(b, k, B.flat[at]); B.par[at] = b.par; }
            }
            grid.sync();
            nbr = ctl->nbr[0];
        } else {
            // ---- waves: seed, then one wave per term whose last site is m (branch order is not significant)
            for (int p = tid; p < nbr; p += nth) sr_seed(Ln.groups[p], p, B.br[0][p]);
            if (tid == 0) { ctl->nbr[0] = 0; ctl->nbr[1] = 0; }
            grid.sync();
            for (int pos = b0; pos < b1; pos++, w++) {
                int *counter = &ctl->nbr[w % 3];
                if (tid == 0) ctl->nbr[(w + 1) % 3] = 0;        // next wave's counter; last read two barriers ago
                const SrBranch *in = B.br[cur];
                SrBranch *out = B.br[cur ^ 1];
                for (int bw = tid - lane; bw < nbr; bw += nth) {          // warp-uniform: one allocation per warp
                    const int b = bw + lane;
                    SrBranch tmp[2];
                    const int nn = b < nbr ? sr_branch(H, o, b0, pos, in[b], tmp) : 0;
                    const unsigned m1 = __ballot_sync(0xffffffffu, nn >= 1), m2 = __ballot_sync(0xffffffffu, nn == 2);
                    const int tot = __popc(m1) + __popc(m2);
                    int base = 0;
                    if (lane == 0 && tot) base = atomicAdd(counter, tot);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned lt = (1u << lane) - 1u;
                    const int at = base + __popc(m1 & lt) + __popc(m2 & lt);
                    if (base + tot <= B.cap) for (int i = 0; i < nn; i++) out[at + i] = tmp[i];
                }
                grid.sync();
                nbr = *counter;
                if (nbr > B.cap) break;                          // uniform
                cur ^= 1;
            }
        }
        SGS_PROF(ctl, 0);
        if (nbr > B.cap) { err = 4; break; }
        if (CL && nbr > kBailC * nth) { bail = true; break; } // nothing of this level is stored yet
        const int total = nbr;
        DdView D = B.D;
        if (MODE == kBlock && 2 * total <= kSmemSlots) { D.tab = stab; D.mask = kSmemSlots - 1; }
        if (!paths) {
            // ---- projection of every branch = the children, with their parent
            for (int c = tid; c < total; c += nth) { sr_project(B.br[cur][c], k, B.flat[c]); B.par[c] = B.br[cur][c].par; }
            grid.sync();
        }
        SGS_PROF(ctl, 1);
        // ---- distinct groups
        {
            const PrefixSpec ps{0, 2 * (k + 1), 0, 0};
            for (int c = tid; c < total; c += nth) dd_insert(D, gen, B.flat, c, ps);
        }
        grid.sync();
        SGS_PROF(ctl, 2);
        for (int c = tid; c < total; c += nth) dd_lead(D, c, &ctl->nuniq);
        grid.sync();
        SGS_PROF(ctl, 3);
        const int nuniq = ctl->nuniq;
        // ---- key order; level storage; clear the per-group counters
        dd_rank_tiles(D, B.flat, nuniq, wid, nw, lane);
        for (int g = tid; g <= nuniq; g += nth) B.cnt[g] = 0;
        if (tid == 0) {
            DpLevelDev &Lv = B.levels[m];
            Lv.groups = (const DpGroup *)sr_take(ctl, B.pool, sizeof(DpGroup) * (size_t)max(nuniq, 1));
            Lv.map_off = (const int *)sr_take(ctl, B.pool, sizeof(int) * (size_t)(nuniq + 1));
            Lv.count = nuniq;
            ctl->err_snap = ctl->err;
        }
        grid.sync();
        SGS_PROF(ctl, 4);
        err = ctl->err_snap;
        if (err) break;
        // ---- groups in key order; a (group, parent) link is counted once
        if (tid == 0) { ctl->nbr[0] = 0; ctl->nbr[1] = 0; ctl->nbr[2] = 0; ctl->nuniq = 0; }   // for the next level; all read before the last barrier
        {
            DpGroup *groups = const_cast<DpGroup *>(B.levels[m].groups);
            for (int c = tid; c < total; c += nth) {
                const int ld = D.leader[c], g = D.rank[ld];
                if (ld == c) groups[g] = B.flat[c];
                const bool first = dd_pair_first(D, gen + 1, D.leader, B.par, c);
                B.first[c] = first;
                if (first) atomicAdd(&B.cnt[g], 1);
            }
        }
        grid.sync();
        SGS_PROF(ctl, 5);
        // ---- link offsets
        if (blockIdx.x == 0) {
            const int nl = block_scan(B.cnt, B.moff, nuniq, part);
            int *map_off = const_cast<int *>(B.levels[m].map_off);
            for (int g = threadIdx.x; g <= nuniq; g += blockDim.x) { map_off[g] = B.moff[g]; if (g < nuniq) B.cnt[g] = B.moff[g]; }
            if (threadIdx.x == 0) {
                DpLevelDev &Lv = B.levels[m];
                Lv.map_idx = (const int *)sr_take(ctl, B.pool, sizeof(int) * (size_t)max(nl, 1));
                Lv.nlinks = nl;
                ctl->err_snap = ctl->err;
            }
        }
        grid.sync();
        SGS_PROF(ctl, 6);
        err = ctl->err_snap;
        if (err) break;
        // ---- fill the links (cnt[g] is now the group's write cursor)
        {
            int *map_idx = const_cast<int *>(B.levels[m].map_idx);
            for (int c = tid; c < total; c += nth)
                if (B.first[c]) map_idx[atomicAdd(&B.cnt[D.rank[D.leader[c]]], 1)] = B.par[c];
        }
        grid.sync();
        SGS_PROF(ctl, 7);
        gen += 2;
        unsorted = m;
    }
    if (unsorted >= 0) sr_sort_links(B.levels[unsorted], tid, nth);            // the last level this kernel filled
    grid.sync();
    if (tid == 0) {
        if (err) atomicMax(&ctl->err, err);
        ctl->gen = gen + 2;                  // an aborted level may have left claims of generations gen, gen + 1 in the table
        ctl->done_m = m + 1;
        if (bail) ctl->bailed |= 1 << MODE;
    }
}

}  // namespace sgs
