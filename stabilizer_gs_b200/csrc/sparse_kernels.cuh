// sparse_kernels.cuh — sm_100a kernels of the sparse exact stabilizer-ground-state search.
//
// What they replace (reference): FullState::{include_pauli, exclude_pauli, evolve}
// (cpp/state.cpp:578-687) together with everything they call — Pauli::commute/dot (cpp/stab.cpp:314-345),
// Stab::canonicalize/decompose/include (cpp/stab.cpp:437-511), linear::Gauss_elimination/solve_modN
// (cpp/linear.cpp:21-121) and the StabEnergies sign tables (cpp/stab.cpp:667-843).
//
// Design (not a port):
//   The reference's include/exclude tree has one leaf per closed commuting group (= distinct GF(2) span
//   of a pairwise-commuting subset of terms).  We enumerate exactly those groups, each once, by
//   prefix-preserving closure extension: a group G with canonical generating terms g_1 < ... < g_m
//   (g_i = lowest-index term of G outside <g_1..g_{i-1}>) is extended by a term j > g_m that commutes
//   with G, lies outside G, and whose coset j*G contains no lower-index term.
//
//   k_adjacency : T x T commutation bit-matrix (POPC of a & swapzx(b)).
//   k_general   : one warp evaluates ONE group from scratch, for any input: echelon basis with phase
//                 and coefficient tracking, membership/sign of every commuting term, exact minimum over
//                 generator signs, valid children.  Used for the top levels of the tree (work-item
//                 generation) and as the exact path for groups that contain linearly dependent terms.
//   k_dfs       : the hot kernel.  Persistent warps pull subtree roots from a queue and walk them
//                 depth-first; a tree level is a compacted list (residue mod G, term index) of the terms
//                 still commuting with G, one lane per term, kept in shared memory.  A child costs one
//                 pass over the parent's list: POPC commutation test, conditional XOR with the new
//                 generator's residue at its pivot bit, __ballot_sync compaction, __match_any_sync
//                 duplicate-residue detection.  While residues stay pairwise distinct no term outside
//                 the generating set can enter the group, so E_min = -sum |w_g| exactly; the first
//                 duplicate hands that node to k_general.  Once a list has <= 96 entries, <= 64 of them
//                 candidates, the warp proves by Gaussian elimination that no dependent term can appear anywhere
//                 below it (level_b); the subtree is then the set of cliques of the candidates' commutation
//                 graph, walked by the lanes independently with bit masks (lb_walk_edges).  Template flags: SPLIT (multi-shard runs:
//                 heavy or uncertifiable roots hand their children to a next round), PRUNE (optional
//                 branch and bound against a device-wide incumbent).
//   k_class_count / k_class_scatter : heavy-first order of the subtree roots (multi-shard runs).
//   k_reduce_best / k_finalize : deterministic argmin over per-warp records, then canonical RREF with
//                 sign tracking, per-term signs and the reference's first-minimum sign pattern.
#pragma once
#include <cstdint>

#include "packed.cuh"

namespace sgs {

constexpr int kMaxG = 64;            // max rank (generators per group) in the sparse search
constexpr int kMaxFrames = kMaxG + 1;
constexpr int kCapM = 256;           // max non-trivial (dependent) members of one group in k_general
constexpr int kMaxBrute = 26;        // max entangled sign variables minimised by enumeration
constexpr unsigned kFlagUnclean = 1; // node has members outside its generating set
constexpr unsigned kFlagExpandOnly = 2;   // already counted and evaluated: only emit children

struct BestRec {
    double E;
    int m;
    int valid;
    uint16_t g[kMaxG];
};

struct SparseStats {
    unsigned long long hist[kMaxG + 1];   // candidates per rank
    unsigned long long slow_nodes;        // candidates evaluated by k_general
    unsigned long long paths[9];          // which code paths of the walk ran (kPath*; summed over shards like hist)
    unsigned int error;                   // SGS_ERR_* (positive) set by a kernel
    unsigned int pad;
    unsigned long long dbg[8];            // SGS_TRACE diagnostics of the walk (see k_dfs)
};

struct SparseDev {
    int n, T, TW, S;             // TW = ceil(T/64); S = node record stride in u16 (2 + max rank)
    const uint64_t *rows;        // [T * W] terms in paulis_list order
    const double *wgt;           // [T]
    const double *wabs;          // [T]
    const uint64_t *adj;         // [T * TW] commutation bit-matrix (diagonal 0)
    BestRec *wbest;              // per-warp best records
    SparseStats *stats;
    uint16_t *slow;              // work stack of nodes needing k_general
    unsigned int *slow_count;
    unsigned int slow_cap;
    int trace;                   // SGS_TRACE: per-root timing diagnostics into stats->dbg
    unsigned long long *inc;     // branch-and-bound mode: best energy found so far on this device (order-encoded double)
    unsigned long long *wtime;   // SGS_WALK_TRACE builds: [2 * warps] start / end time of every walk warp (globaltimer)
    float split_cliques;         // multi-shard runs: expected cliques from which a root hands its children to the next k_dfs round
};

enum { kErrOverflow = 8, kErrLimit = 6 };
constexpr int kStatsSummed = kMaxG + 2 + 9;   // leading u64 words of SparseStats that add up over shards (hist, slow_nodes, paths)
// SparseStats::paths — diagnostics the parity tests assert on (sgs_result.walk_paths)
enum {
    kPathWalk32 = 0,     // certified frames walked with 32-bit clique masks (<= 32 candidates)
    kPathWalk64 = 1,     // certified frames walked with 64-bit clique masks
    kPathFoldOk = 2,     // multi-word rows: independence proven on the folded 64-bit image
    kPathFoldFull = 3,   // multi-word rows: the image was dependent, the full-width elimination decided
    kPathDepCert = 4,    // dependent residues, certified because no dependency is a clique
    kPathNoCert = 5,     // frames level B could not certify (commuting dependency / > kMaxDep dependencies)
    kPathLevelA = 6,     // warp-cooperative child frames pushed below a root
    kPathWide = 7,       // certified frames of more than 64 entries (three rows per lane in the certificate)
    kPathSplit = 8       // roots whose children were handed to a next k_dfs round
};
constexpr int kNumPaths = 9;

SGS_D unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// order-preserving encoding of a double into an unsigned key (atomicMin on the key = min on the double)
SGS_D unsigned long long enc_double(double x)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
SGS_D double dec_double(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
    return __longlong_as_double((long long)u);
}
// pruning threshold of the branch-and-bound mode: a subtree whose energy lower bound exceeds it cannot reach
// the incumbent, not even tie it (the margin covers the different summation orders of bound and energy)
SGS_D double prune_threshold(double inc) { return inc >= 1e299 ? 1e300 : inc + 1e-9 * (1.0 + fabs(inc)); }

SGS_D uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// lexicographic order on generating tuples with "end of tuple" sorting last: the order in which the
// reference's include-first tree lists its leaves (tie-break).
SGS_D bool tuple_before(const uint16_t *a, int na, const uint16_t *b, int nb)
{
    for (int i = 0;; i++) {
        if (i == na) return false;
        if (i == nb) return true;
        if (a[i] != b[i]) return a[i] < b[i];
    }
}

SGS_D uint32_t tuple_hash(const uint16_t *g, int m)
{
    uint32_t h = 2166136261u;
    for (int i = 0; i < m; i++) { h ^= g[i]; h *= 16777619u; }
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return h;
}

// ------------------------------------------------------------------------------------------------
// K3: commutation bit-matrix.  One thread per (term i, 64-term word).
template <int W>
__global__ void k_adjacency(int T, int TW, const uint64_t *__restrict__ rows, uint64_t *__restrict__ adj)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * TW) return;
    int i = idx / TW, wj = idx % TW;
    uint64_t sw[W];
#pragma unroll
    for (int k = 0; k < W; k++) sw[k] = swapzx(rows[(size_t)i * W + k]);
    uint64_t bits = 0;
    for (int b = 0; b < 64; b++) {
        int j = wj * 64 + b;
        if (j >= T) break;
        if (j == i) continue;
        if (!row_anticommute_sw<W>(rows + (size_t)j * W, sw)) bits |= 1ULL << b;
    }
    adj[idx] = bits;
}

// push one node record onto the slow work stack (warp-uniform call)
// (cold paths are kept out of line: the hot kernel's code must stay inside the instruction caches)
__device__ __noinline__ void push_slow(const SparseDev &D, const uint16_t *g, int m, unsigned flags, int lane)
{
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(D.slow_count, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= D.slow_cap) {
        if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrOverflow);
        return;
    }
    uint16_t *rec = D.slow + (size_t)slot * D.S;
    if (lane == 0) { rec[0] = (uint16_t)m; rec[1] = (uint16_t)flags; }
    for (int i = lane; i < m; i += 32) rec[2 + i] = g[i];
}

// alive terms of the group generated by g[0..m): bitset AND of adjacency rows, compacted into tt[]
// (ascending term index).  Returns the count, or -1 if it exceeds cap.
// The AND is taken per 64-bit word (one lane per word), the expansion of the result into indices per SLICE of
// 16 bits (up to 8 words = 512 terms per round) or 32 bits (16 words = 1024 terms per round), one slice per lane:
// a list holds about T / 2^m entries, so a word has a dozen set bits but a slice one or two — the serial
// "lowest bit, clear it, store" loop is as long as the fullest slice, not the fullest word.
SGS_D int alive_list(const SparseDev &D, const uint16_t *g, int m, uint16_t *tt, int cap, int lane)
{
    int total = 0;
    const bool narrow = D.TW <= 8;                       // 16-bit slices: 4 per word
    const int wpr = narrow ? 8 : 16;                     // words per round
    for (int w0 = 0; w0 < D.TW; w0 += wpr) {
        const int w = w0 + lane;
        uint64_t a = 0;
        if (lane < wpr && w < D.TW) {
            a = ~0ULL;
            const int rem = D.T - w * 64;
            if (rem < 64) a = (1ULL << rem) - 1;
            for (int i = 0; i < m; i++) a &= D.adj[(size_t)g[i] * D.TW + w];
        }
        // my slice: lane j reads word j / 4 (narrow) or j / 2 of this round
        const uint64_t aw = __shfl_sync(0xffffffffu, a, narrow ? (lane >> 2) : (lane >> 1));
        uint32_t sl = narrow ? (uint32_t)(aw >> (16 * (lane & 3))) & 0xffffu : (uint32_t)(aw >> (32 * (lane & 1)));
        const int bit0 = w0 * 64 + (narrow ? 16 : 32) * lane;
        const int c = __popc(sl);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        int off = total + inc - c;
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (total + tot > cap) return -1;
        while (sl) {
            const int b = __ffs((int)sl) - 1;
            sl &= sl - 1;
            tt[off++] = (uint16_t)(bit0 + b);
        }
        total += tot;
    }
    __syncwarp();
    return total;
}

// ------------------------------------------------------------------------------------------------
// k_general: exact from-scratch evaluation of one closed commuting group per warp.
// smem per warp (bytes): see general_smem_bytes().
__host__ __device__ inline int general_hash_slots(int T)
{
    int s = 64;
    while (s < 2 * T) s <<= 1;
    return s;
}

template <int W>
__host__ __device__ inline size_t general_smem_bytes(int T)
{
    size_t b = 0;
    b += (size_t)kMaxG * W * 8;      // basis rows
    b += (size_t)kMaxG * 8;          // basis coefficient masks
    b += (size_t)kMaxG * 8;          // lin[]
    b += (size_t)T * W * 8;          // alive residues
    b += (size_t)kCapM * 8 * 2;      // extras: mask, weight
    b += (size_t)kCapM * 4;          // extras: compressed mask
    b += (size_t)((T + 3) / 4) * 8;  // alive term indices (u16)
    b += kMaxG * 2 + 64;             // pivots, phases, vpos
    b += kMaxG * 2;                  // best tuple
    b += (kMaxG + 1) * 4 + 8;        // hist
    b += (size_t)general_hash_slots(T) * 4;   // residue hash table (coset classes)
    b += (size_t)((T + 7) / 8) * 8;           // class-has-another-member flags
    return (b + 15) & ~(size_t)15;
}

// Coset classes of the alive residues in O(na): an open-addressing table maps each distinct residue to the
// LOWEST list index carrying it.  leader[i] = that index; multi[j] != 0 iff class j has more than one member.
template <int W>
SGS_D uint32_t residue_hash(const uint64_t *r)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int k = 0; k < W; k++) { h ^= r[k]; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 29; }
    return (uint32_t)(h ^ (h >> 32));
}

template <int W>
SGS_D void residue_classes(const uint64_t *R, int na, uint32_t *tab, int slots, uint8_t *multi, int lane)
{
    for (int i = lane; i < slots; i += 32) tab[i] = 0xffffffffu;
    for (int i = lane; i < na; i += 32) multi[i] = 0;
    __syncwarp();
    for (int i = lane; i < na; i += 32) {
        uint32_t s = residue_hash<W>(R + (size_t)i * W) & (uint32_t)(slots - 1);
        for (;;) {
            uint32_t cur = tab[s];
            if (cur == 0xffffffffu) {
                cur = atomicCAS(&tab[s], 0xffffffffu, (uint32_t)i);
                if (cur == 0xffffffffu) break;                       // first of its class in this slot
            }
            if (row_equal<W>(R + (size_t)cur * W, R + (size_t)i * W)) { atomicMin(&tab[s], (uint32_t)i); break; }
            s = (s + 1) & (uint32_t)(slots - 1);
        }
    }
    __syncwarp();
}

template <int W>
SGS_D int residue_leader(const uint64_t *R, int i, const uint32_t *tab, int slots)
{
    uint32_t s = residue_hash<W>(R + (size_t)i * W) & (uint32_t)(slots - 1);
    for (;;) {
        const uint32_t cur = tab[s];
        if (row_equal<W>(R + (size_t)cur * W, R + (size_t)i * W)) return (int)cur;
        s = (s + 1) & (uint32_t)(slots - 1);
    }
}

template <int W>
__global__ void __launch_bounds__(128)
k_general(SparseDev D, const uint16_t *__restrict__ in, unsigned nin, uint16_t *__restrict__ out,
          unsigned int *out_count, unsigned out_cap, int shard, int nshards, int smem_per_warp, int own_only, unsigned int *fetch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int T = D.T;
    unsigned char *p = smem_raw + (size_t)warp * smem_per_warp;
    uint64_t *B = (uint64_t *)p;            p += (size_t)kMaxG * W * 8;
    uint64_t *cB = (uint64_t *)p;           p += (size_t)kMaxG * 8;
    double *lin = (double *)p;              p += (size_t)kMaxG * 8;
    uint64_t *R = (uint64_t *)p;            p += (size_t)T * W * 8;
    uint64_t *exC = (uint64_t *)p;          p += (size_t)kCapM * 8;
    double *exA = (double *)p;              p += (size_t)kCapM * 8;
    uint32_t *exV = (uint32_t *)p;          p += (size_t)kCapM * 4;
    uint16_t *tt = (uint16_t *)p;           p += (size_t)((T + 3) / 4) * 8;
    uint8_t *pivB = (uint8_t *)p;           p += kMaxG;
    uint8_t *qB = (uint8_t *)p;             p += kMaxG;
    uint8_t *vpos = (uint8_t *)p;           p += 64;
    uint16_t *bestG = (uint16_t *)p;        p += kMaxG * 2;
    uint32_t *hist = (uint32_t *)p;         p += (kMaxG + 1) * 4 + 8;
    uint32_t *htab = (uint32_t *)p;         p += (size_t)general_hash_slots(T) * 4;
    uint8_t *multi = (uint8_t *)p;
    const int hslots = general_hash_slots(T);

    const unsigned gw = blockIdx.x * wpb + warp, nw = gridDim.x * wpb;
    BestRec *myrec = D.wbest + gw;
    double bestE = myrec->valid ? myrec->E : 1e300;
    int bestM = myrec->valid ? myrec->m : 0;
    int haveBest = myrec->valid;
    for (int i = lane; i < bestM; i += 32) bestG[i] = myrec->g[i];
    for (int i = lane; i <= kMaxG; i += 32) hist[i] = 0;
    unsigned slowNodes = 0;
    __syncwarp();

    // fetch != null (parent-sharded level: a shard evaluates one node in nshards, at ~15 us each): nodes are handed out
    // one by one, so that no warp is left with several owned nodes while others have none
    for (unsigned node = gw;; node += nw) {
        if (fetch) {
            unsigned f = 0;
            if (lane == 0) f = atomicAdd(fetch, 1u);
            node = __shfl_sync(0xffffffffu, f, 0);
        }
        if (node >= nin) break;
        if (*(volatile unsigned int *)&D.stats->error) break;      // a limit was hit somewhere: the search is void, stop early
        const uint16_t *rec = in + (size_t)node * D.S;
        const int m = rec[0];
        const unsigned flags = rec[1];
        const uint16_t *g = rec + 2;
        if (m > kMaxG || 2 + m > D.S) {           // a record no queue slot can hold: a corrupted or over-deep node
            if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrLimit);
            continue;
        }
        const bool mine = nshards <= 1 || (int)(tuple_hash(g, m) % (unsigned)nshards) == shard;
        if (own_only && !mine) continue;          // parent-sharded level: the owner evaluates and expands it

        // ---- 1. echelon basis of the generators with phase and coefficient tracking (K1) ----
        // basis[i] = exact operator product of the generator terms selected by cB[i]; pivot = its
        // lowest set bit; later basis rows are clear at every earlier pivot.
        for (int i = 0; i < m; i++) {
            uint64_t v[W];
            const int t = g[i];
#pragma unroll
            for (int k = 0; k < W; k++) v[k] = D.rows[(size_t)t * W + k];
            int q = row_plus_phase<W>(v);
            uint64_t c = 1ULL << i;
            for (int b = 0; b < i; b++) {
                if (row_bit<W>(v, pivB[b])) {
                    uint64_t bb[W];
#pragma unroll
                    for (int k = 0; k < W; k++) bb[k] = B[b * W + k];
                    row_mul<W>(v, q, bb, qB[b]);
                    c ^= cB[b];
                }
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < W; k++) B[i * W + k] = v[k];
                cB[i] = c;
                qB[i] = (uint8_t)q;
                pivB[i] = (uint8_t)row_lowbit<W>(v);
                lin[i] = D.wgt[t];
            }
            __syncwarp();
        }

        // ---- 2. terms commuting with every generator (K3, via the adjacency bitsets) ----
        int na = alive_list(D, g, m, tt, T, lane);

        // ---- 3. reduce each one modulo the group (K2): member (with sign) or alive residue ----
        double cst = 0.0;
        int ne = 0, nkeep = 0;
        bool overflow = false;
        for (int i0 = 0; i0 < na; i0 += 32) {
            const int i = i0 + lane;
            uint64_t v[W];
            int q = 0, t = 0;
            uint64_t c = 0;
            bool act = i < na;
            if (act) {
                t = tt[i];
#pragma unroll
                for (int k = 0; k < W; k++) v[k] = D.rows[(size_t)t * W + k];
                q = row_plus_phase<W>(v);
                for (int b = 0; b < m; b++) {
                    if (row_bit<W>(v, pivB[b])) {
                        uint64_t bb[W];
#pragma unroll
                        for (int k = 0; k < W; k++) bb[k] = B[b * W + k];
                        row_mul<W>(v, q, bb, qB[b]);
                        c ^= cB[b];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < W; k++) v[k] = 1;
            }
            const bool member = act && row_is_zero<W>(v);
            const bool keep = act && !member;
            // P_t * prod(used basis rows) = i^q * I with q in {0,2}: P_t = (-1)^{q/2} * prod_{k in c} P_{g_k}
            const double a = member ? ((q & 2) ? -D.wgt[t] : D.wgt[t]) : 0.0;
            unsigned mb = __ballot_sync(0xffffffffu, member);
            while (mb) {                                   // deterministic (term) order
                const int src = __ffs(mb) - 1;
                mb &= mb - 1;
                const uint64_t cc = __shfl_sync(0xffffffffu, c, src);
                const double aa = __shfl_sync(0xffffffffu, a, src);
                const int pc = popc64(cc);
                if (pc == 0) cst += aa;
                else if (pc == 1) { if (lane == 0) lin[ctz64(cc)] += aa; }
                else {
                    if (ne < kCapM) { if (lane == 0) { exC[ne] = cc; exA[ne] = aa; } }
                    else overflow = true;
                    ne++;
                }
                __syncwarp();
            }
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) {
                const int pos = nkeep + __popc(kb & lanemask_lt());
#pragma unroll
                for (int k = 0; k < W; k++) R[(size_t)pos * W + k] = v[k];
                tt[pos] = (uint16_t)t;
            }
            nkeep += __popc(kb);
            __syncwarp();
        }
        const int nmem = na - nkeep;
        na = nkeep;
        if (overflow) { if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrLimit); continue; }

        // ---- 4. exact minimum over generator signs (K5) ----
        // E(s) = cst + sum_k lin[k] s_k + sum_e exA[e] prod_{k in exC[e]} s_k, s_k = sign of generator term k.
        // Variables outside V = union of the extras' supports minimise independently.
        if (!(flags & kFlagExpandOnly) && out != nullptr) {       // out == nullptr: count-only pass (children are only counted)
            uint64_t V = 0;
            for (int e = 0; e < ne; e++) V |= exC[e];
            double E = cst;
            for (int k = 0; k < m; k++) if (!((V >> k) & 1)) E -= fabs(lin[k]);
            const int nv = popc64(V);
            if (nv > kMaxBrute) { if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrLimit); continue; }
            if (nv > 0) {
                if (lane == 0) {
                    int j = 0;
                    for (int k = 0; k < m; k++) if ((V >> k) & 1) vpos[j++] = (uint8_t)k;
                    for (int e = 0; e < ne; e++) {
                        uint32_t cv = 0;
                        for (int jj = 0; jj < nv; jj++) if ((exC[e] >> vpos[jj]) & 1) cv |= 1u << jj;
                        exV[e] = cv;
                    }
                }
                __syncwarp();
                double mn = 1e300;
                for (uint32_t pat = lane; pat < (1u << nv); pat += 32) {
                    double val = 0.0;
                    for (int jj = 0; jj < nv; jj++) val += ((pat >> jj) & 1) ? -lin[vpos[jj]] : lin[vpos[jj]];
                    for (int e = 0; e < ne; e++) val += (__popc(pat & exV[e]) & 1) ? -exA[e] : exA[e];
                    mn = fmin(mn, val);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
                E += mn;
            }
            if (mine) {
                if (lane == 0) hist[m]++;
                slowNodes++;
                if (!haveBest || E < bestE || (E == bestE && tuple_before(g, m, bestG, bestM))) {
                    haveBest = 1; bestE = E; bestM = m;
                    if (D.inc && lane == 0) atomicMin(D.inc, enc_double(E));
                    __syncwarp();
                    for (int i = lane; i < m; i += 32) bestG[i] = g[i];
                    __syncwarp();
                }
            }
        }

        // ---- 5. children: alive terms above the core whose coset holds no lower-index term ----
        const int core = m ? (int)g[m - 1] : -1;
        const bool parentUnclean = (flags & kFlagUnclean) || nmem > 0;
        // coset classes of the alive residues: child i is valid iff it is the lowest index of its class;
        // a class with further members brings dependent terms into the child (it is then "unclean")
        residue_classes<W>(R, na, htab, hslots, multi, lane);
        for (int i = lane; i < na; i += 32) {
            const int lead = residue_leader<W>(R, i, htab, hslots);
            if (lead != i) multi[lead] = 1;
        }
        __syncwarp();
        for (int i0 = 0; i0 < na; i0 += 32) {
            const int i = i0 + lane;
            bool valid = false, higher = false;
            if (i < na && (int)tt[i] > core) {
                valid = residue_leader<W>(R, i, htab, hslots) == i;
                higher = multi[i] != 0;
            }
            const unsigned vb = __ballot_sync(0xffffffffu, valid);
            if (vb == 0) continue;
            if (m >= kMaxG || 2 + m + 1 > D.S) {          // a child of rank m + 1 fits no record (rank limit): report, never write
                if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrLimit);
                break;
            }
            if (out == nullptr) { if (lane == 0) atomicAdd(out_count, (unsigned)__popc(vb)); continue; }
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(out_count, (unsigned)__popc(vb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(vb) > out_cap) {
                if (lane == 0) atomicMax(&D.stats->error, (unsigned)kErrOverflow);
                continue;
            }
            if (valid) {
                uint16_t *o = out + (size_t)(base + __popc(vb & lanemask_lt())) * D.S;
                o[0] = (uint16_t)(m + 1);
                o[1] = (uint16_t)((parentUnclean || higher) ? kFlagUnclean : 0);
                for (int k = 0; k < m; k++) o[2 + k] = g[k];
                o[2 + m] = tt[i];
            }
        }
        __syncwarp();
    }

    // ---- write back the warp's best record and statistics ----
    __syncwarp();
    if (haveBest) {
        if (lane == 0) { myrec->E = bestE; myrec->m = bestM; myrec->valid = 1; }
        for (int i = lane; i < bestM; i += 32) myrec->g[i] = bestG[i];
    }
    for (int i = lane; i <= kMaxG; i += 32)
        if (hist[i]) atomicAdd(&D.stats->hist[i], (unsigned long long)hist[i]);
    if (lane == 0 && slowNodes) atomicAdd(&D.stats->slow_nodes, (unsigned long long)slowNodes);
}

// ------------------------------------------------------------------------------------------------
// Heavy-first order of the subtree roots.  The size of a root's subtree grows steeply with the number of
// alive terms after its last generator, i.e. falls with the index of that generator; the expansion kernel
// emits roots in atomic order.  A counting sort into kRootClasses classes of the last generator index gives
// the walk a longest-first schedule (order inside a class is arbitrary: results do not depend on it).
constexpr int kRootClasses = 8;
SGS_D int root_class(const uint16_t *rec, int T) { const int m = rec[0]; return m ? (int)(((unsigned)rec[1 + m] * kRootClasses) / (unsigned)max(T, 1)) : 0; }

__global__ void k_class_count(const uint16_t *__restrict__ items, unsigned n, const unsigned *n_dev, int S, int T, unsigned *cls)
{
    if (n_dev) n = *n_dev;
    __shared__ unsigned loc[kRootClasses];
    if (threadIdx.x < kRootClasses) loc[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&loc[min(root_class(items + (size_t)i * S, T), kRootClasses - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < kRootClasses && loc[threadIdx.x]) atomicAdd(&cls[threadIdx.x], loc[threadIdx.x]);
}

__global__ void k_class_scatter(const uint16_t *__restrict__ items, unsigned n, const unsigned *n_dev, int S, int T, const unsigned *cls,
                                unsigned *cursor, unsigned *__restrict__ perm)
{
    if (n_dev) n = *n_dev;
    __shared__ unsigned loc[kRootClasses], base[kRootClasses];
    if (threadIdx.x < kRootClasses) loc[threadIdx.x] = 0;
    __syncthreads();
    // block-local counts first, one global atomic per class and block, then the scatter
    const unsigned per = (n + gridDim.x - 1) / gridDim.x, lo = min(n, blockIdx.x * per), hi = min(n, lo + per);
    for (unsigned i = lo + threadIdx.x; i < hi; i += blockDim.x)
        atomicAdd(&loc[min(root_class(items + (size_t)i * S, T), kRootClasses - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < kRootClasses) {
        unsigned off = 0;
        for (int c = 0; c < (int)threadIdx.x; c++) off += cls[c];
        base[threadIdx.x] = off + (loc[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], loc[threadIdx.x]) : 0u);
        loc[threadIdx.x] = 0;
    }
    __syncthreads();
    for (unsigned i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int c = min(root_class(items + (size_t)i * S, T), kRootClasses - 1);
        perm[base[c] + atomicAdd(&loc[c], 1u)] = i;
    }
}

// ------------------------------------------------------------------------------------------------
// k_dfs: the hot kernel.
// Shared memory of one warp.  The arena is what bounds the occupancy of the hot kernel, so everything that is not
// a list entry is kept small: the root's echelon basis and the level-B commutation matrix share one region (the
// basis is dead once the root's residues are written), the warp-cooperative level A keeps at most kMaxFramesA
// frames (a deeper chain of uncertifiable lists goes to the exact kernel), the suffix sums exist only in the
// branch-and-bound build.
constexpr int kMaxFramesA = 24;
constexpr int kMaxEdgesSmem = 256;      // = kMaxEdges (defined with the walk)
template <int W>
__host__ __device__ inline size_t dfs_scratch_bytes() { return (size_t)(32 * W * 8 > 64 * 8 + 32 * 4 ? 32 * W * 8 : 64 * 8 + 32 * 4); }
template <int W>
__host__ __device__ inline size_t dfs_smem_bytes(int cap, bool prune)
{
    size_t b = 0;
    b += (size_t)cap * W * 8;          // residues
    b += 64 * 8;                       // |w| of the candidates of the level-B frame being walked
    b += dfs_scratch_bytes<W>();       // root basis, later the level-B commutation bit-matrix (64-bit or 32-bit form)
    b += (size_t)kMaxEdgesSmem * 2;    // edge list of the level-B walk (one chunk)
    b += (size_t)kMaxFramesA * 8;      // frame E
    b += (size_t)kMaxFramesA * 4 * 3;  // frame base, a, q
    b += (size_t)(kMaxG + 1) * 4 + 4;  // hist
    b += (size_t)cap * 2;              // term indices
    b += (size_t)kMaxG * 2 * 2;        // path, best tuple
    b += 64;                           // basis pivots
    b += 8;                            // alignment slack
    if (prune) b += 64 * 8;            // suffix sums of the candidate weights (branch-and-bound mode)
    b += (kNumPaths + 1) * 4;          // code-path counters
    return (b + 15) & ~(size_t)15;
}

// any two equal residues among R[0..n)?  (no zero residue can occur: see the caller)
// Rows of 32 entries: __match_any_sync inside a row, broadcast-compare against the lower rows.
template <int W>
__device__ __forceinline__ bool has_duplicate_body(const uint64_t *R, int n, int lane)
{
    bool dup = false;
#pragma unroll 1
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int cnt = min(32, n - r0);
        if (lane < cnt) {
            const unsigned mask = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
            unsigned mm = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < W; k++) mm &= __match_any_sync(mask, R[(r0 + lane) * W + k]);
            dup |= (mm & (mm - 1)) != 0;
        }
        if (r0 > 0) {
#pragma unroll 1
            for (int e = r0; e < r0 + cnt; e++) {
                uint64_t ref[W];
#pragma unroll
                for (int k = 0; k < W; k++) ref[k] = R[e * W + k];
#pragma unroll 1
                for (int i = lane; i < r0; i += 32) dup |= row_equal<W>(R + i * W, ref);
            }
        }
    }
    return __any_sync(0xffffffffu, dup);
}
template <int W>
__device__ __noinline__ bool has_duplicate_cold(const uint64_t *R, int n, int lane) { return has_duplicate_body<W>(R, n, lane); }
// The same question for a list of 33..128 entries in O(n): the entries are inserted row by row (32 lanes at a time)
// into an open-addressing table of 256 16-bit slots (tab: the edge list's shared memory, unused between walks).  No
// atomics: the lanes that found their slot empty all write, one store lands, everyone reads back; a lane that meets
// another entry — an earlier row's or the winner of this round — compares residues with it and probes on.  (The
// all-pairs scheme above costs ~n^2/32 broadcast compares: 1 800 instructions at n = 125 against ~300 here.)
template <int W>
__device__ __noinline__ bool has_duplicate_hash(const uint64_t *R, int n, uint16_t *tab, int lane)
{
    for (int i = lane; i < 128; i += 32) ((uint32_t *)tab)[i] = 0;
    __syncwarp();
    bool dup = false;
#pragma unroll 1
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int i = r0 + lane;
        bool pending = i < n;
        uint64_t v[W];
        uint64_t hh = 0x9E3779B97F4A7C15ULL;
#pragma unroll
        for (int k = 0; k < W; k++) { v[k] = pending ? R[i * W + k] : 0; hh = (hh ^ v[k]) * 0xBF58476D1CE4E5B9ULL; hh ^= hh >> 31; }
        unsigned h = (unsigned)(hh >> 40) & 255u;
#pragma unroll 1
        while (__any_sync(0xffffffffu, pending)) {
            const unsigned occ = pending ? tab[h] : 0u;
            const bool tryw = pending && occ == 0u;
            if (pending && occ != 0u) {                       // an entry of an earlier row (or round) sits here
                dup |= row_equal<W>(R + (occ - 1u) * W, v);
                h = (h + 1u) & 255u;
            }
            if (tryw) tab[h] = (uint16_t)(i + 1);
            __syncwarp();
            if (tryw) {
                const unsigned got = tab[h];
                if (got == (unsigned)(i + 1)) pending = false;
                else { dup |= row_equal<W>(R + (got - 1u) * W, v); h = (h + 1u) & 255u; }
            }
            if (dup) pending = false;
            __syncwarp();
        }
    }
    return __any_sync(0xffffffffu, dup);
}
template <int W>
SGS_D bool has_duplicate(const uint64_t *R, int n, uint16_t *tab, int lane)
{
#ifndef SGS_HASH_W1
#define SGS_HASH_W1 0      // (one-word rows: the extra out-of-line call costs 3 % of config 4 through register allocation; their big lists are rare)
#endif
    if ((W > 1 || SGS_HASH_W1) && n > 32 && n <= 128) return has_duplicate_hash<W>(R, n, tab, lane);
    if (W == 1) return has_duplicate_body<W>(R, n, lane);
    return has_duplicate_cold<W>(R, n, lane);
}

// ------------------------------------------------------------------------------------------------
// Level B of the walk: a frame (compacted list of the n <= kMaxLB terms still commuting with the group, their
// residues mod the group) whose residues are LINEARLY INDEPENDENT over GF(2) cannot produce a dependent
// term anywhere below it: no alive term can fall into a descendant group and no two can merge into one
// coset.  Every descendant is then "group + a clique of the commutation graph of the list", its members are
// exactly its generators and E_min = E_frame - sum |w|.  The warp (1) certifies independence by Gaussian
// elimination across lanes (pivot row broadcast with shuffles, lowest-set-bit pivot, XOR elimination),
// (2) builds the candidates' commutation bit-matrix with POPC, then (3) the lanes walk the cliques privately with
// 32- or 64-bit masks, work items dealt from a shared list.  (In its final form the certificate is taken over
// the candidates only, see level_b.)  Returns 0 (nothing counted) if the frame cannot be certified.
constexpr int kMaxLB = 96;          // entries of a level-B frame (three rows per lane); its candidates: <= 64 (the walk's mask width)
constexpr int kFoldCand = 52;       // multi-word rows: the 64-bit image is tried up to this many candidates (it fails with probability ~ entries * 2^(candidates - 64))
constexpr int kMaxDep = 6;
constexpr float kSplitCliques = 2500.f;   // expected cliques from which a root's children go to the next k_dfs round (multi-shard runs)

SGS_D int ctz_mask(uint32_t v) { return __ffs((int)v) - 1; }
SGS_D int ctz_mask(uint64_t v) { return __ffsll((long long)v) - 1; }
SGS_D int popc_mask(uint32_t v) { return __popc(v); }
SGS_D int popc_mask(uint64_t v) { return __popcll(v); }

// is clique A (a bit mask over the list positions) before clique B in the include-first leaf order?  The sorted
// vertex sequences agree below the lowest differing bit; the clique that HAS that bit continues with a smaller
// vertex, or continues where the other has ended ("end of tuple" sorts last): it comes first.
template <typename M>
SGS_D bool clique_before(M A, M B) { const M x = A ^ B; return (A & (x & ((M)0 - x))) != 0; }

// The clique walk of a certified frame (vertices 0..ncand-1 = list positions nlow..n-1, ADm[v] = neighbours of v
// with a higher index).  Work items are the EDGES (f, l) of the candidates' commutation graph — "every clique whose
// two lowest vertices are f < l" — listed in shared memory (EL[e] = l | f << 8, l ascending: heavy items first;
// kMaxEdges per chunk, a frame with more edges is walked chunk by chunk) and dealt DYNAMICALLY: whenever lanes are
// idle at the top of an iteration they take the next items off the list, numbered by a ballot (the list cursor is a
// warp-uniform register: no atomics).  An edge's subtree holds 3-4 cliques on average, so the lanes stay level where
// a deal by vertex leaves three quarters of them idle (measured: 7 of 32 lanes busy per iteration).
//
// The loop is WARP-CONVERGED: all 32 lanes run the same iteration (take an item if idle, at most one pop, then one
// vertex) until the list is empty and the last lane has finished; the branches inside are short and rejoin.  (Left
// to themselves the lanes drift apart through the pop / leaf / push branches and the walk issues with 3-6 active
// lanes per instruction: ncu, profiles/r02_before_k_dfs_config4_source_top.csv.)  The two innermost frames live in
// registers (c0/w0: frame d, c1/w1: frame d - 1), the local-memory stack holds frames 0..d-2.  A frame whose last
// vertex is being expanded is not kept (rest == 0: nothing to come back to), so a pop always lands on a frame with
// work.  The member mask needs no stack: the members above l are added in increasing order and all lie below the
// lowest remaining candidate of the frame a pop returns to, except the vertex that frame had just expanded — the
// highest of them.  Candidates per rank are counted in byte fields of a register (clique sizes 3-6; larger ones go
// straight to hc[]) and flushed every 255 iterations.  Keeps the heaviest clique; ties go to the clique that comes
// first in the leaf order (clique_before), so neither the deal nor the order of the walk matters.
// PRUNE (branch and bound): SW[v] = sum of the weights of the vertices >= v; a branch whose best conceivable
// weight (current + every vertex from its lowest remaining candidate on) stays below `need` is not entered.
constexpr int kMaxEdges = kMaxEdgesSmem;
SGS_D int clz_mask(uint32_t v) { return __clz((int)v); }
SGS_D int clz_mask(uint64_t v) { return __clzll((long long)v); }

template <typename M, bool PRUNE>
SGS_D void lb_walk_edges(const M *ADm, const double *WAc, const uint16_t *EL, int nedges, int depth, uint32_t *hc, double &bw, M &bm,
                         uint32_t ltmask, const double *SW, double need)
{
    constexpr int kBits = (int)sizeof(M) * 8;
    M cand[kBits];
    double ws[kBits];
    int d = 0, next = 0, iter = 0;
    M c0 = 0, m0 = 0, c1 = 0;
    uint32_t cnt = 0;
    double w0 = 0.0, w1 = 0.0;
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 4; k++) hc[depth + 3 + k] += (cnt >> (8 * k)) & 255u;
        cnt = 0;
    };
#pragma unroll 1
    for (;;) {
        const bool idle = c0 == 0 && d == 0;
        const unsigned nb = __ballot_sync(0xffffffffu, idle);
        if (nb) {
            if (next >= nedges) { if (nb == 0xffffffffu) break; }
            else {
                const int e = next + __popc(nb & ltmask);
                next += __popc(nb);
                if (idle && e < nedges) {
                    const unsigned it = EL[e];
                    const int l = it & 255, f = it >> 8;
                    c0 = ADm[l] & ADm[f];
                    w0 = WAc[l] + WAc[f];
                    m0 = ((M)1 << l) | ((M)1 << f);
                    if (c0 == 0 && w0 >= bw && (w0 > bw || clique_before(m0, bm))) { bw = w0; bm = m0; }
                }
            }
        }
        if (c0 == 0 && d > 0) {
            // pop: the frame below has candidates left (a frame is only kept with rest != 0)
            d--;
            c0 = c1; w0 = w1;
            m0 &= (c0 & ((M)0 - c0)) - (M)1;                               // members below the frame's lowest remaining candidate ...
            m0 &= ~((M)1 << (kBits - 1 - clz_mask(m0)));                   // ... minus the highest: the vertex expanded last
            if (d >= 1) { c1 = cand[d - 1]; w1 = ws[d - 1]; }
        }
        if (c0 != 0) {
            const int v = ctz_mask(c0);
            const M rest = c0 & (c0 - 1);
            const M nc = rest & ADm[v];
            const double nw = w0 + WAc[v];
            const int sz = popc_mask(m0);                      // the new clique has sz + 1 vertices
            if (sz <= 5) cnt += 1u << (8 * (sz - 2));
            else hc[depth + sz + 1] += 1;
            if (nc == 0) {
                if (nw >= bw) {
                    const M cur = m0 | ((M)1 << v);
                    if (nw > bw || clique_before(cur, bm)) { bw = nw; bm = cur; }
                }
                c0 = rest;
            } else if (PRUNE && nw + SW[ctz_mask(nc)] < need) {
                c0 = rest;       // nothing below can reach the incumbent; the clique itself cannot either (it is lighter still)
            } else {
                if (rest != 0) {
                    if (d >= 1) { cand[d - 1] = c1; ws[d - 1] = w1; }
                    c1 = rest; w1 = w0;
                    d++;
                }
                c0 = nc; w0 = nw; m0 |= (M)1 << v;
            }
        }
        if (++iter == 255) { flush(); iter = 0; }
    }
    flush();
}

// One certified frame: list the edges of the candidates' graph (vertex l = lane and lane + 32 lists its lower
// neighbours dn0 / dn1) in chunks of at most kMaxEdges and walk every chunk.  Singletons and pairs are counted
// here; a singleton is offered as the best clique (it is, if its vertex is isolated).
template <typename M, bool PRUNE>
SGS_D void lb_walk_frame(const M *ADm, const double *WAc, uint16_t *EL, int ncand, M dn0, M dn1, int depth, uint32_t *hc, double &bw, M &bm,
                         int lane, const double *SW, double need)
{
    const int cnt0 = popc_mask(dn0), cnt1 = popc_mask(dn1);
    int inc0 = cnt0, inc1 = cnt1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y0 = __shfl_up_sync(0xffffffffu, inc0, o), y1 = __shfl_up_sync(0xffffffffu, inc1, o);
        if (lane >= o) { inc0 += y0; inc1 += y1; }
    }
    const int tot0 = __shfl_sync(0xffffffffu, inc0, 31), total = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
    const int ex0 = inc0 - cnt0, ex1 = tot0 + inc1 - cnt1;       // offset of my vertices' first edge in the full list
    if (lane == 0) { hc[depth + 1] += (uint32_t)ncand; hc[depth + 2] += (uint32_t)total; }
    if (lane < ncand) { bw = WAc[lane]; bm = (M)1 << lane; }
    if (lane + 32 < ncand && WAc[lane + 32] > bw) { bw = WAc[lane + 32]; bm = (M)1 << (lane + 32); }
    const uint32_t ltmask = (1u << lane) - 1u;
    int start = 0;
#pragma unroll 1
    while (start < total) {
        // the chunk: every vertex whose edges fit completely (offsets are monotone: a prefix of the remaining vertices)
        const bool in0 = cnt0 > 0 && ex0 >= start && ex0 + cnt0 - start <= kMaxEdges;
        const bool in1 = cnt1 > 0 && ex1 >= start && ex1 + cnt1 - start <= kMaxEdges;
        const int end = __reduce_max_sync(0xffffffffu, max(in0 ? ex0 + cnt0 : start, in1 ? ex1 + cnt1 : start));
        if (in0) { int o = ex0 - start; for (M x = dn0; x; x &= x - 1) EL[o++] = (uint16_t)(lane | (ctz_mask(x) << 8)); }
        if (in1) { int o = ex1 - start; for (M x = dn1; x; x &= x - 1) EL[o++] = (uint16_t)((lane + 32) | (ctz_mask(x) << 8)); }
        __syncwarp();
        lb_walk_edges<M, PRUNE>(ADm, WAc, EL, end - start, depth, hc, bw, bm, ltmask, SW, need);
        __syncwarp();
        start = end;
    }
}

// The certificate of level B on one-word rows (row i in p0 of lane i, row i + 32 in p1 of lane i): are the CANDIDATE
// rows (positions nlow..n-1) linearly dependent, or does a LOWER row (positions 0..nlow-1) lie in their span?
// Forward elimination across the lanes with the candidates as the only pivots — the lower rows are reduced along
// and must stay non-zero — two pivots per step so that the shuffle -> lowest bit -> test -> XOR chains of
// consecutive pivots overlap (the loop is latency-bound).  Half the pivots of an elimination of all n rows.
SGS_D bool rows_dependent_1w(uint64_t p0, uint64_t p1, uint64_t p2, int n, int nlow, int lane)
{
    const bool two = n > 32, three = n > 64;
    const bool low0 = lane < nlow, low1 = lane + 32 < nlow, low2 = lane + 64 < nlow;
    int u = nlow;
#pragma unroll 1
    for (; u + 1 < n; u += 2) {
        const uint64_t c0 = __shfl_sync(0xffffffffu, u < 32 ? p0 : (u < 64 ? p1 : p2), u & 31);
        uint64_t c1 = __shfl_sync(0xffffffffu, u + 1 < 32 ? p0 : (u + 1 < 64 ? p1 : p2), (u + 1) & 31);
        if (c0 == 0) return true;
        const uint64_t l0 = c0 & (0 - c0);
        if (c1 & l0) c1 ^= c0;
        if (c1 == 0) return true;
        const uint64_t l1 = c1 & (0 - c1);
        if (lane > u + 1 || low0) {
            if (p0 & l0) p0 ^= c0;
            if (p0 & l1) p0 ^= c1;
        }
        if (two && (lane + 32 > u + 1 || low1)) {
            if (p1 & l0) p1 ^= c0;
            if (p1 & l1) p1 ^= c1;
        }
        if (three && (lane + 64 > u + 1 || low2)) {
            if (p2 & l0) p2 ^= c0;
            if (p2 & l1) p2 ^= c1;
        }
    }
    if (u < n) {
        const uint64_t c0 = __shfl_sync(0xffffffffu, u < 32 ? p0 : (u < 64 ? p1 : p2), u & 31);
        if (c0 == 0) return true;
        const uint64_t l0 = c0 & (0 - c0);
        if (low0 && (p0 & l0)) p0 ^= c0;
        if (low1 && (p1 & l0)) p1 ^= c0;
        if (low2 && (p2 & l0)) p2 ^= c0;
    }
    return __any_sync(0xffffffffu, (low0 && p0 == 0) || (low1 && p1 == 0) || (low2 && p2 == 0));
}

// expected number of cliques of a graph with n vertices and e edges (edge density p): sum_k C(n,k) p^(k(k-1)/2)
SGS_D float expected_cliques(int n, int e)
{
    const float p = n > 1 ? (float)e / (0.5f * (float)n * (float)(n - 1)) : 0.f;
    float term = (float)n, sum = term, pk = 1.f;
    for (int k = 2; k <= n && k <= 24; k++) {
        pk *= p;                                   // p^(k-1)
        term *= (float)(n - k + 1) / (float)k * pk;
        sum += term;
        if (term < 0.5f) break;
    }
    return sum;
}

// cold halves of level_b's certificate, kept out of line (instruction-cache footprint of the hot kernel):
// the full-width elimination of multi-word rows, same scheme as rows_dependent_1w ...
template <int W>
__device__ __noinline__ bool rows_dependent_full(const uint64_t *R, int n, int nlow, int lane)
{
    const bool two = n > 32, three = n > 64;
    uint64_t e0[W], e1[W], e2[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        e0[k] = lane < n ? R[lane * W + k] : 0;
        e1[k] = lane + 32 < n ? R[(lane + 32) * W + k] : 0;
        e2[k] = lane + 64 < n ? R[(lane + 64) * W + k] : 0;
    }
#pragma unroll 1
    for (int u = nlow; u < n; u++) {
        uint64_t cu[W], lm[W];
        bool nz = false;
#pragma unroll
        for (int k = 0; k < W; k++) {
            cu[k] = __shfl_sync(0xffffffffu, u < 32 ? e0[k] : (u < 64 ? e1[k] : e2[k]), u & 31);
            lm[k] = nz ? 0 : (cu[k] & (0 - cu[k]));
            nz |= cu[k] != 0;
        }
        if (!nz) return true;
        uint64_t h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
        for (int k = 0; k < W; k++) { h0 |= e0[k] & lm[k]; h1 |= e1[k] & lm[k]; h2 |= e2[k] & lm[k]; }
        if ((lane > u || lane < nlow) && h0) {
#pragma unroll
            for (int k = 0; k < W; k++) e0[k] ^= cu[k];
        }
        if (two && (lane + 32 > u || lane + 32 < nlow) && h1) {
#pragma unroll
            for (int k = 0; k < W; k++) e1[k] ^= cu[k];
        }
        if (three && (lane + 64 > u || lane + 64 < nlow) && h2) {
#pragma unroll
            for (int k = 0; k < W; k++) e2[k] ^= cu[k];
        }
    }
    uint64_t z0 = 0, z1 = 0, z2 = 0;
#pragma unroll
    for (int k = 0; k < W; k++) { z0 |= e0[k]; z1 |= e1[k]; z2 |= e2[k]; }
    return __any_sync(0xffffffffu, (lane < nlow && z0 == 0) || (lane + 32 < nlow && z1 == 0) || (lane + 64 < nlow && z2 == 0));
}

// ... and the analysis of a frame whose certificate failed: an elimination of ALL n residues that tracks which
// original residues are XORed into each row yields the dependencies themselves.  A dependent term can only ever
// appear below the frame if the support of some kernel vector (XOR of dependencies) is a CLIQUE of the commutation
// graph: the terms of a dependency inside one group commute pairwise.  All 2^ndep - 1 of them are checked (at most
// kMaxDep dependencies).  Returns true if the frame is certified after all.
template <int W>
__device__ __forceinline__ bool level_b_dependent_body(const uint64_t *R, int n, int lane)
{
    const bool two = n > 32;
    // commutation rows of all n entries (the form is symmetric: popc(swap(a) & b) = popc(a & swap(b)))
    uint64_t adj0 = 0, adj1 = 0;
    {
        uint64_t s0[W], s1[W];
#pragma unroll
        for (int k = 0; k < W; k++) {
            s0[k] = lane < n ? swapzx(R[lane * W + k]) : 0;
            s1[k] = lane + 32 < n ? swapzx(R[(lane + 32) * W + k]) : 0;
        }
#pragma unroll 1
        for (int u = 0; u < n; u++) {
            uint64_t ru[W];
#pragma unroll
            for (int k = 0; k < W; k++) ru[k] = R[u * W + k];
            adj0 |= (uint64_t)(row_anticommute_sw<W>(ru, s0) ^ 1) << u;
            if (two) adj1 |= (uint64_t)(row_anticommute_sw<W>(ru, s1) ^ 1) << u;
        }
        const uint64_t vmask = n >= 64 ? ~0ULL : ((1ULL << n) - 1ULL);
        adj0 = lane < n ? (adj0 & vmask & ~(1ULL << lane)) : 0;
        adj1 = lane + 32 < n ? (adj1 & vmask & ~(1ULL << (lane + 32))) : 0;
    }
    uint64_t e0[W], e1[W];
    uint64_t cmb0 = lane < n ? (1ULL << lane) : 0, cmb1 = lane + 32 < n ? (1ULL << (lane + 32)) : 0;
#pragma unroll
    for (int k = 0; k < W; k++) {
        e0[k] = lane < n ? R[lane * W + k] : 0;
        e1[k] = lane + 32 < n ? R[(lane + 32) * W + k] : 0;
    }
    uint64_t deps[kMaxDep];
    int ndep = 0;
#pragma unroll 1
    for (int u = 0; u < n; u++) {
        uint64_t cu[W], lm[W];
        bool nz = false;
#pragma unroll
        for (int k = 0; k < W; k++) {
            cu[k] = __shfl_sync(0xffffffffu, u < 32 ? e0[k] : e1[k], u & 31);
            lm[k] = nz ? 0 : (cu[k] & (0 - cu[k]));
            nz |= cu[k] != 0;
        }
        const uint64_t cmu = __shfl_sync(0xffffffffu, u < 32 ? cmb0 : cmb1, u & 31);
        if (!nz) {
            if (ndep == kMaxDep) return false;
#pragma unroll
            for (int k = 0; k < kMaxDep; k++) if (k == ndep) deps[k] = cmu;
            ndep++;
            continue;
        }
        uint64_t h0 = 0, h1 = 0;
#pragma unroll
        for (int k = 0; k < W; k++) { h0 |= e0[k] & lm[k]; h1 |= e1[k] & lm[k]; }
        if (lane > u && h0) {
#pragma unroll
            for (int k = 0; k < W; k++) e0[k] ^= cu[k];
            cmb0 ^= cmu;
        }
        if (lane + 32 > u && h1) {
#pragma unroll
            for (int k = 0; k < W; k++) e1[k] ^= cu[k];
            cmb1 ^= cmu;
        }
    }
#pragma unroll 1
    for (int gcode = 1; gcode < (1 << ndep); gcode++) {
        uint64_t S = 0;
#pragma unroll
        for (int k = 0; k < kMaxDep; k++) if ((gcode >> k) & 1) S ^= deps[k];
        const bool in0 = (S >> lane) & 1, in1 = (S >> (lane + 32)) & 1;
        const bool bad = (in0 && (S & ~adj0 & ~(1ULL << lane)) != 0) || (in1 && (S & ~adj1 & ~(1ULL << (lane + 32))) != 0);
        if (!__any_sync(0xffffffffu, bad)) return false;          // a commuting dependency: not certified
    }
    return true;
}

// Multi-word rows: out of line (their hot code is twice as long and spills out of the instruction caches: -13 %
// when this sits inside level_b); one-word rows: inline (measured 6 % faster there).
template <int W>
__device__ __noinline__ bool level_b_dependent_cold(const uint64_t *R, int n, int lane) { return level_b_dependent_body<W>(R, n, lane); }
template <int W>
SGS_D bool level_b_dependent(const uint64_t *R, int n, int lane)
{
    if (W == 1) return level_b_dependent_body<W>(R, n, lane);
    return level_b_dependent_cold<W>(R, n, lane);
}

// (not inlined: it is called from two places of k_dfs and a single copy keeps the hot loops inside the
//  instruction cache)
template <int W, bool PRUNE>
__device__ __noinline__ int level_b(const uint64_t *R, const uint16_t *TTc, const double *__restrict__ wabs, double *WC, uint64_t *AD, uint32_t *AD32, uint16_t *EL, int n, int nlow, int depth,
                  uint32_t *hc, double &bw, uint64_t &bm, int lane, float split_cliques, double *SW, double need, uint32_t *pc)
{
    // 1. the certificate: the residues of the CANDIDATES (positions nlow..n-1) are linearly independent and no
    //    lower-index entry lies in their span.  Then, for every clique K of candidates, the group <G, K> has rank
    //    |K| more, holds no alive term besides K (such a term's residue would lie in span(K)) and every step of
    //    its canonical generating sequence is valid (a lower-index term in the coset of a new generator would lie
    //    in the span of K as well).  Dependencies among the lower entries themselves are irrelevant: they can
    //    never become generators below this frame.  Own rows: entries lane and lane + 32; step u broadcasts the
    //    (already reduced) copy of candidate u and clears its lowest set bit from the later candidates and from
    //    the lower rows.
    // (the candidates' weights |w| for the walk: gathered from the term table into the warp's scratch while the
    //  certificate runs — the lists carry no weights)
    if (lane < n - nlow) WC[lane] = wabs[TTc[lane]];
    if (lane + 32 < n - nlow) WC[lane + 32] = wabs[TTc[lane + 32]];
    uint64_t e0[W], e1[W], e2[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        e0[k] = lane < n ? R[lane * W + k] : 0;
        e1[k] = lane + 32 < n ? R[(lane + 32) * W + k] : 0;
        e2[k] = lane + 64 < n ? R[(lane + 64) * W + k] : 0;
    }
    bool dependent = false;
    // Multi-word rows: if the images under a linear map onto 64 bits (the words folded together by rotations)
    // are independent, so are the rows — one-word elimination at half the cost or less; with <= 62 entries it
    // rarely fails, and when it does the full-width elimination below decides.
    bool proven = false;
    if (W > 1 && n - nlow <= kFoldCand) {
        uint64_t p0 = e0[0], p1 = e1[0], p2 = e2[0];
#pragma unroll
        for (int k = 1; k < W; k++) {
            const int rot = (23 * k) & 63;
            p0 ^= (e0[k] << rot) | (e0[k] >> (64 - rot));
            p1 ^= (e1[k] << rot) | (e1[k] >> (64 - rot));
            p2 ^= (e2[k] << rot) | (e2[k] >> (64 - rot));
        }
        const bool dep = rows_dependent_1w(p0, p1, p2, n, nlow, lane);
        proven = !dep;
        if (lane == 0) pc[dep ? kPathFoldFull : kPathFoldOk]++;
    }
    if (W == 1) {
        dependent = rows_dependent_1w(e0[0], e1[0], e2[0], n, nlow, lane);
        proven = true;
    }
    if (!proven) dependent = rows_dependent_full<W>(R, n, nlow, lane);      // (multi-word rows whose folded image was inconclusive)
    if (dependent) {
        // rare (a handful per thousand frames on random instances): out of line, the hot code must stay inside the instruction caches
        // (the dependency analysis works on 64-bit entry masks: a longer list goes back to the caller uncertified)
        if (n > 64 || !level_b_dependent<W>(R, n, lane)) { if (lane == 0) pc[kPathNoCert]++; return 0; }
        if (lane == 0) pc[kPathDepCert]++;
    }
    // 2. commutation graph of the candidates only (list positions nlow..n-1; vertex = position - nlow): lane
    //    owns vertices lane and lane + 32, keeps the higher-numbered neighbours (cliques grow upward)
    const int ncand = n - nlow;
    const uint64_t *Rc = R + (size_t)nlow * W;
    if (PRUNE) {
        // certified: below this frame members = generators, so no group can weigh more than all candidates
        // together.  Suffix sums of the candidate weights (warp scan), then the frame-level bound.
        __syncwarp();
        double x1 = lane + 32 < ncand ? WC[lane + 32] : 0.0, x0 = lane < ncand ? WC[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y1 = __shfl_down_sync(0xffffffffu, x1, o), y0 = __shfl_down_sync(0xffffffffu, x0, o);
            if (lane + o < 32) { x1 += y1; x0 += y0; }
        }
        x0 += __shfl_sync(0xffffffffu, x1, 0);
        if (__shfl_sync(0xffffffffu, x0, 0) < need) return 1;          // the whole frame is out of reach
        SW[lane] = x0; SW[lane + 32] = x1;
        __syncwarp();
    }
#ifndef SGS_W2_ONE_WALK
#define SGS_W2_ONE_WALK 0
#endif
    if (ncand <= 32 && !(SGS_W2_ONE_WALK && W > 1)) {
        uint64_t sc[W];
#pragma unroll
        for (int k = 0; k < W; k++) sc[k] = lane < ncand ? swapzx(Rc[lane * W + k]) : 0;
        uint32_t a32 = 0;
#pragma unroll 1
        for (int u = 0; u < ncand; u++) {
            uint64_t ru[W];
#pragma unroll
            for (int k = 0; k < W; k++) ru[k] = Rc[u * W + k];
            a32 |= (uint32_t)(row_anticommute_sw<W>(ru, sc) ^ 1) << u;
        }
        const uint32_t up32 = lane < ncand ? (a32 & ~((2u << lane) - 1u)) : 0u;
        if (split_cliques > 0.f && ncand >= 12) {      // too big a walk for one warp of a short kernel: let the caller split it
            int e = __popc(up32);
            for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if (expected_cliques(ncand, e) > split_cliques) return 2;
        }
        if (lane < ncand) AD32[lane] = up32;
        if (lane == 0) { pc[kPathWalk32]++; if (n > 64) pc[kPathWide]++; }
        __syncwarp();
        // (bw / bm are references into the caller's frame, i.e. local memory: the walk works on register copies)
        uint32_t bm32 = 0;
        double bwl = -1.0;
        lb_walk_frame<uint32_t, PRUNE>(AD32, WC, EL, ncand, lane < ncand ? (a32 & ((1u << lane) - 1u)) : 0u, 0u, depth, hc, bwl, bm32, lane, SW, need);
        bm = (uint64_t)bm32;
        bw = bwl;
    } else {
        uint64_t sc0[W], sc1[W];
#pragma unroll
        for (int k = 0; k < W; k++) {
            sc0[k] = lane < ncand ? swapzx(Rc[lane * W + k]) : 0;
            sc1[k] = lane + 32 < ncand ? swapzx(Rc[(lane + 32) * W + k]) : 0;
        }
        uint64_t a0 = 0, a1 = 0;
#pragma unroll 1
        for (int u = 0; u < ncand; u++) {
            uint64_t ru[W];
#pragma unroll
            for (int k = 0; k < W; k++) ru[k] = Rc[u * W + k];
            a0 |= (uint64_t)(row_anticommute_sw<W>(ru, sc0) ^ 1) << u;
            a1 |= (uint64_t)(row_anticommute_sw<W>(ru, sc1) ^ 1) << u;
        }
        if (lane >= ncand) a0 = 0;
        const uint64_t up0 = a0 & ~((2ULL << lane) - 1ULL);
        const uint64_t up1 = (lane + 32 < ncand && lane != 31) ? (a1 & ~((2ULL << (lane + 32)) - 1ULL)) : 0ULL;
        if (split_cliques > 0.f) {
            int e = popc64(up0) + popc64(up1);
            for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if (expected_cliques(ncand, e) > split_cliques) return 2;
        }
        AD[lane] = up0;
        if (lane + 32 < ncand) AD[lane + 32] = up1;
        if (lane == 0) { pc[kPathWalk64]++; if (n > 64) pc[kPathWide]++; }
        __syncwarp();
        uint64_t bmr = 0;
        double bwl = -1.0;
        lb_walk_frame<uint64_t, PRUNE>(AD, WC, EL, ncand, a0 & ((1ULL << lane) - 1ULL), lane + 32 < ncand ? (a1 & ((1ULL << (lane + 32)) - 1ULL)) : 0ULL,
                                       depth, hc, bwl, bmr, lane, SW, need);
        bm = bmr;
        bw = bwl;
    }
    __syncwarp();
    return 1;
}

// cold half of try_best: ties by tuple order, then the copy of the tuple
__device__ __noinline__ bool best_update(double Eval, int len, double bestE, int bestM, uint16_t *bestG, const uint16_t *path, int lane)
{
    __syncwarp();
    if (Eval == bestE && !tuple_before(path, len, bestG, bestM)) return false;
    for (int i = lane; i < len; i += 32) bestG[i] = path[i];
    __syncwarp();
    return true;
}

// cold tail of a level-B walk whose best clique reaches the incumbent: warp argmax of (weight, then the clique that
// comes first in the leaf order), the clique's terms appended to path[depth..).  Returns the clique size and weight.
__device__ __noinline__ int level_b_pick(double bw, uint64_t bm, const uint16_t *TTlb, uint16_t *path, int depth, int lane, double *wout)
{
    double w = bm ? bw : -1.0;
    uint64_t mask = bm;
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) {
        const double ow = __shfl_xor_sync(0xffffffffu, w, d);
        const uint64_t om = __shfl_xor_sync(0xffffffffu, mask, d);
        if (ow > w || (ow == w && om != mask && clique_before(om, mask))) { w = ow; mask = om; }
    }
    if (lane == 0) {
        uint64_t mm = mask;
        int l2 = depth;
        while (mm) { path[l2++] = TTlb[ctz64(mm)]; mm &= mm - 1; }
    }
    __syncwarp();
    *wout = w;
    return popc64(mask);
}

// launch geometry: blocks of kDfsWpb warps (the warps of a block share nothing: small blocks only make the shared-
// memory granularity finer); kDfsMinBlocks caps the registers (65536 / (32 * kDfsWpb * kDfsMinBlocks) per thread)
#ifndef SGS_DFS_WPB
#define SGS_DFS_WPB 4
#endif
#ifndef SGS_DFS_MINB
#define SGS_DFS_MINB 6
#endif
#ifndef SGS_DFS_MINB1
#define SGS_DFS_MINB1 8
#endif
constexpr int kDfsWpb = SGS_DFS_WPB;
enum { kDfsOwnSelf = 1, kDfsExpand = 2 };
// one-word rows: 64 registers (8 blocks of 4 warps per SM) measured 2 % faster than 80; multi-word rows spill there (-13 %)
template <int W> constexpr int dfs_min_blocks() { return W == 1 ? SGS_DFS_MINB1 : SGS_DFS_MINB; }

template <int W, bool SPLIT, bool PRUNE>
__global__ void __launch_bounds__(32 * kDfsWpb, dfs_min_blocks<W>())
k_dfs(SparseDev D, const uint16_t *__restrict__ items, unsigned nitems, unsigned int *item_counter,
      int shard, int nshards, int cap, int smem_per_warp, const unsigned int *nitems_dev, const unsigned int *__restrict__ perm,
      uint16_t *retry_out_, unsigned int *retry_count, unsigned retry_cap, const unsigned int *valid_end, int mode)
{
    // mode bit 0 (kDfsOwnSelf): with several shards a root belongs to the shard of its OWN tuple (expansion launch);
    // mode bit 1 (kDfsExpand): expansion launch — a root is counted and evaluated, its valid children are written to
    // retry_out as the roots of the walk; nothing is walked here.  This is the last level of the top-of-tree expansion
    // done incrementally (one alive list, one duplicate check per parent) instead of from scratch per parent by k_general.
    uint16_t *const retry_out = SPLIT ? retry_out_ : nullptr;      // single-shard build: the hand-off code is compiled out
    // retry_out != null: a root whose residues cannot be certified (a commuting dependency: its subtree needs
    // the warp-cooperative level A, hundreds of instructions per candidate) is not walked here; its children
    // are appended to retry_out as roots of the next k_dfs round, so that one such subtree is shared by many
    // warps instead of pinning one warp for up to a millisecond.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const unsigned ltmask = lanemask_lt();
    if (nitems_dev) nitems = min(nitems, *nitems_dev);   // queue length produced on the device (parent-sharded expansion, retry round)
    if (valid_end) nitems = min(nitems, *valid_end);
    unsigned char *p = smem_raw + (size_t)warp * smem_per_warp;
    uint64_t *R = (uint64_t *)p;        p += (size_t)cap * W * 8;
    double *WC = (double *)p;           p += 64 * 8;      // |w| of the candidates of the level-B frame being walked
    uint64_t *B = (uint64_t *)p;        // root basis; dead once the root's residues are written, then:
    uint64_t *AD = (uint64_t *)p;       // level-B commutation bit-matrix, 64-bit form
    uint32_t *AD32 = (uint32_t *)(p + 64 * 8);   // ... and 32-bit form
    p += dfs_scratch_bytes<W>();
    uint16_t *EL = (uint16_t *)p;       p += (size_t)kMaxEdgesSmem * 2;
    double *fE = (double *)p;           p += (size_t)kMaxFramesA * 8;
    int *fbase = (int *)p;              p += (size_t)kMaxFramesA * 4;
    int *fa = (int *)p;                 p += (size_t)kMaxFramesA * 4;
    int *fq = (int *)p;                 p += (size_t)kMaxFramesA * 4;
    uint32_t *hist = (uint32_t *)p;     p += (size_t)(kMaxG + 1) * 4 + 4;
    uint16_t *TT = (uint16_t *)p;       p += (size_t)cap * 2;
    uint16_t *path = (uint16_t *)p;     p += (size_t)kMaxG * 2;
    uint16_t *bestG = (uint16_t *)p;    p += (size_t)kMaxG * 2;
    uint8_t *pivB = (uint8_t *)p;       p += 64;
    p = smem_raw + (((size_t)(p - smem_raw) + 7) & ~(size_t)7);
    double *SW = (double *)p;           // [64] suffix sums of the candidate weights (branch-and-bound build only)
    if (PRUNE) p += 64 * 8;
    uint32_t *pc = (uint32_t *)p;       // [kNumPaths] code-path counters (kPath*)
    if (lane < kNumPaths) pc[lane] = 0;

    const unsigned gw = blockIdx.x * wpb + warp;
    BestRec *myrec = D.wbest + gw;
    double bestE = myrec->valid ? myrec->E : 1e300;     // 1e300: no incumbent
    int bestM = myrec->valid ? myrec->m : 0;
    for (int i = lane; i < bestM; i += 32) bestG[i] = myrec->g[i];
    for (int i = lane; i <= kMaxG; i += 32) hist[i] = 0;
    __syncwarp();

    // record path[0..len) as the best tuple if Eval beats the incumbent (ties: tuple order)
    double thr = 1e300;                                   // branch and bound: see prune_threshold()
    auto try_best = [&](double Eval, int len) {
        if (Eval > bestE) return;
        if (best_update(Eval, len, bestE, bestM, bestG, path, lane)) {
            bestE = Eval; bestM = len;
            if (PRUNE) {
                if (lane == 0) atomicMin(D.inc, enc_double(Eval));
                thr = fmin(thr, prune_threshold(Eval));
            }
        }
    };

    uint32_t hc[kMaxG + 4];                              // per-lane candidate counts by rank (level B)
#pragma unroll 1
    for (int i = 0; i < kMaxG + 4; i++) hc[i] = 0;

    // run level B on the list at lb (n entries, candidates from position nlow, group rank `depth`,
    // energy Ef); false if the residues are linearly dependent
    auto run_level_b = [&](int lb, int n, int nlow, int depth, double Ef, float split) -> int {
        double bw = -1.0;
        uint64_t bm = 0;
        // a clique of weight w gives the energy Ef - w: it reaches the incumbent iff w >= Ef - thr
        const int rc = level_b<W, PRUNE>(R + lb * W, TT + lb + nlow, D.wabs, WC, AD, AD32, EL, n, nlow, depth, hc, bw, bm, lane, split, SW, Ef - thr, pc);
        if (rc != 1) return rc;
        // does any lane's best clique reach the incumbent?
        const bool cand = bm != 0 && (Ef - bw) <= bestE;
        if (__any_sync(0xffffffffu, cand)) {
            double w;
            const int sz = level_b_pick(bw, bm, TT + lb + nlow, path, depth, lane, &w);      // (bm: mask over the candidates)
            try_best(Ef - w, depth + sz);
        }
        return 1;
    };

#ifndef SGS_WALK_TRACE
#define SGS_WALK_TRACE 0      // per-root timing diagnostics (SGS_TRACE=1) cost ~5 % of the walk: opt-in at build time
#endif
    const bool trace = SGS_WALK_TRACE && D.trace != 0;
    unsigned long long troot = 0;
    bool tfail = false;
    unsigned tinfo = 0;
    const unsigned tround = valid_end ? (retry_out ? 2u : 3u) : 1u;
    const unsigned long long twarp0 = trace ? gtimer() : 0ull;
#pragma unroll 1
    for (;;) {
        // ---- fetch the next subtree root owned by this shard ----
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(item_counter, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (trace) {
            const unsigned long long tn = gtimer();
            if (troot) {
                const unsigned long long dt = tn - troot;
                if (lane == 0) {
                    atomicMax(&D.stats->dbg[7], (dt << 24) | ((unsigned long long)tinfo << 1) | (tfail ? 1ull : 0ull));
                    atomicMax(&D.stats->dbg[0], dt);
                    atomicAdd(&D.stats->dbg[3], dt);
                    if (tfail) { atomicAdd(&D.stats->dbg[1], 1ull); atomicAdd(&D.stats->dbg[2], dt); }
                    if (dt > 50000ull) { atomicAdd(&D.stats->dbg[5], 1ull); atomicAdd(&D.stats->dbg[6], dt); }
                }
            }
            troot = tn; tfail = false; tinfo = 0;
        }
        if (it >= nitems) break;
        // (a limit was hit somewhere: the search is void, stop early — polled every 16th root: the flag is an L2 round trip
        //  in front of every root's dependent loads)
        if ((it & 15u) == 0 && *(volatile unsigned int *)&D.stats->error) break;
        if (PRUNE) thr = fmin(prune_threshold(bestE), prune_threshold(dec_double(*(volatile unsigned long long *)D.inc)));
        const uint16_t *rec = items + (size_t)(perm ? perm[it] : it) * D.S;
        const int m0 = rec[0];
        const unsigned flags = rec[1];
        // a subtree root belongs to the shard of its PARENT (hash of the tuple without its last generator): the same
        // partition whether the last expansion level was replicated (first, adaptive run of a plan) or expanded by
        // the parent's owner only (recorded launch sequence)
        if (nshards > 1 && (int)(tuple_hash(rec + 2, (mode & kDfsOwnSelf) ? m0 : (m0 > 0 ? m0 - 1 : 0)) % (unsigned)nshards) != shard) continue;
        __syncwarp();
        for (int i = lane; i < m0; i += 32) path[i] = rec[2 + i];
        __syncwarp();
        if ((flags & kFlagUnclean) || m0 > 32) { push_slow(D, path, m0, flags, lane); continue; }

        // ---- root: echelon basis of the m0 generators.  Lane i owns generator i; step b broadcasts
        // the (already reduced) row b and clears its pivot bit from the later rows. ----
        uint64_t bv[W];
        double wsum = 0.0;
        if (lane < m0) {
            const int t = path[lane];
#pragma unroll
            for (int k = 0; k < W; k++) bv[k] = D.rows[t * W + k];
            wsum = D.wabs[t];
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) bv[k] = 0;
        }
#pragma unroll 1
        for (int b = 0; b < m0; b++) {
            uint64_t rb[W];
#pragma unroll
            for (int k = 0; k < W; k++) rb[k] = __shfl_sync(0xffffffffu, bv[k], b);
            const int pv = row_lowbit<W>(rb);
            if (lane == b) pivB[b] = (uint8_t)pv;
            if (lane > b && lane < m0 && row_bit<W>(bv, pv)) {
#pragma unroll
                for (int k = 0; k < W; k++) bv[k] ^= rb[k];
            }
        }
        if (lane < m0) {
#pragma unroll
            for (int k = 0; k < W; k++) B[lane * W + k] = bv[k];
        }
        __syncwarp();
        // E0 = -sum |w_g| in generator order (deterministic: same order on every shard)
        double E0 = 0.0;
#pragma unroll 1
        for (int b = 0; b < m0; b++) E0 -= __shfl_sync(0xffffffffu, wsum, b);

        int a = alive_list(D, path, m0, TT, cap, lane);
        if (a < 0) { push_slow(D, path, m0, flags, lane); continue; }
        bool member = false;
        int q0 = 0;
        double wall = 0.0;                                  // weight of all alive entries (branch-and-bound mode)
        const int core0 = m0 ? (int)path[m0 - 1] : -1;
#pragma unroll 1
        for (int i = lane; i < a; i += 32) {
            uint64_t v[W];
            const int t = TT[i];
#pragma unroll
            for (int k = 0; k < W; k++) v[k] = D.rows[t * W + k];
#pragma unroll 1
            for (int b = 0; b < m0; b++) {
                if (row_bit<W>(v, pivB[b])) {
#pragma unroll
                    for (int k = 0; k < W; k++) v[k] ^= B[b * W + k];
                }
            }
            member |= row_is_zero<W>(v);
            q0 += (t <= core0);
#pragma unroll
            for (int k = 0; k < W; k++) R[i * W + k] = v[k];
            if (PRUNE) wall += D.wabs[t];
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, member)) { push_slow(D, path, m0, flags | kFlagUnclean, lane); continue; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q0 += __shfl_xor_sync(0xffffffffu, q0, d);

        if (lane == 0) hist[m0]++;
        if (PRUNE) {
            // every group below contains the root's generators plus alive terms only (as generators or as dependent
            // members), each worth at most |w|: E >= E0 - sum of all alive weights
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) wall += __shfl_xor_sync(0xffffffffu, wall, d);
            if (E0 - wall > thr) continue;
        }
        tinfo = (tround << 20) | ((unsigned)m0 << 14) | ((unsigned)min(a, 127) << 7) | (unsigned)min(a - q0, 127);
        if (q0 >= a) { try_best(E0, m0); continue; }                      // no extension: a leaf
        // (a certified root whose clique walk alone would outlast the rest of a short kernel is split the same way
        //  when a next round exists: level_b answers 2)
        if (!(mode & kDfsExpand) && a <= kMaxLB && a - q0 <= 64 && m0 + (a - q0) < kMaxG && run_level_b(0, a, q0, m0, E0, retry_out ? D.split_cliques : 0.f) == 1) continue;
        tfail = true;
        if (has_duplicate<W>(R, a, EL, lane)) {
            try_best(E0, m0);
            push_slow(D, path, m0, flags | kFlagExpandOnly, lane);
            continue;
        }
        if (retry_out && (a - q0 >= 4 || (mode & kDfsExpand)) && m0 + 1 < kMaxG) {
            // every entry from q0 on is a valid child (residues pairwise distinct); each counts itself as a root
            unsigned slot = 0;
            if (lane == 0) slot = atomicAdd(retry_count, (unsigned)(a - q0));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot + (unsigned)(a - q0) <= retry_cap) {
                if (lane == 0) pc[kPathSplit]++;
                for (int j = q0 + lane; j < a; j += 32) {
                    uint16_t *o = retry_out + (size_t)(slot + (unsigned)(j - q0)) * D.S;
                    o[0] = (uint16_t)(m0 + 1);
                    o[1] = 0;
                    for (int k = 0; k < m0; k++) o[2 + k] = path[k];
                    o[2 + m0] = TT[j];
                }
                continue;
            }
            if (lane == 0) atomicMin(retry_count + 1, slot);               // queue full: records end here; walk it locally
        }
        if (lane == 0) hist[m0 + 1] += (uint32_t)(a - q0);

        // ---- depth-first walk; the current frame (base, a, q, E) lives in registers ----
        int fd = 0, base = 0, q = q0;
        double E = E0;
#pragma unroll 1
        for (;;) {
            if (q >= a) {                                   // frame exhausted: pop
                if (fd == 0) break;
                fd--;
                base = fbase[fd]; a = fa[fd]; q = fq[fd]; E = fE[fd];
                continue;
            }
            const int jq = q++;
            // (the warp-cooperative descent is where an over-deep instance spends its time: poll the error flag)
            if ((jq & 7) == 0 && *(volatile unsigned int *)&D.stats->error) break;
            uint64_t rj[W], sw[W];
#pragma unroll
            for (int k = 0; k < W; k++) { rj[k] = R[(base + jq) * W + k]; sw[k] = swapzx(rj[k]); }
            const double Ec = E - D.wabs[TT[base + jq]];

            // pass 1: which list entries commute with the new generator (POPC), and is any above it?
            // (roomy: the child list fits behind its parent whatever its length — pass 2 counts as it writes and
            //  pass 1 is skipped; a child without extension then wrote a list for nothing, which is rare this high up)
            const int depth = m0 + fd;                      // rank of the frame's group
            const int cb = base + a;
#ifndef SGS_ROOMY_W1
#define SGS_ROOMY_W1 1
#endif
            const bool roomy = (W > 1 || SGS_ROOMY_W1) && a > 32 && cb + a <= cap && depth + 1 < kMaxG && fd + 1 < kMaxFramesA;
            int nal = 0, nlow = 0;
            unsigned hiAny = 0;
            if (roomy) {
            } else if (a <= 32) {
                bool com = false;
                if (lane < a && lane != jq) com = !row_anticommute_sw<W>(R + (base + lane) * W, sw);
                const unsigned bal = __ballot_sync(0xffffffffu, com);
                hiAny = (bal >> jq) >> 1;
                nal = __popc(bal);
                nlow = nal - __popc(hiAny);
            } else {
#pragma unroll 1
                for (int r0 = 0; r0 < a; r0 += 32) {
                    const int i = r0 + lane;
                    bool com = false;
                    if (i < a && i != jq) com = !row_anticommute_sw<W>(R + (base + i) * W, sw);
                    const unsigned bal = __ballot_sync(0xffffffffu, com);
                    const int rel = jq - r0;                // position of jq inside this row
                    if (rel >= 32) nlow += __popc(bal);
                    else if (rel >= 0) { nlow += __popc(bal & ((1u << rel) - 1u)); hiAny |= (bal >> rel) >> 1; }
                    else hiAny |= bal;
                    nal += __popc(bal);
                }
            }
            if (!roomy && hiAny == 0) {                     // the child has no extension: a leaf of the tree
                if (Ec <= bestE) {
                    if (lane == 0) path[depth] = TT[base + jq];
                    try_best(Ec, depth + 1);
                }
                continue;
            }
            if (lane == 0) path[depth] = TT[base + jq];
            if (!roomy && (cb + nal > cap || depth + 1 >= kMaxG || fd + 1 >= kMaxFramesA)) {     // out of list / frame space: the exact kernel takes over
                __syncwarp();
                try_best(Ec, depth + 1);
                push_slow(D, path, depth + 1, kFlagExpandOnly, lane);
                continue;
            }
            // pass 2: child list = surviving entries, reduced by the new generator at its pivot bit
            const int pv = row_lowbit<W>(rj);
            int off = cb;
            double csum = 0.0;                              // weight of the child's list (branch-and-bound mode)
#pragma unroll 1
            for (int r0 = 0; r0 < a; r0 += 32) {
                const int i = r0 + lane;
                bool com = false;
                uint64_t v[W];
                if (i < a && i != jq) {
#pragma unroll
                    for (int k = 0; k < W; k++) v[k] = R[(base + i) * W + k];
                    com = !row_anticommute_sw<W>(v, sw);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, com);
                if (roomy) {
                    const int rel = jq - r0;
                    if (rel >= 32) nlow += __popc(bal);
                    else if (rel >= 0) { nlow += __popc(bal & ((1u << rel) - 1u)); hiAny |= (bal >> rel) >> 1; }
                    else hiAny |= bal;
                }
                if (com) {
                    if (row_bit<W>(v, pv)) {
#pragma unroll
                        for (int k = 0; k < W; k++) v[k] ^= rj[k];
                    }
                    const int pos = off + __popc(bal & ltmask);
#pragma unroll
                    for (int k = 0; k < W; k++) R[pos * W + k] = v[k];
                    TT[pos] = TT[base + i];
                    if (PRUNE) csum += D.wabs[TT[base + i]];
                }
                off += __popc(bal);
            }
            __syncwarp();
            if (roomy) {
                nal = off - cb;
                if (hiAny == 0) {                           // (a leaf after all)
                    try_best(Ec, depth + 1);
                    continue;
                }
            }
            if (PRUNE) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
                if (Ec - csum > thr) { try_best(Ec, depth + 1); continue; }     // (the child itself is still a candidate)
            }
            if (nal <= kMaxLB && nal - nlow <= 64 && depth + 1 + (nal - nlow) < kMaxG && run_level_b(cb, nal, nlow, depth + 1, Ec, 0.f) == 1) continue;
            // residues pairwise distinct ⇒ every child of this child is a new group whose only members
            // are its generators.  Otherwise the exact kernel expands it.
            if (has_duplicate<W>(R + cb * W, nal, EL, lane)) {
                try_best(Ec, depth + 1);
                push_slow(D, path, depth + 1, kFlagExpandOnly, lane);
                continue;
            }
            if (lane == 0) {
                hist[depth + 2] += (uint32_t)(nal - nlow);
                fbase[fd] = base; fa[fd] = a; fq[fd] = q; fE[fd] = E;
                pc[kPathLevelA]++;
            }
            __syncwarp();
            fd++;
            base = cb; a = nal; q = nlow; E = Ec;
        }
    }

    __syncwarp();
    if (trace && lane == 0) {                              // warp lifetimes: idle share = 1 - sum / (warps x kernel time)
        const unsigned long long te = gtimer();
        atomicAdd(&D.stats->dbg[4], te - twarp0);
        if (tround == 1u) { D.wtime[2 * gw] = twarp0; D.wtime[2 * gw + 1] = te; }
    }
    if (bestE < 1e300) {
        if (lane == 0) { myrec->E = bestE; myrec->m = bestM; myrec->valid = 1; }
        for (int i = lane; i < bestM; i += 32) myrec->g[i] = bestG[i];
    }
#pragma unroll 1
    for (int i = 0; i <= kMaxG; i++) {
        unsigned v = hc[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) hist[i] += v;
    }
    __syncwarp();
    for (int i = lane; i <= kMaxG; i += 32)
        if (hist[i]) atomicAdd(&D.stats->hist[i], (unsigned long long)hist[i]);
    if (lane < kNumPaths && pc[lane]) atomicAdd(&D.stats->paths[lane], (unsigned long long)pc[lane]);
}

// ------------------------------------------------------------------------------------------------
// deterministic argmin over the per-warp records: minimum energy, ties by tuple order.  One block; the
// order (E, tuple) is total, so the tree reduction gives the same winner whatever the pairing.
SGS_D bool rec_better(const BestRec *recs, int a, int b)     // is record a strictly better than b?
{
    if (a < 0) return false;
    if (b < 0) return true;
    if (recs[a].E != recs[b].E) return recs[a].E < recs[b].E;
    return a != b && tuple_before(recs[a].g, recs[a].m, recs[b].g, recs[b].m);
}

// what one shard contributes to the combination: its best record and the counters that add up over shards.
// One ncclAllGather of these records is the search's only collective.
struct ShardRec {
    BestRec best;
    unsigned long long sums[kStatsSummed];
};

// block argmin over records: (E, index) pairs travel through shared memory, the records themselves are only touched
// again on exact ties (tuple order).  Returns the winner's index in every thread (-1: no valid record).
SGS_D int block_argmin(const BestRec *__restrict__ recs, int nrecs, int stride_bytes, double *sE, int *sI)
{
    const int tid = threadIdx.x;
    auto rec = [&](int i) { return (const BestRec *)((const char *)recs + (size_t)i * stride_bytes); };
    auto better = [&](double ea, int a, double eb, int b) {       // is (ea, a) strictly better than (eb, b)?
        if (a < 0) return false;
        if (b < 0) return true;
        if (ea != eb) return ea < eb;
        return a != b && tuple_before(rec(a)->g, rec(a)->m, rec(b)->g, rec(b)->m);
    };
    double e = 0.0;
    int best = -1;
    // (the records are 144 bytes apart: eight independent loads in flight per thread instead of a chain of misses)
    for (int i0 = tid; i0 < nrecs; i0 += 8 * blockDim.x) {
        double ev[8];
        int vv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + k * blockDim.x;
            vv[k] = i < nrecs ? rec(i)->valid : 0;
            ev[k] = i < nrecs ? rec(i)->E : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (vv[k] && better(ev[k], i0 + k * (int)blockDim.x, e, best)) { e = ev[k]; best = i0 + k * blockDim.x; }
    }
    sE[tid] = e; sI[tid] = best;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
        if (tid < s2 && better(sE[tid + s2], sI[tid + s2], sE[tid], sI[tid])) { sE[tid] = sE[tid + s2]; sI[tid] = sI[tid + s2]; }
        __syncthreads();
    }
    return sI[0];
}

// per-shard reduction: best of the per-warp records + this shard's counters -> one ShardRec
__global__ void __launch_bounds__(1024) k_reduce_pack(const BestRec *__restrict__ recs, int nrecs, const SparseStats *__restrict__ stats,
                                                      ShardRec *__restrict__ out)
{
    __shared__ double sE[1024];
    __shared__ int sI[1024];
    const int w = block_argmin(recs, nrecs, (int)sizeof(BestRec), sE, sI);
    const int tid = threadIdx.x;
    if (w >= 0) {
        const uint32_t *src = (const uint32_t *)(recs + w);
        uint32_t *dst = (uint32_t *)&out->best;
        for (int i = tid; i < (int)(sizeof(BestRec) / 4); i += blockDim.x) dst[i] = src[i];
    } else if (tid == 0) { out->best.valid = 0; out->best.m = 0; out->best.E = 0.0; }
    const unsigned long long *ss = (const unsigned long long *)stats;
    for (int i = tid; i < kStatsSummed; i += blockDim.x) out->sums[i] = ss[i];
}

// ------------------------------------------------------------------------------------------------
// k_finalize: the winning group → the reference's printed result.
//   canonical generators = RREF of the generator rows, rows sorted by pivot, with tracked signs
//   (Stab::canonicalize, cpp/stab.cpp:437-449 via Gauss_elimination_full cpp/linear.cpp:129);
//   sign pattern = FIRST minimum of the 2^m table indexed with generator 0 as the most significant bit
//   and '+' = 0 (cpp/state.cpp:681-682, cpp/stab.cpp:37-45); per-term signs = Stab::energy
//   (cpp/stab.cpp:579-587, printed by cpp/sparse.cpp:13-16).
struct FinalOut {
    double energy;
    int m;
    int error;
    uint64_t rows[kMaxG * 4];
    int8_t signs[kMaxG];
};

// Phase 0 combines the shards: argmin over the nshard records (minimum energy, ties by tuple order — the same winner
// on every rank), counters summed into D.stats; the winner is left in *gbest for the host.
template <int W>
__global__ void __launch_bounds__(256)
k_finalize(SparseDev D, const ShardRec *__restrict__ shards, int nshard, BestRec *gbest, const int *__restrict__ order, FinalOut *out,
           int8_t *__restrict__ term_signs, uint64_t *__restrict__ memX, double *__restrict__ memA,
           int *__restrict__ memT, int *memCount)
{
    __shared__ uint64_t C[kMaxG * W], G[kMaxG * W];
    __shared__ int qC[kMaxG], pivC[kMaxG];
    __shared__ double redE[256];
    __shared__ unsigned long long redI[256];
    const int tid = threadIdx.x, T = D.T;
    {
        __shared__ int sI[256];
        const int w = block_argmin(&shards[0].best, nshard, (int)sizeof(ShardRec), redE, sI);
        if (w >= 0) {
            const uint32_t *src = (const uint32_t *)&shards[w].best;
            uint32_t *dst = (uint32_t *)gbest;
            for (int i = tid; i < (int)(sizeof(BestRec) / 4); i += blockDim.x) dst[i] = src[i];
        } else if (tid == 0) { gbest->valid = 0; gbest->m = 0; gbest->E = 0.0; }
        unsigned long long *ss = (unsigned long long *)D.stats;
        for (int i = tid; i < kStatsSummed; i += blockDim.x) {
            unsigned long long a = 0;
            for (int r = 0; r < nshard; r++) a += shards[r].sums[i];
            ss[i] = a;
        }
        __syncthreads();
        if (w < 0) { if (tid == 0) { out->error = 0; out->m = 0; out->energy = 0.0; *memCount = 0; } return; }
    }
    const BestRec *best = gbest;
    const int m = best->m;

    if (tid < 32) {
        // K1: reduced row echelon form, pivot = lowest set bit, with phase tracking.  One warp: the earlier rows are
        // spread over the lanes.  The group is abelian, so a product of its elements (phase included) does not depend
        // on the order of the factors, and because every earlier pivot column is set in its own row only, WHICH earlier
        // rows reduce the new one can be read off the new row before any of them is applied.
        const int lane = tid;
        if (lane == 0) out->error = 0;
        for (int b = lane; b < m; b += 32) {                  // all generator rows at once: one round trip, not m
            const int t = best->g[b];
            for (int k = 0; k < W; k++) G[b * W + k] = D.rows[(size_t)t * W + k];
        }
        __syncwarp();
        for (int i = 0; i < m; i++) {
            uint64_t v[W], acc[W];
            for (int k = 0; k < W; k++) { v[k] = G[i * W + k]; acc[k] = 0; }
            int q = row_plus_phase<W>(v), qa = 0;
            for (int b = lane; b < i; b += 32)
                if (row_bit<W>(v, pivC[b])) row_mul<W>(acc, qa, C + b * W, qC[b]);
            for (int off = 16; off > 0; off >>= 1) {
                uint64_t o[W];
                for (int k = 0; k < W; k++) o[k] = __shfl_xor_sync(0xffffffffu, acc[k], off);
                const int oq = __shfl_xor_sync(0xffffffffu, qa, off);
                row_mul<W>(acc, qa, o, oq);
            }
            row_mul<W>(v, q, acc, qa);
            const int pv = row_lowbit<W>(v);
            for (int b = lane; b < i; b += 32)                // clear the new pivot from earlier rows
                if (row_bit<W>(C + b * W, pv)) row_mul<W>(C + b * W, qC[b], v, q);
            if (lane == 0) {
                for (int k = 0; k < W; k++) C[i * W + k] = v[k];
                qC[i] = q; pivC[i] = pv;
            }
            __syncwarp();
        }
        {                                                     // sort rows by pivot: every row counts the smaller pivots
            uint64_t v[2][W];
            int pos[2], pvs[2];
            for (int h = 0; h < 2; h++) {
                const int b = lane + 32 * h;
                pos[h] = 0; pvs[h] = 0;
                if (b < m) {
                    pvs[h] = pivC[b];
                    for (int k = 0; k < W; k++) v[h][k] = C[b * W + k];
                    for (int j = 0; j < m; j++) pos[h] += pivC[j] < pvs[h];
                }
            }
            __syncwarp();
            // canonicalize_signs (cpp/stab.cpp:685-691): tables are indexed against '+' generators
            for (int h = 0; h < 2; h++)
                if (lane + 32 * h < m) {
                    for (int k = 0; k < W; k++) C[pos[h] * W + k] = v[h][k];
                    pivC[pos[h]] = pvs[h];
                    qC[pos[h]] = row_plus_phase<W>(v[h]);
                }
        }
    }
    __syncthreads();

    // K2: decompose every term over the '+' canonical generators.  The members are written in ascending term order
    // (ballot + per-warp counts: no atomics, no sort), which is the summation order of the energies below.
    __shared__ int wcnt[8];
    int nm = 0;
    for (int t0 = 0; t0 < T; t0 += blockDim.x) {
        const int t = t0 + tid;
        bool member = false;
        uint64_t x = 0;
        int q = 0;
        if (t < T) {
            uint64_t v[W];
            for (int k = 0; k < W; k++) v[k] = D.rows[(size_t)t * W + k];
            q = row_plus_phase<W>(v);
            bool com = true;
            for (int i = 0; i < m; i++) com &= !row_anticommute<W>(v, C + i * W);
            if (com) {
                for (int i = 0; i < m; i++)
                    if (row_bit<W>(v, pivC[i])) { row_mul<W>(v, q, C + i * W, qC[i]); x |= 1ULL << (m - 1 - i); }
                member = row_is_zero<W>(v);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, member);
        if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { before += w < (tid >> 5) ? wcnt[w] : 0; all += wcnt[w]; }
        if (member) {   // P_t = tau * prod_{i in x} C_i^+, tau = (-1)^{q/2}; bit 30 of memT carries tau < 0
            const int slot = nm + before + __popc(bal & ((1u << (tid & 31)) - 1u));
            memT[slot] = t | ((q & 2) ? (1 << 30) : 0); memX[slot] = x; memA[slot] = (q & 2) ? -D.wgt[t] : D.wgt[t];
        }
        if (t < T) term_signs[order[t]] = 0;
        nm += all;
        __syncthreads();
    }
    if (tid == 0) *memCount = nm;
    __syncthreads();

    // K5: first minimum over the 2^m sign patterns
    double myE = 1e300;
    unsigned long long myI = ~0ULL;
    if (m <= kMaxBrute) {
        for (unsigned long long idx = tid; idx < (1ULL << m); idx += blockDim.x) {
            double e = 0.0;
            for (int i = 0; i < nm; i++) e += (popc64(idx & memX[i]) & 1) ? -memA[i] : memA[i];
            if (e < myE) { myE = e; myI = idx; }
        }
    } else {
        // more generators than the table limit: exact only if every member is a generator on its own
        // (then the signs decouple); otherwise report the limit.
        if (tid == 0) {
            bool ok = nm == m;
            unsigned long long idx = 0;
            double e = 0.0;
            // solve x-matrix (unit-triangular after sorting by pivot is not guaranteed): Gaussian solve over GF(2)
            uint64_t Xm[kMaxG]; int want[kMaxG];
            for (int i = 0; i < nm && ok; i++) { Xm[i] = memX[i]; want[i] = memA[i] > 0.0 ? 1 : 0; e -= fabs(memA[i]); }
            // find idx with parity(idx & Xm[i]) = want[i] for all i
            for (int c = 0, r = 0; c < m && ok; c++) {
                int pr = -1;
                for (int i = r; i < nm; i++) if ((Xm[i] >> c) & 1) { pr = i; break; }
                if (pr < 0) { ok = false; break; }
                uint64_t tx = Xm[pr]; Xm[pr] = Xm[r]; Xm[r] = tx;
                int tw = want[pr]; want[pr] = want[r]; want[r] = tw;
                for (int i = 0; i < nm; i++) if (i != r && ((Xm[i] >> c) & 1)) { Xm[i] ^= Xm[r]; want[i] ^= want[r]; }
                r++;
            }
            if (ok) { for (int i = 0; i < m; i++) if (want[i]) idx |= 1ULL << ctz64(Xm[i]); myE = e; myI = idx; }
            else out->error = kErrLimit;
        }
    }
    redE[tid] = myE; redI[tid] = myI;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) {
            if (redE[tid + s] < redE[tid] || (redE[tid + s] == redE[tid] && redI[tid + s] < redI[tid])) {
                redE[tid] = redE[tid + s]; redI[tid] = redI[tid + s];
            }
        }
        __syncthreads();
    }
    const unsigned long long idx = redI[0];
    if (tid == 0) {
        out->energy = redE[0];
        out->m = m;
        for (int i = 0; i < m; i++) {
            for (int k = 0; k < W; k++) out->rows[i * W + k] = C[i * W + k];
            out->signs[i] = ((idx >> (m - 1 - i)) & 1) ? -1 : 1;
        }
    }
    for (int i = tid; i < nm; i += blockDim.x) {
        const int tau = (memT[i] >> 30) & 1;
        const int flip = popc64(idx & memX[i]) & 1;
        term_signs[order[memT[i] & 0xffff]] = (int8_t)((tau ^ flip) ? -1 : 1);
    }
}

}  // namespace sgs
