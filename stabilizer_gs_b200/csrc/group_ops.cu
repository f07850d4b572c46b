// group_ops.cu — device utilities behind sgs_group_* / sgs_pauli_commute / sgs_measure_int_peaks.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "group_ops.hpp"
#include "packed.cuh"

namespace sgs {

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

namespace {

constexpr int kMaxRows = 128;

// K1: Stab::canonicalize (cpp/stab.cpp:437-449): RREF with pivot = lowest set bit, rows sorted by pivot,
// signs carried by phase-tracked products (Pauli::dot, cpp/stab.cpp:320-330).  One warp, m <= 128; the rows of a
// stabilizer group commute, so products may be taken in any order.  Rows enter one by one into a fully reduced
// basis (no row has a bit at another row's pivot): which basis rows multiply into the newcomer is therefore
// decided by the newcomer's own bits, the lanes multiply their share of them and a butterfly combines the
// partial products; clearing the new pivot from the basis is one row per lane.
template <int W>
__global__ void k_canonicalize(int m, uint64_t *rows, int8_t *signs, int *rank)
{
    __shared__ uint64_t C[kMaxRows * W];
    __shared__ int q[kMaxRows], piv[kMaxRows], dst[kMaxRows];
    const int lane = threadIdx.x;
    int r = 0;
    for (int i = 0; i < m; i++) {
        uint64_t v[W];
        for (int k = 0; k < W; k++) v[k] = rows[(size_t)i * W + k];
        int qv = row_plus_phase<W>(v);
        if (signs[i] < 0) qv = (qv + 2) & 3;
        uint64_t acc[W];
        int qa = 0;
        for (int k = 0; k < W; k++) acc[k] = 0;
        for (int b = lane; b < r; b += 32)
            if (row_bit<W>(v, piv[b])) row_mul<W>(acc, qa, C + b * W, q[b]);
        for (int o = 16; o; o >>= 1) {
            uint64_t oth[W];
            for (int k = 0; k < W; k++) oth[k] = __shfl_xor_sync(0xffffffffu, acc[k], o);
            const int qo = __shfl_xor_sync(0xffffffffu, qa, o);
            row_mul<W>(acc, qa, oth, qo);
        }
        row_mul<W>(v, qv, acc, qa);
        if (row_is_zero<W>(v)) continue;                       // uniform: every lane holds the same v
        const int pv = row_lowbit<W>(v);
        for (int b = lane; b < r; b += 32)
            if (row_bit<W>(C + b * W, pv)) row_mul<W>(C + b * W, q[b], v, qv);
        if (lane == 0) {
            for (int k = 0; k < W; k++) C[r * W + k] = v[k];
            q[r] = qv; piv[r] = pv;
        }
        r++;
        __syncwarp();
    }
    // rows in pivot order: position = number of smaller pivots (pivots are distinct)
    for (int b = lane; b < r; b += 32) {
        int c = 0;
        for (int j = 0; j < r; j++) c += piv[j] < piv[b];
        dst[b] = c;
    }
    __syncwarp();
    for (int b = lane; b < r; b += 32) {
        for (int k = 0; k < W; k++) rows[(size_t)dst[b] * W + k] = C[b * W + k];
        signs[dst[b]] = (int8_t)row_sign<W>(C + b * W, q[b]);
    }
    if (lane == 0) *rank = r;
}

// K2: Stab::energy / decompose (cpp/stab.cpp:493-505,579-587) for a batch; the group is canonical.
template <int W>
__global__ void k_group_energy(int m, const uint64_t *__restrict__ rows, const int8_t *__restrict__ signs, int B,
                               const uint64_t *__restrict__ P, const int8_t *__restrict__ Psigns, int8_t *__restrict__ out,
                               const int *m_dev = nullptr)
{
    if (m_dev) m = *m_dev;               // rank left on the device by k_canonicalize
    __shared__ uint64_t C[kMaxRows * W];
    __shared__ int q[kMaxRows], piv[kMaxRows];
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        for (int k = 0; k < W; k++) C[i * W + k] = rows[(size_t)i * W + k];
        int qq = row_plus_phase<W>(C + i * W);
        if (signs[i] < 0) qq = (qq + 2) & 3;
        q[i] = qq;
        piv[i] = row_lowbit<W>(C + i * W);
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint64_t v[W];
    for (int k = 0; k < W; k++) v[k] = P[(size_t)b * W + k];
    int qv = row_plus_phase<W>(v);
    if (Psigns && Psigns[b] < 0) qv = (qv + 2) & 3;
    for (int i = 0; i < m; i++)
        if (row_bit<W>(v, piv[i])) row_mul<W>(v, qv, C + i * W, q[i]);
    out[b] = row_is_zero<W>(v) ? (int8_t)((qv & 2) ? -1 : 1) : (int8_t)0;
}

// K3: Pauli::commute (cpp/stab.cpp:314-318) for a batch of pairs
template <int W>
__global__ void k_commute(int B, const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint8_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    out[i] = row_anticommute<W>(a + (size_t)i * W, b + (size_t)i * W) ? 0 : 1;
}

// issue-rate microbenchmarks: 8 independent dependency chains per thread
__global__ void k_peak_lop3(unsigned iters, unsigned seed, unsigned *sink, long long *cycles)
{
    unsigned x0 = seed + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, x4 = x0 * 11, x5 = x0 * 13, x6 = x0 * 17, x7 = x0 * 19;
    const unsigned a = seed * 2654435761u, b = ~seed;
    const long long t0 = clock64();
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x0) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(x1) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x2) : "r"(b), "r"(a));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xca;" : "+r"(x3) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x4) : "r"(a), "r"(x0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(x5) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x6) : "r"(b), "r"(x1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x7) : "r"(a), "r"(b));
        }
    }
    const long long t1 = clock64();
    const unsigned r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (r == 0x12345678u) sink[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void k_peak_popc(unsigned iters, unsigned seed, unsigned *sink, long long *cycles)
{
    unsigned x0 = seed + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, x4 = x0 * 11, x5 = x0 * 13, x6 = x0 * 17, x7 = x0 * 19;
    const long long t0 = clock64();
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile("popc.b32 %0, %0;" : "+r"(x0));
            asm volatile("popc.b32 %0, %0;" : "+r"(x1));
            asm volatile("popc.b32 %0, %0;" : "+r"(x2));
            asm volatile("popc.b32 %0, %0;" : "+r"(x3));
            asm volatile("popc.b32 %0, %0;" : "+r"(x4));
            asm volatile("popc.b32 %0, %0;" : "+r"(x5));
            asm volatile("popc.b32 %0, %0;" : "+r"(x6));
            asm volatile("popc.b32 %0, %0;" : "+r"(x7));
        }
    }
    const long long t1 = clock64();
    const unsigned r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (r == 0x12345678u) sink[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int words_for(int n) { int W = (2 * n + 63) / 64; return W == 3 ? 4 : (W < 1 ? 1 : W); }

template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)); }
};

// repack rows between the caller's stride (ceil(2n/64)) and the kernel's W (3 words are run as 4)
std::vector<uint64_t> widen(const uint64_t *src, int count, int Win, int W)
{
    std::vector<uint64_t> v((size_t)(count ? count : 1) * W, 0);
    for (int i = 0; i < count; i++)
        for (int k = 0; k < Win; k++) v[(size_t)i * W + k] = src[(size_t)i * Win + k];
    return v;
}

}  // namespace

Status group_canonicalize(int device, int n, int m, uint64_t *rows, int8_t *signs, int *rank)
{
    if (m > kMaxRows) return Status::fail(SGS_ERR_LIMIT, "at most 128 generator rows");
    SGS_CUDA_TRY(cudaSetDevice(device));
    const int Win = (2 * n + 63) / 64, W = words_for(n);
    std::vector<uint64_t> h = widen(rows, m, Win, W);
    DevBuf<uint64_t> dr; DevBuf<int8_t> ds; DevBuf<int> dk;
    SGS_CUDA_TRY(dr.alloc(h.size())); SGS_CUDA_TRY(ds.alloc(m)); SGS_CUDA_TRY(dk.alloc(1));
    SGS_CUDA_TRY(cudaMemcpy(dr.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    if (m) SGS_CUDA_TRY(cudaMemcpy(ds.p, signs, m, cudaMemcpyHostToDevice));
    if (W == 1) k_canonicalize<1><<<1, 32>>>(m, dr.p, ds.p, dk.p);
    else if (W == 2) k_canonicalize<2><<<1, 32>>>(m, dr.p, ds.p, dk.p);
    else k_canonicalize<4><<<1, 32>>>(m, dr.p, ds.p, dk.p);
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaMemcpy(h.data(), dr.p, h.size() * 8, cudaMemcpyDeviceToHost));
    if (m) SGS_CUDA_TRY(cudaMemcpy(signs, ds.p, m, cudaMemcpyDeviceToHost));
    SGS_CUDA_TRY(cudaMemcpy(rank, dk.p, sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++)
        for (int k = 0; k < Win; k++) rows[(size_t)i * Win + k] = i < *rank ? h[(size_t)i * W + k] : 0;
    return Status();
}

// ------------------------------------------------------------------------------------------------
// Wide groups (n > 127 or more than 128 generators: the ground state of a long k-local chain).  Same results as
// k_canonicalize / k_group_energy — the RREF of a row space is unique and so is the sign of each of its elements —
// with rows of any width kept in global memory.
//   k_rref_wide     one block: Gauss-Jordan on the bit matrix [rows | I] by columns (lowest bit first), so that the
//                   rows come out sorted by pivot and the identity half records which generators multiply into each
//                   canonical row (Gauss_elimination_full, cpp/linear.cpp:129)
//   k_products_wide one warp per canonical row: the phase-tracked product of the selected signed generators
//                   (Stab::transform, cpp/stab.cpp:442-449; Pauli::dot :320-330)
//   k_signs_wide    one warp per term: reduction over the canonical rows, pivot by pivot (Stab::energy :579-587)
__global__ void __launch_bounds__(1024) k_rref_wide(int m, int nbits, int W, int Wu, uint64_t *Mx, int *perm, int *piv, int *rank_out,
                                                    int *row_of_bit)
{
    __shared__ int s_min[32];
    __shared__ int s_piv;
    extern __shared__ unsigned char s_flag[];          // [m] row has the pivot bit
    const int tid = threadIdx.x, nth = blockDim.x, Wt = W + Wu;
    for (int i = tid; i < m; i += nth) perm[i] = i;
    for (int b = tid; b < nbits; b += nth) row_of_bit[b] = -1;
    __syncthreads();
    int r = 0;
    for (int c = 0; c < nbits && r < m; c++) {
        const int cw = c >> 6;
        const uint64_t cb = 1ULL << (c & 63);
        int best = m;
        for (int i = r + tid; i < m; i += nth) if (Mx[(size_t)perm[i] * Wt + cw] & cb) { best = i; break; }
        best = __reduce_min_sync(0xffffffffu, best);
        if ((tid & 31) == 0) s_min[tid >> 5] = best;
        __syncthreads();
        if (tid < 32) {
            int v = tid < (nth >> 5) ? s_min[tid] : m;
            v = __reduce_min_sync(0xffffffffu, v);
            if (tid == 0) s_piv = v;
        }
        __syncthreads();
        const int pv = s_piv;
        if (pv >= m) continue;                           // (uniform: every thread read s_piv after the barrier)
        if (tid == 0) { const int t = perm[r]; perm[r] = perm[pv]; perm[pv] = t; piv[r] = c; row_of_bit[c] = r; }
        __syncthreads();
        const int pr = perm[r];
        for (int i = tid; i < m; i += nth) s_flag[i] = (i != r && (Mx[(size_t)perm[i] * Wt + cw] & cb)) ? 1 : 0;
        __syncthreads();
        for (int idx = tid; idx < m * Wt; idx += nth) {
            const int i = idx / Wt, w = idx - i * Wt;
            if (s_flag[i]) Mx[(size_t)perm[i] * Wt + w] ^= Mx[(size_t)pr * Wt + w];
        }
        __syncthreads();
        r++;
    }
    if (tid == 0) *rank_out = r;
}

// G: the m signed generators (W words each), qG their phases; canonical row i = product over the set bits of its
// combination mask.  Lane l holds words l, l + 32, ... of the accumulator (W <= 32 * kWideWpl).
constexpr int kWideWpl = 8;
__global__ void k_products_wide(int m, int W, int Wu, const uint64_t *__restrict__ Mx, const int *__restrict__ perm, const int *rank,
                                const uint64_t *__restrict__ G, const int *__restrict__ qG, uint64_t *__restrict__ C, int *__restrict__ qC,
                                int8_t *__restrict__ signs)
{
    const int lane = threadIdx.x & 31, i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, Wt = W + Wu;
    if (i >= *rank) return;
    const uint64_t *U = Mx + (size_t)perm[i] * Wt + W;
    uint64_t acc[kWideWpl];
#pragma unroll
    for (int k = 0; k < kWideWpl; k++) acc[k] = 0;
    int q = 0;
    for (int uw = 0; uw < Wu; uw++) {
        uint64_t bits = U[uw];
        while (bits) {
            const int j = uw * 64 + ctz64(bits);
            bits &= bits - 1;
            if (j >= m) break;
            uint64_t par = 0;
#pragma unroll
            for (int k = 0; k < kWideWpl; k++) {
                const int w = lane + 32 * k;
                if (w < W) { const uint64_t g = G[(size_t)j * W + w]; par ^= (acc[k] >> 1) & g & kEven; acc[k] ^= g; }
            }
            const unsigned p = __reduce_xor_sync(0xffffffffu, (unsigned)parity64(par));
            q = (q + qG[j] + 2 * (int)(p & 1)) & 3;
        }
    }
    int ny = 0;
#pragma unroll
    for (int k = 0; k < kWideWpl; k++) {
        const int w = lane + 32 * k;
        if (w < W) { C[(size_t)i * W + w] = acc[k]; ny += popc64(acc[k] & (acc[k] >> 1) & kEven); }
    }
    ny = __reduce_add_sync(0xffffffffu, ny);
    if (lane == 0) {
        const int sg = ((q + ny) & 3) == 0 ? 1 : -1;
        signs[i] = (int8_t)sg;
        qC[i] = (4 - (ny & 3)) & 3;                      // phase of the row taken with sign '+' ...
        if (sg < 0) qC[i] = (qC[i] + 2) & 3;             // ... and with its own sign
    }
}

__global__ void k_signs_wide(int T, int W, int nbits, const int *rank, const uint64_t *__restrict__ C, const int *__restrict__ qC,
                             const int *__restrict__ row_of_bit, const uint64_t *__restrict__ P, int8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    (void)rank;
    uint64_t v[kWideWpl];
    int ny = 0;
#pragma unroll
    for (int k = 0; k < kWideWpl; k++) {
        const int w = lane + 32 * k;
        v[k] = w < W ? P[(size_t)t * W + w] : 0;
        ny += popc64(v[k] & (v[k] >> 1) & kEven);
    }
    ny = __reduce_add_sync(0xffffffffu, ny);
    int q = (4 - (ny & 3)) & 3;
    bool member = true;
    for (;;) {
        // lowest set bit of v (every element of the row space has its lowest bit at a pivot; RREF rows carry no other pivot)
        int low = nbits;
#pragma unroll
        for (int k = kWideWpl - 1; k >= 0; k--) if (v[k]) low = (lane + 32 * k) * 64 + ctz64(v[k]);
        low = __reduce_min_sync(0xffffffffu, low);
        if (low >= nbits) break;                          // v == identity: a member
        const int i = row_of_bit[low];
        if (i < 0) { member = false; break; }
        uint64_t par = 0;
#pragma unroll
        for (int k = 0; k < kWideWpl; k++) {
            const int w = lane + 32 * k;
            if (w < W) { const uint64_t g = C[(size_t)i * W + w]; par ^= (v[k] >> 1) & g & kEven; v[k] ^= g; }
        }
        const unsigned p = __reduce_xor_sync(0xffffffffu, (unsigned)parity64(par));
        q = (q + qC[i] + 2 * (int)(p & 1)) & 3;
    }
    if (lane == 0) out[t] = member ? (int8_t)((q & 2) ? -1 : 1) : (int8_t)0;
}

Status group_finish_wide(cudaStream_t st, int n, int m, uint64_t *rows, int8_t *signs, int *rank, int T, const uint64_t *P, int8_t *term_signs)
{
    const int W = (2 * n + 63) / 64, Wu = (std::max(m, 1) + 63) / 64, Wt = W + Wu, nbits = 2 * n;
    if (W > 32 * kWideWpl) return Status::fail(SGS_ERR_LIMIT, "ground-state group: more than 8192 qubits");
    const int mm = std::max(m, 1), Tn = std::max(T, 1);
    std::vector<uint64_t> hM((size_t)mm * Wt, 0);
    std::vector<int> hq(mm, 0);
    for (int i = 0; i < m; i++) {
        int ny = 0;
        for (int w = 0; w < W; w++) { const uint64_t x = rows[(size_t)i * W + w]; hM[(size_t)i * Wt + w] = x; ny += __builtin_popcountll(x & (x >> 1) & kEven); }
        hM[(size_t)i * Wt + W + (i >> 6)] |= 1ULL << (i & 63);
        hq[i] = ((4 - (ny & 3)) & 3);
        if (signs[i] < 0) hq[i] = (hq[i] + 2) & 3;
    }
    DevBuf<uint64_t> dM, dG, dC, dP; DevBuf<int> dperm, dpiv, drank, drob, dqG, dqC; DevBuf<int8_t> dsg, dout;
    SGS_CUDA_TRY(dM.alloc(hM.size())); SGS_CUDA_TRY(dG.alloc((size_t)mm * W)); SGS_CUDA_TRY(dC.alloc((size_t)mm * W)); SGS_CUDA_TRY(dP.alloc((size_t)Tn * W));
    SGS_CUDA_TRY(dperm.alloc(mm)); SGS_CUDA_TRY(dpiv.alloc(mm)); SGS_CUDA_TRY(drank.alloc(1)); SGS_CUDA_TRY(drob.alloc(std::max(nbits, 1)));
    SGS_CUDA_TRY(dqG.alloc(mm)); SGS_CUDA_TRY(dqC.alloc(mm)); SGS_CUDA_TRY(dsg.alloc(mm)); SGS_CUDA_TRY(dout.alloc(Tn));
    SGS_CUDA_TRY(cudaMemcpyAsync(dM.p, hM.data(), hM.size() * 8, cudaMemcpyHostToDevice, st));
    if (m) SGS_CUDA_TRY(cudaMemcpyAsync(dG.p, rows, (size_t)m * W * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(dqG.p, hq.data(), (size_t)mm * 4, cudaMemcpyHostToDevice, st));
    if (T) SGS_CUDA_TRY(cudaMemcpyAsync(dP.p, P, (size_t)T * W * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaFuncSetAttribute(k_rref_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(mm, 1024)));
    k_rref_wide<<<1, 1024, mm, st>>>(m, nbits, W, Wu, dM.p, dperm.p, dpiv.p, drank.p, drob.p);
    k_products_wide<<<(mm * 32 + 127) / 128, 128, 0, st>>>(m, W, Wu, dM.p, dperm.p, drank.p, dG.p, dqG.p, dC.p, dqC.p, dsg.p);
    k_signs_wide<<<(Tn * 32 + 127) / 128, 128, 0, st>>>(T, W, nbits, drank.p, dC.p, dqC.p, drob.p, dP.p, dout.p);
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaMemcpyAsync(rank, drank.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<uint64_t> hc((size_t)mm * W, 0);
    SGS_CUDA_TRY(cudaMemcpy(hc.data(), dC.p, hc.size() * 8, cudaMemcpyDeviceToHost));
    if (m) SGS_CUDA_TRY(cudaMemcpy(signs, dsg.p, m, cudaMemcpyDeviceToHost));
    if (T) SGS_CUDA_TRY(cudaMemcpy(term_signs, dout.p, T, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++)
        for (int w = 0; w < W; w++) rows[(size_t)i * W + w] = i < *rank ? hc[(size_t)i * W + w] : 0;
    return Status();
}

// K1 + K2 back to back on the caller's stream and workspace (no allocation, one synchronisation): canonicalise
// the m generator rows, then the sign of every one of the T terms under the canonical group.
size_t group_finish_ws_bytes(int T) { return (size_t)kMaxRows * 4 * 8 + (size_t)std::max(T, 1) * 4 * 8 + kMaxRows + std::max(T, 1) + 64 + 1024; }

Status group_finish(cudaStream_t st, void *ws, int n, int m, uint64_t *rows, int8_t *signs, int *rank, int T, const uint64_t *P,
                    int8_t *term_signs)
{
    if (m > kMaxRows || n + 1 > 128) return group_finish_wide(st, n, m, rows, signs, rank, T, P, term_signs);
    const int Win = (2 * n + 63) / 64, W = words_for(n), Tn = std::max(T, 1);
    char *base = (char *)ws;
    uint64_t *d_rows = (uint64_t *)base; base += (size_t)kMaxRows * 4 * 8;
    uint64_t *d_P = (uint64_t *)base; base += (size_t)Tn * 4 * 8;
    int *d_rank = (int *)base; base += 64;
    int8_t *d_signs = (int8_t *)base; base += (kMaxRows + 255) & ~255;
    int8_t *d_out = (int8_t *)base;
    std::vector<uint64_t> hr = widen(rows, m, Win, W), hp = widen(P, T, Win, W);
    SGS_CUDA_TRY(cudaMemcpyAsync(d_rows, hr.data(), hr.size() * 8, cudaMemcpyHostToDevice, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(d_P, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice, st));
    if (m) SGS_CUDA_TRY(cudaMemcpyAsync(d_signs, signs, m, cudaMemcpyHostToDevice, st));
    const int grid = (Tn + 127) / 128;
    if (W == 1) { k_canonicalize<1><<<1, 32, 0, st>>>(m, d_rows, d_signs, d_rank); k_group_energy<1><<<grid, 128, 0, st>>>(0, d_rows, d_signs, T, d_P, nullptr, d_out, d_rank); }
    else if (W == 2) { k_canonicalize<2><<<1, 32, 0, st>>>(m, d_rows, d_signs, d_rank); k_group_energy<2><<<grid, 128, 0, st>>>(0, d_rows, d_signs, T, d_P, nullptr, d_out, d_rank); }
    else { k_canonicalize<4><<<1, 32, 0, st>>>(m, d_rows, d_signs, d_rank); k_group_energy<4><<<grid, 128, 0, st>>>(0, d_rows, d_signs, T, d_P, nullptr, d_out, d_rank); }
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaMemcpyAsync(hr.data(), d_rows, hr.size() * 8, cudaMemcpyDeviceToHost, st));
    if (m) SGS_CUDA_TRY(cudaMemcpyAsync(signs, d_signs, m, cudaMemcpyDeviceToHost, st));
    if (T) SGS_CUDA_TRY(cudaMemcpyAsync(term_signs, d_out, T, cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaMemcpyAsync(rank, d_rank, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGS_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < m; i++)
        for (int k = 0; k < Win; k++) rows[(size_t)i * Win + k] = i < *rank ? hr[(size_t)i * W + k] : 0;
    return Status();
}

Status group_energy(int device, int n, int m, const uint64_t *rows, const int8_t *signs, int B, const uint64_t *P,
                    const int8_t *Psigns, int8_t *out)
{
    if (m > kMaxRows) return Status::fail(SGS_ERR_LIMIT, "at most 128 generator rows");
    if (B == 0) return Status();
    SGS_CUDA_TRY(cudaSetDevice(device));
    const int Win = (2 * n + 63) / 64, W = words_for(n);
    std::vector<uint64_t> hr = widen(rows, m, Win, W), hp = widen(P, B, Win, W);
    DevBuf<uint64_t> dr, dp; DevBuf<int8_t> ds, dps, dout;
    SGS_CUDA_TRY(dr.alloc(hr.size())); SGS_CUDA_TRY(dp.alloc(hp.size()));
    SGS_CUDA_TRY(ds.alloc(m)); SGS_CUDA_TRY(dps.alloc(B)); SGS_CUDA_TRY(dout.alloc(B));
    SGS_CUDA_TRY(cudaMemcpy(dr.p, hr.data(), hr.size() * 8, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(dp.p, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
    if (m) SGS_CUDA_TRY(cudaMemcpy(ds.p, signs, m, cudaMemcpyHostToDevice));
    if (Psigns) SGS_CUDA_TRY(cudaMemcpy(dps.p, Psigns, B, cudaMemcpyHostToDevice));
    const int8_t *psg = Psigns ? dps.p : nullptr;
    const int grid = (B + 127) / 128;
    if (W == 1) k_group_energy<1><<<grid, 128>>>(m, dr.p, ds.p, B, dp.p, psg, dout.p);
    else if (W == 2) k_group_energy<2><<<grid, 128>>>(m, dr.p, ds.p, B, dp.p, psg, dout.p);
    else k_group_energy<4><<<grid, 128>>>(m, dr.p, ds.p, B, dp.p, psg, dout.p);
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaMemcpy(out, dout.p, B, cudaMemcpyDeviceToHost));
    return Status();
}

Status pauli_commute_batch(int device, int n, int B, const uint64_t *a, const uint64_t *b, uint8_t *out)
{
    if (B == 0) return Status();
    SGS_CUDA_TRY(cudaSetDevice(device));
    const int Win = (2 * n + 63) / 64, W = words_for(n);
    std::vector<uint64_t> ha = widen(a, B, Win, W), hb = widen(b, B, Win, W);
    DevBuf<uint64_t> da, db; DevBuf<uint8_t> dout;
    SGS_CUDA_TRY(da.alloc(ha.size())); SGS_CUDA_TRY(db.alloc(hb.size())); SGS_CUDA_TRY(dout.alloc(B));
    SGS_CUDA_TRY(cudaMemcpy(da.p, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
    SGS_CUDA_TRY(cudaMemcpy(db.p, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
    const int grid = (B + 255) / 256;
    if (W == 1) k_commute<1><<<grid, 256>>>(B, da.p, db.p, dout.p);
    else if (W == 2) k_commute<2><<<grid, 256>>>(B, da.p, db.p, dout.p);
    else k_commute<4><<<grid, 256>>>(B, da.p, db.p, dout.p);
    SGS_CUDA_TRY(cudaGetLastError());
    SGS_CUDA_TRY(cudaMemcpy(out, dout.p, B, cudaMemcpyDeviceToHost));
    return Status();
}

Status device_flush_l2(int device)
{
    static void *buf[16] = {nullptr};
    const size_t bytes = 512ull << 20;       // > 126 MB L2
    if (device < 0 || device >= 16) return Status::fail(SGS_ERR_ARG, "device out of range");
    SGS_CUDA_TRY(cudaSetDevice(device));
    if (!buf[device]) SGS_CUDA_TRY(cudaMalloc(&buf[device], bytes));
    static int tick = 0;
    SGS_CUDA_TRY(cudaMemset(buf[device], ++tick & 0xff, bytes));
    SGS_CUDA_TRY(cudaDeviceSynchronize());
    return Status();
}

Status measure_int_peaks(int device, double *lop3, double *popc, double *sm_mhz)
{
    SGS_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop{};
    SGS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    DevBuf<unsigned> sink; DevBuf<long long> cyc;
    SGS_CUDA_TRY(sink.alloc(1)); SGS_CUDA_TRY(cyc.alloc(1));
    cudaEvent_t e0, e1;
    SGS_CUDA_TRY(cudaEventCreate(&e0)); SGS_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8, block = 256;
    const unsigned iters = 1 << 14;
    const double ops = (double)grid * block * iters * 32.0;
    double best[2] = {0, 0}, mhz = 0;
    for (int which = 0; which < 2; which++) {
        for (int rep = 0; rep < 4; rep++) {
            SGS_CUDA_TRY(cudaEventRecord(e0));
            if (which == 0) k_peak_lop3<<<grid, block>>>(iters, 12345u + rep, sink.p, cyc.p);
            else k_peak_popc<<<grid, block>>>(iters, 12345u + rep, sink.p, cyc.p);
            SGS_CUDA_TRY(cudaEventRecord(e1));
            SGS_CUDA_TRY(cudaEventSynchronize(e1));
            float ms = 0;
            SGS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 0) continue;
            best[which] = std::max(best[which], ops / (ms * 1e-3));
            long long c = 0;
            SGS_CUDA_TRY(cudaMemcpy(&c, cyc.p, sizeof(c), cudaMemcpyDeviceToHost));
            // block 0 ran for c cycles out of a kernel that took ms: a lower bound of the running clock
            mhz = std::max(mhz, (double)c / (ms * 1e-3) / 1e6);
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (lop3) *lop3 = best[0];
    if (popc) *popc = best[1];
    // running SM clock implied by the LOP3 rate: 64 INT32 lanes per SM per clock (alu pipe, rt_SMSP = 2)
    (void)mhz;
    if (sm_mhz) *sm_mhz = best[0] / ((double)prop.multiProcessorCount * 64.0) / 1e6;
    return Status();
}

}  // namespace sgs
