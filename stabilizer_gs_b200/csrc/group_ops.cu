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

// K1 + K2 back to back on the caller's stream and workspace (no allocation, one synchronisation): canonicalise
// the m generator rows, then the sign of every one of the T terms under the canonical group.
size_t group_finish_ws_bytes(int T) { return (size_t)kMaxRows * 4 * 8 + (size_t)std::max(T, 1) * 4 * 8 + kMaxRows + std::max(T, 1) + 64 + 1024; }

Status group_finish(cudaStream_t st, void *ws, int n, int m, uint64_t *rows, int8_t *signs, int *rank, int T, const uint64_t *P,
                    int8_t *term_signs)
{
    if (m > kMaxRows) return Status::fail(SGS_ERR_LIMIT, "at most 128 generator rows");
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
