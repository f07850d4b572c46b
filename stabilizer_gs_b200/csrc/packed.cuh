// packed.cuh — packed symplectic Pauli rows on the device (and host).
//
// Layout: W little-endian u64 words per row, bit 2j = z_j, bit 2j+1 = x_j, i.e. the
// reference's Pauli::array() column order [z0,x0,z1,x1,...] (cpp/stab.cpp:194-199), so the lowest set
// bit is the reference's Gaussian-elimination pivot (cpp/linear.cpp:50).  A signed Pauli is
// i^q * prod_j Z_j^{z_j} X_j^{x_j} with q = ZX_phase in Z_4 (cpp/stab.cpp:176-192).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SGS_HD __host__ __device__ __forceinline__
#define SGS_D __device__ __forceinline__
#else
#define SGS_HD inline
#define SGS_D inline
#endif

namespace sgs {

constexpr uint64_t kEven = 0x5555555555555555ULL;

SGS_HD int popc64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

// parity of the set bits: the two halves are folded first, so one POPC (a quarter-rate instruction) instead of two
SGS_HD int parity64(uint64_t v)
{
    const uint32_t x = (uint32_t)v ^ (uint32_t)(v >> 32);
#if defined(__CUDA_ARCH__)
    return __popc(x) & 1;
#else
    return __builtin_popcount(x) & 1;
#endif
}

SGS_HD int ctz64(uint64_t v)   // v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// exchange the z and x bit of every site: a & swapzx(b) has a set bit per anticommuting (z,x) pair
SGS_HD uint64_t swapzx(uint64_t b) { return ((b >> 1) & kEven) | ((b & kEven) << 1); }

// number of Y letters (z = x = 1), cpp/stab.cpp:172-174
template <int W>
SGS_HD int row_ny(const uint64_t *r)
{
    int c = 0;
#pragma unroll
    for (int i = 0; i < W; i++) c += popc64(r[i] & (r[i] >> 1) & kEven);
    return c;
}

// ZX_phase of a '+' signed string: group_phase 0 → q = -nY mod 4 (cpp/stab.cpp:180-182)
template <int W>
SGS_HD int row_plus_phase(const uint64_t *r) { return (4 - (row_ny<W>(r) & 3)) & 3; }

// printed sign from q: '+' iff (q + nY) mod 4 == 0 (cpp/stab.cpp:176-192,356)
template <int W>
SGS_HD int row_sign(const uint64_t *r, int q) { return ((q + row_ny<W>(r)) & 3) == 0 ? 1 : -1; }

// Pauli::commute (cpp/stab.cpp:314-318): 1 if the two strings ANTIcommute. bsw = swapzx of b's words.
template <int W>
SGS_HD int row_anticommute_sw(const uint64_t *a, const uint64_t *bsw)
{
    uint64_t acc = 0;                              // only the parity matters: fold the words, one POPC
#pragma unroll
    for (int i = 0; i < W; i++) acc ^= a[i] & bsw[i];
    return parity64(acc);
}

template <int W>
SGS_HD int row_anticommute(const uint64_t *a, const uint64_t *b)
{
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < W; i++) acc ^= a[i] & swapzx(b[i]);
    return parity64(acc);
}

// Pauli::dot (cpp/stab.cpp:320-330): a <- a * b; q_a + q_b + 2 * (x_a . z_b)
template <int W>
SGS_HD void row_mul(uint64_t *a, int &qa, const uint64_t *b, int qb)
{
    uint64_t acc = 0;                              // 2 * (x_a . z_b) mod 4: only the parity of the count matters
#pragma unroll
    for (int i = 0; i < W; i++) acc ^= (a[i] >> 1) & b[i] & kEven;
    qa = (qa + qb + 2 * parity64(acc)) & 3;
#pragma unroll
    for (int i = 0; i < W; i++) a[i] ^= b[i];
}

template <int W>
SGS_HD bool row_is_zero(const uint64_t *r)
{
    uint64_t a = 0;
#pragma unroll
    for (int i = 0; i < W; i++) a |= r[i];
    return a == 0;
}

template <int W>
SGS_HD int row_lowbit(const uint64_t *r)   // index of the lowest set bit; row must be non-zero
{
    if (W == 1) return ctz64(r[0]);
#pragma unroll
    for (int i = 0; i < W; i++)
        if (r[i]) return i * 64 + ctz64(r[i]);
    return -1;
}

template <int W>
SGS_HD int row_bit(const uint64_t *r, int c)
{
    if (W == 1) return (int)((r[0] >> c) & 1);
    uint64_t w = r[0];                       // select without dynamic indexing (keeps rows in registers)
#pragma unroll
    for (int i = 1; i < W; i++) w = ((c >> 6) == i) ? r[i] : w;
    return (int)((w >> (c & 63)) & 1);
}

template <int W>
SGS_HD bool row_equal(const uint64_t *a, const uint64_t *b)
{
    uint64_t d = 0;
#pragma unroll
    for (int i = 0; i < W; i++) d |= a[i] ^ b[i];
    return d == 0;
}

}  // namespace sgs
