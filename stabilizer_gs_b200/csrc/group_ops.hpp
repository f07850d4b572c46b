// group_ops.hpp — small device utilities behind the C ABI: batched commutation, group
// canonicalisation / membership, and the integer-pipe issue-rate microbenchmarks.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.hpp"

namespace sgs {

int device_count();
Status group_canonicalize(int device, int n, int m, uint64_t *rows, int8_t *signs, int *rank);
Status group_energy(int device, int n, int m, const uint64_t *rows, const int8_t *signs, int B, const uint64_t *P,
                    const int8_t *Psigns, int8_t *out);
// canonicalise + per-term signs on the caller's stream with a caller-provided device workspace of
// group_finish_ws_bytes(T) bytes: no allocation, one synchronisation
size_t group_finish_ws_bytes(int T);
Status group_finish(cudaStream_t st, void *ws, int n, int m, uint64_t *rows, int8_t *signs, int *rank, int T, const uint64_t *P,
                    int8_t *term_signs);
Status pauli_commute_batch(int device, int n, int B, const uint64_t *a, const uint64_t *b, uint8_t *out);
Status device_flush_l2(int device);
Status measure_int_peaks(int device, double *lop3_ops_per_s, double *popc_ops_per_s, double *sm_mhz);

}  // namespace sgs
