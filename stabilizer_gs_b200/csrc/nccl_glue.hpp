// nccl_glue.hpp — the one collective step of the sharded searches: NCCL all-gather of the per-shard
// (energy, generating tuple) minima and all-reduce(sum) of the candidate counters, over NVLink/NVSwitch.
// Works single-process (ncclCommInitAll over the local devices) and multi-process (one communicator
// rank per local device, bootstrapped from a shared ncclUniqueId).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.hpp"

namespace sgs {

class NcclWorld {
  public:
    NcclWorld();
    ~NcclWorld();
    Status init(const std::vector<int> &devices, int rank, int world, const void *unique_id);
    Status all_gather_bytes(const std::vector<void *> &send, const std::vector<void *> &recv, size_t bytes,
                            const std::vector<cudaStream_t> &streams);
    Status all_reduce_sum_u64(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams);
    Status all_reduce_max_f64(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams);
    static Status unique_id(void *out128);

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace sgs
