// nccl_glue.hpp — the one collective step of the sharded searches: NCCL all-gather of the per-shard
// (energy, generating tuple) minima and all-reduce(sum) of the candidate counters, over NVLink/NVSwitch.
// Works single-process (ncclCommInitAll over the local devices) and multi-process (one communicator
// rank per local device, bootstrapped from a shared ncclUniqueId).
#pragma once
#include <cuda_runtime.h>

#include <memory>
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
    Status broadcast_arrays(const std::vector<std::vector<void *>> &bufs, const std::vector<size_t> &bytes, const std::vector<cudaStream_t> &streams);
    Status all_reduce_sum_i32(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams);
    Status all_reduce_max_f64(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams);
    // one NCCL group: in-place all-gathers of several arrays (array a of local device i: block `bytes[a]` per rank at
    // base[i][a] + global_rank * bytes[a]) and an all-reduce(max) of one int per device — the per-site exchange of
    // the multi-GPU k-local DP
    Status gather_arrays_and_max(const std::vector<std::vector<void *>> &base, const std::vector<size_t> &bytes,
                                 const std::vector<void *> &max_word, const std::vector<cudaStream_t> &streams);
    int global_rank(int local) const;      // rank of local device `local` in the communicator
    int size() const;                      // ranks in the communicator
    static Status unique_id(void *out128);
    // process-wide cache of communicators keyed by (unique id, rank, world, devices): an id can bootstrap only one
    static Status get(const std::vector<int> &devices, int rank, int world, const void *unique_id, std::shared_ptr<NcclWorld> &out);

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace sgs
