// nccl_glue.cpp — see nccl_glue.hpp.
#include "nccl_glue.hpp"

#include <nccl.h>

#include <cstring>

namespace sgs {

struct NcclWorld::Impl {
    std::vector<ncclComm_t> comms;
    std::vector<int> devices;
};

#define SGS_NCCL_TRY(expr)                                                                              \
    do {                                                                                                \
        ncclResult_t r__ = (expr);                                                                      \
        if (r__ != ncclSuccess)                                                                         \
            return Status::fail(SGS_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(r__));     \
    } while (0)

NcclWorld::NcclWorld() : impl_(new Impl()) {}

NcclWorld::~NcclWorld()
{
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        cudaSetDevice(impl_->devices[i]);
        ncclCommDestroy(impl_->comms[i]);
    }
    delete impl_;
}

Status NcclWorld::unique_id(void *out128)
{
    static_assert(sizeof(ncclUniqueId) <= SGS_NCCL_ID_BYTES, "ncclUniqueId larger than SGS_NCCL_ID_BYTES");
    ncclUniqueId id;
    SGS_NCCL_TRY(ncclGetUniqueId(&id));
    memset(out128, 0, SGS_NCCL_ID_BYTES);
    memcpy(out128, &id, sizeof(id));
    return Status();
}

Status NcclWorld::init(const std::vector<int> &devices, int rank, int world, const void *unique_id)
{
    impl_->devices = devices;
    const int ng = (int)devices.size();
    impl_->comms.assign(ng, nullptr);
    if (world <= 1 && unique_id == nullptr) {
        SGS_NCCL_TRY(ncclCommInitAll(impl_->comms.data(), ng, devices.data()));
        return Status();
    }
    if (unique_id == nullptr) return Status::fail(SGS_ERR_ARG, "multi-process NCCL needs a shared unique id");
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    SGS_NCCL_TRY(ncclGroupStart());
    for (int i = 0; i < ng; i++) {
        cudaSetDevice(devices[i]);
        ncclResult_t r = ncclCommInitRank(&impl_->comms[i], world * ng, id, rank * ng + i);
        if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
    }
    SGS_NCCL_TRY(ncclGroupEnd());
    return Status();
}

Status NcclWorld::all_gather_bytes(const std::vector<void *> &send, const std::vector<void *> &recv, size_t bytes,
                                   const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        ncclResult_t r = ncclAllGather(send[i], recv[i], bytes, ncclChar, impl_->comms[i], streams[i]);
        if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclAllGather: ") + ncclGetErrorString(r)); }
    }
    SGS_NCCL_TRY(ncclGroupEnd());
    return Status();
}

Status NcclWorld::all_reduce_sum_u64(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        ncclResult_t r = ncclAllReduce(buf[i], buf[i], count, ncclUint64, ncclSum, impl_->comms[i], streams[i]);
        if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r)); }
    }
    SGS_NCCL_TRY(ncclGroupEnd());
    return Status();
}

Status NcclWorld::all_reduce_max_f64(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        ncclResult_t r = ncclAllReduce(buf[i], buf[i], count, ncclDouble, ncclMax, impl_->comms[i], streams[i]);
        if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r)); }
    }
    SGS_NCCL_TRY(ncclGroupEnd());
    return Status();
}

}  // namespace sgs
