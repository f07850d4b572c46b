// nccl_glue.cpp — see nccl_glue.hpp.
#include "nccl_glue.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace sgs {

// NCCL is bound at run time, not link time: inside a process that already carries an NCCL (e.g. a
// torch.distributed launcher, whose bundled libnccl.so.2 is newer than the system one) the loaded copy is
// used; otherwise libnccl.so.2 is opened on first use.  Single-GPU searches never touch NCCL.
namespace {
struct NcclApi {
    decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&::ncclCommInitAll) CommInitAll = nullptr;
    decltype(&::ncclCommInitRank) CommInitRank = nullptr;
    decltype(&::ncclCommDestroy) CommDestroy = nullptr;
    decltype(&::ncclGroupStart) GroupStart = nullptr;
    decltype(&::ncclGroupEnd) GroupEnd = nullptr;
    decltype(&::ncclAllGather) AllGather = nullptr;
    decltype(&::ncclAllReduce) AllReduce = nullptr;
    decltype(&::ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &api()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    void *h = nullptr;
    if (!dlsym(RTLD_DEFAULT, "ncclGetUniqueId")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { a.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return a; }
    }
    auto sym = [&](const char *name) -> void * {
        void *p = h ? dlsym(h, name) : dlsym(RTLD_DEFAULT, name);
        if (!p) a.why = std::string("NCCL symbol missing: ") + name;
        return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitAll && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.AllGather &&
           a.AllReduce && a.GetErrorString;
    return a;
}
}  // namespace

#define ncclGetUniqueId api().GetUniqueId
#define ncclCommInitAll api().CommInitAll
#define ncclCommInitRank api().CommInitRank
#define ncclCommDestroy api().CommDestroy
#define ncclGroupStart api().GroupStart
#define ncclGroupEnd api().GroupEnd
#define ncclAllGather api().AllGather
#define ncclAllReduce api().AllReduce
#define ncclGetErrorString api().GetErrorString
#define SGS_NEED_NCCL() do { if (!api().ok) return Status::fail(SGS_ERR_NCCL, api().why.empty() ? "NCCL unavailable" : api().why); } while (0)

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
    for (size_t i = 0; api().ok && i < impl_->comms.size(); i++) {
        if (!impl_->comms[i]) continue;
        cudaSetDevice(impl_->devices[i]);
        ncclCommDestroy(impl_->comms[i]);
    }
    delete impl_;
}

Status NcclWorld::unique_id(void *out128)
{
    static_assert(sizeof(ncclUniqueId) <= SGS_NCCL_ID_BYTES, "ncclUniqueId larger than SGS_NCCL_ID_BYTES");
    SGS_NEED_NCCL();
    ncclUniqueId id;
    SGS_NCCL_TRY(ncclGetUniqueId(&id));
    memset(out128, 0, SGS_NCCL_ID_BYTES);
    memcpy(out128, &id, sizeof(id));
    return Status();
}

Status NcclWorld::init(const std::vector<int> &devices, int rank, int world, const void *unique_id)
{
    SGS_NEED_NCCL();
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
