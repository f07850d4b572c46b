// nccl_glue.cpp — see nccl_glue.hpp.
#include "nccl_glue.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>

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
    decltype(&::ncclBroadcast) Broadcast = nullptr;
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
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitAll && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.AllGather &&
           a.AllReduce && a.Broadcast && a.GetErrorString;
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
#define ncclBroadcast api().Broadcast
#define ncclGetErrorString api().GetErrorString
#define SGS_NEED_NCCL() do { if (!api().ok) return Status::fail(SGS_ERR_NCCL, api().why.empty() ? "NCCL unavailable" : api().why); } while (0)

struct NcclWorld::Impl {
    std::vector<ncclComm_t> comms;
    std::vector<int> devices;
    int rank0 = 0, nranks = 1;      // global rank of local device 0, communicator size
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
        impl_->rank0 = 0; impl_->nranks = ng;
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
    impl_->rank0 = rank * ng; impl_->nranks = world * ng;
    return Status();
}

int NcclWorld::global_rank(int local) const { return impl_->rank0 + local; }
int NcclWorld::size() const { return impl_->nranks; }

Status NcclWorld::get(const std::vector<int> &devices, int rank, int world, const void *unique_id, std::shared_ptr<NcclWorld> &out)
{
    std::string key(unique_id ? (const char *)unique_id : "", unique_id ? SGS_NCCL_ID_BYTES : 0);
    key += "|" + std::to_string(rank) + "|" + std::to_string(world);
    for (int d : devices) key += "," + std::to_string(d);
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<NcclWorld>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { out = it->second; return Status(); }
    std::shared_ptr<NcclWorld> w(new NcclWorld());
    Status s = w->init(devices, rank, world, unique_id);
    if (!s.ok()) return s;
    cache[key] = w;
    out = w;
    return Status();
}

Status NcclWorld::gather_arrays_and_max(const std::vector<std::vector<void *>> &base, const std::vector<size_t> &bytes,
                                        const std::vector<void *> &max_word, const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        const int g = impl_->rank0 + (int)i;
        for (size_t a = 0; a < bytes.size(); a++) {
            if (bytes[a] == 0) continue;
            char *recv = (char *)base[i][a];
            ncclResult_t r = ncclAllGather(recv + (size_t)g * bytes[a], recv, bytes[a], ncclChar, impl_->comms[i], streams[i]);
            if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclAllGather: ") + ncclGetErrorString(r)); }
        }
        ncclResult_t r = ncclAllReduce(max_word[i], max_word[i], 1, ncclInt, ncclMax, impl_->comms[i], streams[i]);
        if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r)); }
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

// several arrays of every device overwritten with rank 0's copy (one NCCL group)
Status NcclWorld::broadcast_arrays(const std::vector<std::vector<void *>> &bufs, const std::vector<size_t> &bytes, const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++)
        for (size_t a = 0; a < bytes.size(); a++) {
            if (bytes[a] == 0) continue;
            ncclResult_t r = ncclBroadcast(bufs[i][a], bufs[i][a], bytes[a], ncclChar, 0, impl_->comms[i], streams[i]);
            if (r != ncclSuccess) { ncclGroupEnd(); return Status::fail(SGS_ERR_NCCL, std::string("ncclBroadcast: ") + ncclGetErrorString(r)); }
        }
    SGS_NCCL_TRY(ncclGroupEnd());
    return Status();
}

Status NcclWorld::all_reduce_sum_i32(const std::vector<void *> &buf, size_t count, const std::vector<cudaStream_t> &streams)
{
    SGS_NCCL_TRY(ncclGroupStart());
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        ncclResult_t r = ncclAllReduce(buf[i], buf[i], count, ncclInt, ncclSum, impl_->comms[i], streams[i]);
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
