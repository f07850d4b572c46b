// arena.hpp — process-wide cache of device scratch arenas, one set per search family and device: repeated
// one-shot searches reuse the device allocation instead of paying cudaMalloc/cudaFree on every call.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <vector>

namespace sgs {

struct ArenaCache {
    struct Slot { void *p; size_t bytes; bool busy; };
    std::mutex mu;
    std::map<int, std::vector<Slot>> slots;
    cudaError_t acquire(int dev, size_t bytes, void **out, size_t *got = nullptr)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto &v = slots[dev];
        for (auto &s : v)
            if (!s.busy && s.bytes >= bytes) { s.busy = true; *out = s.p; if (got) *got = s.bytes; return cudaSuccess; }
        for (size_t i = 0; i < v.size();)          // drop idle arenas that are too small
            if (!v[i].busy) { cudaFree(v[i].p); v.erase(v.begin() + i); } else i++;
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
        v.push_back({p, bytes, true});
        *out = p;
        if (got) *got = bytes;
        return cudaSuccess;
    }
    void release(int dev, void *p)
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &s : slots[dev]) if (s.p == p) s.busy = false;
    }
};

inline ArenaCache &sparse_arena() { static ArenaCache a; return a; }
inline ArenaCache &klocal_arena() { static ArenaCache a; return a; }

}  // namespace sgs
