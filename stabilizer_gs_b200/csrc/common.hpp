// common.hpp — error plumbing shared by the host files of libsgs.
#pragma once
#include <cstdio>
#include <string>

#include "../../include/sgs.h"

namespace sgs {

void set_last_error(const std::string &msg);
const char *get_last_error();

struct Status {
    int code = SGS_OK;
    std::string msg;
    bool ok() const { return code == SGS_OK; }
    static Status fail(int c, const std::string &m) { Status s; s.code = c; s.msg = m; return s; }
};

}  // namespace sgs

#define SGS_TRY(expr) do { ::sgs::Status s__ = (expr); if (!s__.ok()) return s__; } while (0)
#define SGS_CUDA_TRY(expr)                                                                                \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            int code__ = (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ||               \
                          e__ == cudaErrorInvalidDevice)                                                  \
                             ? SGS_ERR_NO_DEVICE                                                          \
                             : SGS_ERR_CUDA;                                                              \
            return ::sgs::Status::fail(code__, std::string(#expr) + ": " + cudaGetErrorString(e__) +      \
                                                   " (" __FILE__ ":" + std::to_string(__LINE__) + ")");   \
        }                                                                                                 \
    } while (0)
