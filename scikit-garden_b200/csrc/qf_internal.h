// Internal helpers shared by the host and device translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/qforest_b200.h"

namespace qf {

// Thread-local last-error slot behind qf_last_error().
std::string& last_error_slot();

inline int fail(int code, const char* msg) {
    last_error_slot() = msg;
    return code;
}

inline int fail(int code, const std::string& msg) {
    last_error_slot() = msg;
    return code;
}

}  // namespace qf

#define QF_CUDA_TRY(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _buf[512];                                                                  \
            std::snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr,               \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                       \
            return qf::fail(QF_ERR_CUDA, _buf);                                              \
        }                                                                                    \
    } while (0)
