// canonicalvoting_b200/csrc/common.cuh -- shared host/device helpers of libcvb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cvb200.h"

namespace cvb200 {

// thread-local error text behind cvb200_last_error()
void set_error(const char *fmt, ...);

inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

#define CVB_CUDA(expr)                                          \
    do {                                                        \
        int _rc = ::cvb200::check_cuda((expr), #expr);          \
        if (_rc) return _rc;                                    \
    } while (0)

#define CVB_LAUNCH_CHECK(name)                                            \
    do {                                                                  \
        int _rc = ::cvb200::check_cuda(cudaGetLastError(), name);         \
        if (_rc) return _rc;                                              \
    } while (0)

#define CVB_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ::cvb200::set_error(__VA_ARGS__);   \
            return (code);                      \
        }                                       \
    } while (0)

constexpr int kNumSMs = 148;  // B200

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace cvb200
