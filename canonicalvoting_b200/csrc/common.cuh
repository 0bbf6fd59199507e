// canonicalvoting_b200/csrc/common.cuh -- shared host/device helpers of libcvb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

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

// One-time per-DEVICE setup at a call site (function attributes such as the dynamic shared-memory limit or the non-portable
// cluster size belong to a device's context: a process that touches a second GPU must set them there too).  Thread-safe;
// two racing first calls may both do the setup, which is harmless.
//     static DeviceOnce once;  if (!once.done()) { CVB_CUDA(cudaFuncSetAttribute(...)); once.mark(); }
struct DeviceOnce {
    std::atomic<unsigned long long> bits[4];
    DeviceOnce() { for (auto &b : bits) b.store(0); }
    static int device() { int d = 0; cudaGetDevice(&d); return d & 255; }
    bool done() const { const int d = device(); return (bits[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ull; }
    void mark() { const int d = device(); bits[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace cvb200
