// common.hpp -- shared helpers of the lsio_b200 native library (product code; no oracle).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace lsio {

// Last CUDA error text seen on this thread (returned by lsio_cuda_last_error()).
inline std::string& last_error_text() {
    thread_local std::string s;
    return s;
}
inline std::atomic<uint64_t>& launch_counter() {
    static std::atomic<uint64_t> c{0};
    return c;
}

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s) at %s:%d", what, cudaGetErrorName(e), cudaGetErrorString(e), file, line);
    last_error_text() = buf;
    return (int)e;  // positive cudaError_t
}

#define LSIO_CUDA_TRY(expr)                                                          \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) return ::lsio::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// RAII device switch: the library never leaves the caller's current device changed.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace lsio
