// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/rf_b200.h"

void rf_set_error(const std::string& msg);
int rf_fail(int code, const std::string& msg);

#define RF_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rf_fail(RF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

#define RF_CUDA_LAUNCH_CHECK(name)                                                             \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return rf_fail(RF_ERR_CUDA, std::string("launch ") + name + ": " +                 \
                                            cudaGetErrorString(_e));                           \
    } while (0)


// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: remember it per (call site,
// device) so that a process driving several GPUs sets it on each of them.  Up to 64 devices.
struct rf_dev_once {
    std::atomic<unsigned long long> done{0};
};
template <class F>
inline cudaError_t rf_set_smem_once(rf_dev_once& o, F* func, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (o.done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) o.done.fetch_or(bit, std::memory_order_release);
    return e;
}
