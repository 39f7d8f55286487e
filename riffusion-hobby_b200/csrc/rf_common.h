// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

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
