// Shared host/device helpers for libvd3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vd3d_b200.h"

namespace vd3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define VD3D_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::vd3d::set_error(__VA_ARGS__);                       \
            return VD3D_EINVAL;                                   \
        }                                                         \
    } while (0)

#define VD3D_CHECK_LAUNCH(name)                                                     \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ::vd3d::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return VD3D_ECUDA;                                                      \
        }                                                                           \
        ::vd3d::count_launch();                                                     \
    } while (0)

#define VD3D_CUDA(call)                                                             \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::vd3d::set_error("%s failed: %s", #call, cudaGetErrorString(e__));     \
            return VD3D_ECUDA;                                                      \
        }                                                                           \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace vd3d
