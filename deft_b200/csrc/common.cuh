// Shared helpers for libdeft_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "deft_b200.h"

namespace deft {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of this library (bench.py's gpu_launches)

inline cudaStream_t to_stream(deft_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define DEFT_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            deft::set_error(__VA_ARGS__);                         \
            return DEFT_EINVAL;                                   \
        }                                                         \
    } while (0)

#define DEFT_CHECK_LAUNCH(name)                                                   \
    do {                                                                          \
        deft::count_launch();                                                     \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            deft::set_error("%s: %s", name, cudaGetErrorString(e__));             \
            return DEFT_ECUDA;                                                    \
        }                                                                         \
    } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace deft
