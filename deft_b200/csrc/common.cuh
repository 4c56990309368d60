// Shared helpers for libdeft_b200.so (sm_100a only).
#pragma once
#include <utility>
#include <stdlib.h>
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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel, so what has been granted is
// remembered per device ordinal (a process may drive several GPUs through this library).  `granted` is the caller's
// function-local static (one per kernel instantiation).  Requests <= 48 KB need no attribute.
constexpr int DEFT_MAX_DEVICES = 64;
template <typename K>
inline int ensure_dyn_smem(K kernel, size_t bytes, size_t (&granted)[DEFT_MAX_DEVICES], const char* what) {
    if (bytes <= 48 * 1024) return DEFT_OK;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= DEFT_MAX_DEVICES) {
        set_error("%s: cannot identify the current device", what);
        return DEFT_ECUDA;
    }
    if (bytes <= granted[d]) return DEFT_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("%s: cannot raise the dynamic shared memory limit to %zu bytes: %s", what, bytes, cudaGetErrorString(e));
        return DEFT_ECUDA;
    }
    granted[d] = bytes;
    return DEFT_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }


// ---- programmatic dependent launch (PDL).  Every kernel on the frame path starts with DEFT_PDL_PROLOGUE(): wait until
// the predecessor grid in the stream has completed and its writes are visible (griddepcontrol.wait, a no-op without the
// launch attribute), then let the successor begin launching while this grid runs.  The wait is the FIRST statement,
// before any early return, so a grid never completes before its predecessor -- ordering stays transitive.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define DEFT_PDL_PROLOGUE() do { deft::pdl_wait(); deft::pdl_launch_dependents(); } while (0)

bool pdl_enabled();            // api.cu: deft_set_pdl() / DEFT_PDL
long long* debug_trace_ptr();  // api.cu: deft_debug_set_trace(); nullptr unless a bring-up tool armed it
// A tensor-core dispatcher that answers DEFT_EINVAL ("shape not covered") makes the caller fall back to the fp32 SIMT
// engine: correct, but ~7x slower.  Every such fallback is counted (deft_simt_fallback_count) and, with
// DEFT_DEBUG_ENGINE=1 / deft_debug_set_engine_log(1), reported on stderr, so a shape regression cannot hide.
void note_simt_fallback(const char* op, const char* detail_fmt, ...);
void note_dcn_patch_launch();  // api.cu: deft_dcn_patch_launch_count()
int dcn_patch_mode();          // api.cu: deft_debug_set_dcn_patch() / DEFT_DCN_PATCH (0 off, 1 on for maps that tile well, 2 on for every map)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace deft
