// Shared helpers for the fpsg_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <utility>

#include "../../include/fpsg_b200.h"

namespace fpsg {

// Records a message for fpsg_last_error() (thread-local) and returns `code`.
int fail(int code, const char* fmt, ...);

#define FPSG_CUDA_OK(expr)                                                                            \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return ::fpsg::fail(FPSG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                                __FILE__, __LINE__);                                                  \
    } while (0)

constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

// The canonical FP32 squared distance (oracle/fpsg_oracle.c sqdist_f32): differences, then
// t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t). The *_rn intrinsics are never re-associated or
// re-contracted by nvcc, so the bits match the oracle's fmaf() chain exactly.
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay), dz = __fsub_rn(bz, az);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

// Launch with programmatic stream serialization (PDL): the grid may start while its predecessor in the stream is
// still draining; the kernel calls cudaGridDependencySynchronize() before it touches the predecessor's output.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// 3-input FP32 min (FMNMX3 on sm_100a): one ALU-pipe slot for two comparisons.
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

}  // namespace fpsg
