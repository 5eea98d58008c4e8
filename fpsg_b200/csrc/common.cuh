// Shared helpers for the fpsg_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdlib>
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

// Tuning switches (FPSG_* environment variables used by the ablations under bench_micro/) exist only in builds made
// with -DFPSG_TUNING (FPSG_BUILD_TUNING=1 python fpsg_b200/build.py). The product build compiles them out: no launch
// path reads the environment.
#ifdef FPSG_TUNING
inline const char* tune_env(const char* name) { return std::getenv(name); }
#else
inline const char* tune_env(const char*) { return nullptr; }
#endif

// Timeline stamps (tuning builds only): kernels write %globaltimer values into a buffer registered with
// fpsg_tune_set_trace (api.cu; not part of the product ABI), laid out [kernel][CTA][phase]. bench_micro/timeline.py
// turns them into the step anatomy kept under profiles/.
constexpr int kTraceCtas = 4096;
constexpr int kTracePhases = 8;
unsigned long long* trace_buffer();
#if defined(FPSG_TUNING) && defined(__CUDACC__)
#define FPSG_TRACE(ptr, kernel_id, cta, phase)                                                                   \
    do {                                                                                                         \
        if ((ptr) != nullptr && threadIdx.x == 0 && (cta) < ::fpsg::kTraceCtas) {                                \
            unsigned long long t_;                                                                               \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                               \
            (ptr)[((size_t)(kernel_id) * ::fpsg::kTraceCtas + (cta)) * ::fpsg::kTracePhases + (phase)] = t_;      \
        }                                                                                                        \
    } while (0)
#else
#define FPSG_TRACE(ptr, kernel_id, cta, phase) \
    do {                                       \
    } while (0)
#endif

constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

// A batch made of up to FPSG_MAX_SEGMENTS separately allocated runs of pairs ("segments": the query-set and
// support-set calls of the reference's training loss, src/models/few_shot.py:109-117). Pair b of the batch lives in
// segment s = the last one with b0[s] <= b, at local index b - b0[s]. A plain (B,N,3)/(B,M,3) call is one segment.
struct SegTable {
    const float* p1[FPSG_MAX_SEGMENTS];
    const float* p2[FPSG_MAX_SEGMENTS];
    float* gp1[FPSG_MAX_SEGMENTS];  // backward only
    float* gp2[FPSG_MAX_SEGMENTS];
    int b0[FPSG_MAX_SEGMENTS + 1];  // b0[n] = total number of pairs
    float weight[FPSG_MAX_SEGMENTS];
    int n;
};

#ifdef __CUDACC__
__device__ __forceinline__ int seg_find(const SegTable& t, int b) {
    int s = 0;
#pragma unroll
    for (int q = 1; q < FPSG_MAX_SEGMENTS; ++q) s += (q < t.n && b >= t.b0[q]) ? 1 : 0;
    return s;
}
// cloud b of the p1 / p2 side (npts points of 3 floats per cloud)
__device__ __forceinline__ const float* seg_p1(const SegTable& t, int b, int npts) {
    if (t.n == 1) return t.p1[0] + (size_t)b * npts * 3;
    const int s = seg_find(t, b);
    return t.p1[s] + (size_t)(b - t.b0[s]) * npts * 3;
}
__device__ __forceinline__ const float* seg_p2(const SegTable& t, int b, int npts) {
    if (t.n == 1) return t.p2[0] + (size_t)b * npts * 3;
    const int s = seg_find(t, b);
    return t.p2[s] + (size_t)(b - t.b0[s]) * npts * 3;
}
#endif

inline SegTable single_segment(const float* p1, const float* p2, float* gp1, float* gp2, int B) {
    SegTable t = {};
    t.p1[0] = p1;
    t.p2[0] = p2;
    t.gp1[0] = gp1;
    t.gp2[0] = gp2;
    t.b0[0] = 0;
    t.b0[1] = B;
    t.weight[0] = 1.0f;
    t.n = 1;
    return t;
}

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
