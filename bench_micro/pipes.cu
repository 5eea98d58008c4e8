// Pipe-throughput probes for sm_100a: which FP32 instruction forms issue at what rate per SM
// sub-partition (SMSP). They ground the design of the Chamfer distance tile ("K1").
// Standalone, no torch:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipes pipes.cu
// Prints one JSON line per probe: warp-instructions per cycle per SMSP (clock64) and the SM clock
// inferred from wall time.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

constexpr int ITERS = 2048;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

#define PROBE_PROLOGUE                                         \
    float a[16];                                               \
    _Pragma("unroll") for (int i = 0; i < 16; i++) a[i] = seed + threadIdx.x * 1e-3f + i; \
    float b = seed * 0.5f, c = seed * 0.25f;                   \
    long long t0 = clock64();
#define PROBE_EPILOGUE                                         \
    long long t1 = clock64();                                  \
    float s = 0.f;                                             \
    _Pragma("unroll") for (int i = 0; i < 16; i++) s += a[i];  \
    if (s == 123.456f) out[threadIdx.x] = s;                   \
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;

__global__ void k_ffma(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
    }
    PROBE_EPILOGUE
}
__global__ void k_fadd(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
    }
    (void)c;
    PROBE_EPILOGUE
}
__global__ void k_fmnmx(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]));
    }
    (void)b; (void)c;
    PROBE_EPILOGUE
}
__global__ void k_fmnmx3(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]), "f"(a[(i + 9) & 15]));
    }
    (void)b; (void)c;
    PROBE_EPILOGUE
}
// packed probes: 8 independent 64-bit accumulators, 16 packed instructions per iteration
#define PACKED_PROLOGUE                                                            \
    u64 p[8];                                                                      \
    _Pragma("unroll") for (int i = 0; i < 8; i++) p[i] = pack2(a[2 * i], a[2 * i + 1]); \
    u64 pb = pack2(b, b + 1.f), pc = pack2(c, c + 1.f);
#define PACKED_EPILOGUE _Pragma("unroll") for (int i = 0; i < 8; i++) unpack2(p[i], a[2 * i], a[2 * i + 1]);

__global__ void k_ffma2(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    PACKED_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
    }
    PACKED_EPILOGUE
    PROBE_EPILOGUE
}
__global__ void k_fadd2(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    PACKED_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
    }
    (void)pc;
    PACKED_EPILOGUE
    PROBE_EPILOGUE
}
__global__ void k_fmul2(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    PACKED_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
    }
    (void)pc;
    PACKED_EPILOGUE
    PROBE_EPILOGUE
}
// 8 FFMA2 + 8 FMNMX interleaved: does the ALU-pipe min issue in the shadow of the packed FMA?
__global__ void k_mix_ffma2_fmnmx(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    u64 p[4];
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = pack2(a[2 * i], a[2 * i + 1]);
    u64 pb = pack2(b, b + 1.f), pc = pack2(c, c + 1.f);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[8 + i + 4 * r]) : "f"(a[8 + ((i + 3) & 7)]));
            }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) unpack2(p[i], a[2 * i], a[2 * i + 1]);
    PROBE_EPILOGUE
}
// 8 FFMA + 8 FMNMX interleaved (scalar form of the same question)
__global__ void k_mix_ffma_fmnmx(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(a[8 + i]) : "f"(a[8 + ((i + 3) & 7)]));
        }
    }
    PROBE_EPILOGUE
}
// warp-wide float min through the uniform datapath (CREDUX.MIN.F32), 8 per iteration + the FMNMX that consumes it
__global__ void k_credux(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float r;
            asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(a[i]));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(a[8 + i]) : "f"(r));
        }
    }
    (void)b; (void)c;
    PROBE_EPILOGUE
}
// butterfly min via SHFL: 5 SHFL + 5 FMNMX per value, 2 values per iteration
__global__ void k_shfl_min(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float v = a[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
            a[8 + i] = fminf(a[8 + i], v) + 1.0f;
            a[i] += a[8 + i];
        }
    }
    (void)b; (void)c;
    PROBE_EPILOGUE
}


// 12 FFMA2 + 4 FMNMX3 per iteration: 1 issue slot per FFMA2 -> ~24 cycles (pipe-bound); 2 slots -> ~31
__global__ void k_mixA_12ffma2_4min3(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    u64 p[6];
#pragma unroll
    for (int i = 0; i < 6; i++) p[i] = pack2(a[i], a[i] + 1.f);
    u64 pb = pack2(b, b + 1.f), pc = pack2(c, c + 1.f);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 6; i++) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                if (i % 3 == 0) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[12 + i / 3 + 2 * r]) : "f"(a[8 + i]), "f"(a[9 + i]));
            }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) unpack2(p[i], a[i], a[6]);
    PROBE_EPILOGUE
}
// 24 FFMA + 4 FMNMX3 per iteration (same arithmetic, scalar issue)
__global__ void k_mixB_24ffma_4min3(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    float q[12];
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = a[i];
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(q[i]) : "f"(b), "f"(c));
                if (i % 6 == 0) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[12 + i / 6 + 2 * r]) : "f"(a[8]), "f"(a[9]));
            }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) a[i] = q[i];
    PROBE_EPILOGUE
}
// FMUL2 with operands that stay normal (x *= 1.0)
__global__ void k_fmul2_normal(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    PACKED_PROLOGUE
    pb = pack2(1.0f, 1.0f);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
    }
    (void)pc;
    PACKED_EPILOGUE
    PROBE_EPILOGUE
}
// FSETP + SEL pairs (argmin bookkeeping): 8 + 8 per iteration
__global__ void k_fsetp_sel(float* out, long long* cycles, float seed) {
    PROBE_PROLOGUE
    int id[8];
#pragma unroll
    for (int i = 0; i < 8; i++) id[i] = i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("{ .reg .pred q; setp.lt.f32 q, %1, %2; selp.b32 %0, %3, %0, q; }" : "+r"(id[i]) : "f"(a[i]), "f"(a[8 + i]), "r"(it));
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] += id[i];
    (void)b; (void)c;
    PROBE_EPILOGUE
}

struct Entry { const char* name; void (*fn)(float*, long long*, float); int warp_instr_per_iter; };

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4096 * sizeof(float)));
    CK(cudaMalloc(&cyc, sms * 8 * sizeof(long long)));
    Entry entries[] = {
        {"ffma", k_ffma, 16}, {"fadd", k_fadd, 16}, {"fmnmx", k_fmnmx, 16}, {"fmnmx3", k_fmnmx3, 16},
        {"ffma2", k_ffma2, 16}, {"fadd2", k_fadd2, 16}, {"fmul2", k_fmul2, 16},
        {"mix_ffma2_fmnmx", k_mix_ffma2_fmnmx, 16}, {"mix_ffma_fmnmx", k_mix_ffma_fmnmx, 16},
        {"credux_plus_fmnmx", k_credux, 16}, {"mixA_12ffma2_4min3", k_mixA_12ffma2_4min3, 16}, {"mixB_24ffma_4min3", k_mixB_24ffma_4min3, 28}, {"fmul2_normal", k_fmul2_normal, 16}, {"fsetp_sel", k_fsetp_sel, 16}, {"shfl_min_butterfly", k_shfl_min, 2 * 13},
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    long long* h = (long long*)malloc(sms * 8 * sizeof(long long));
    const int warps_cfg[] = {4, 8, 16};  // warps per CTA, 1 CTA per SM -> 1, 2, 4 warps per SMSP
    for (auto& e : entries) {
        for (int wc : warps_cfg) {
            int threads = wc * 32;
            e.fn<<<sms, threads>>>(out, cyc, 1.0f);  // warm-up
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            e.fn<<<sms, threads>>>(out, cyc, 1.0f);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < sms; i++) mx = h[i] > mx ? h[i] : mx;
            double warp_instr_per_smsp = (double)ITERS * e.warp_instr_per_iter * (wc / 4.0);
            double ipc = warp_instr_per_smsp / (double)mx;
            printf("{\"probe\": \"%s\", \"warps_per_smsp\": %d, \"cycles\": %lld, \"warp_instr_per_cycle_per_smsp\": %.4f, \"ms\": %.4f, \"sm_mhz_est\": %.0f}\n",
                   e.name, wc / 4, mx, ipc, ms, mx / (ms * 1e3));
        }
    }
    return 0;
}
