// Steady-state cost of the Chamfer sweep's inner loop, isolated from memory traffic and tile bookkeeping.
// Each variant runs the same arithmetic (16 rows x 4 columns per step, row minima + column minima) from registers
// and a tiny shared-memory column tile; reports SM cycles per step and warp-instructions/cycle/SMSP equivalents.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o loop_probe loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math_constants.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
constexpr int R = 16;
constexpr int STEPS = 2048;
constexpr int TILE = 64;  // columns in the shared tile, reused cyclically

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay), dz = __fsub_rn(bz, az);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float warp_min_f32(float v) { float m; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v)); return m; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// MODE 0: scalar math, no column reduce.  1: scalar + CREDUX/ballot/STS (the shipped loop).
// MODE 2: packed f32x2 math (rows paired), no column reduce.  3: packed + CREDUX/ballot/STS.
template <int MODE>
__global__ void __launch_bounds__(128, (MODE >= 2 ? 3 : 4)) loop_kernel(const float* __restrict__ rows, const float* __restrict__ cols, float* out, long long* cycles) {
    __shared__ __align__(16) float tile[TILE * 3];
    __shared__ __align__(16) float tile2[TILE * 6];  // duplicated coordinates for the packed variant: x x y y z z
    __shared__ u64 res[4][TILE];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < TILE * 3; i += 128) { tile[i] = cols[i]; tile2[2 * i] = cols[i]; tile2[2 * i + 1] = cols[i]; }
    float px[R], py[R], pz[R], best[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { px[k] = rows[(tid * R + k) * 3]; py[k] = rows[(tid * R + k) * 3 + 1]; pz[k] = rows[(tid * R + k) * 3 + 2]; best[k] = CUDART_INF_F; }
    __syncthreads();
    const long long t0 = clock64();
    if (MODE < 2) {
#pragma unroll 1
        for (int s = 0; s < STEPS; ++s) {
            const int step = s & 15;
            const float4 v0 = *reinterpret_cast<const float4*>(tile + 12 * step);
            const float4 v1 = *reinterpret_cast<const float4*>(tile + 12 * step + 4);
            const float4 v2 = *reinterpret_cast<const float4*>(tile + 12 * step + 8);
            float c0m = CUDART_INF_F, c1m = CUDART_INF_F, c2m = CUDART_INF_F, c3m = CUDART_INF_F;
#pragma unroll
            for (int k = 0; k < R; k += 2) {
                const float a0 = sqdist(px[k], py[k], pz[k], v0.x, v0.y, v0.z);
                const float a1 = sqdist(px[k], py[k], pz[k], v0.w, v1.x, v1.y);
                const float a2 = sqdist(px[k], py[k], pz[k], v1.z, v1.w, v2.x);
                const float a3 = sqdist(px[k], py[k], pz[k], v2.y, v2.z, v2.w);
                const float b0 = sqdist(px[k + 1], py[k + 1], pz[k + 1], v0.x, v0.y, v0.z);
                const float b1 = sqdist(px[k + 1], py[k + 1], pz[k + 1], v0.w, v1.x, v1.y);
                const float b2 = sqdist(px[k + 1], py[k + 1], pz[k + 1], v1.z, v1.w, v2.x);
                const float b3 = sqdist(px[k + 1], py[k + 1], pz[k + 1], v2.y, v2.z, v2.w);
                best[k] = min3(min3(best[k], a0, a1), a2, a3);
                best[k + 1] = min3(min3(best[k + 1], b0, b1), b2, b3);
                c0m = min3(c0m, a0, b0); c1m = min3(c1m, a1, b1); c2m = min3(c2m, a2, b2); c3m = min3(c3m, a3, b3);
            }
            if (MODE == 1) {
                const float m0 = warp_min_f32(c0m), m1 = warp_min_f32(c1m), m2 = warp_min_f32(c2m), m3 = warp_min_f32(c3m);
                const unsigned h0 = __ballot_sync(0xffffffffu, c0m == m0), h1 = __ballot_sync(0xffffffffu, c1m == m1);
                const unsigned h2 = __ballot_sync(0xffffffffu, c2m == m2), h3 = __ballot_sync(0xffffffffu, c3m == m3);
                if (lane == 0) {
                    res[w][4 * step + 0] = ((u64)__float_as_uint(m0) << 32) | h0; res[w][4 * step + 1] = ((u64)__float_as_uint(m1) << 32) | h1;
                    res[w][4 * step + 2] = ((u64)__float_as_uint(m2) << 32) | h2; res[w][4 * step + 3] = ((u64)__float_as_uint(m3) << 32) | h3;
                }
            } else {
                best[0] = min3(best[0], c0m, c1m); best[1] = min3(best[1], c2m, c3m);
            }
        }
    } else {
        // rows paired: P[k2] = (p[2k2], p[2k2+1]); column coordinates duplicated (q, q) straight from shared memory
        u64 PX[R / 2], PY[R / 2], PZ[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) { PX[k] = pack2(px[2 * k], px[2 * k + 1]); PY[k] = pack2(py[2 * k], py[2 * k + 1]); PZ[k] = pack2(pz[2 * k], pz[2 * k + 1]); }
#pragma unroll 1
        for (int s = 0; s < STEPS; ++s) {
            const int step = s & 15;
            float cm[4] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
            u64 QX[4], QY[4], QZ[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const ulonglong2 xy = *reinterpret_cast<const ulonglong2*>(tile2 + (4 * step + c) * 6);  // (x,x) (y,y)
                QX[c] = xy.x; QY[c] = xy.y;
                QZ[c] = *reinterpret_cast<const u64*>(tile2 + (4 * step + c) * 6 + 4);
            }
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
                float lo[4], hi[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const u64 dx = sub2(QX[c], PX[k]), dy = sub2(QY[c], PY[k]), dz = sub2(QZ[c], PZ[k]);
                    u64 t = mul2(dy, dy);
                    t = fma2(dx, dx, t);
                    t = fma2(dz, dz, t);
                    unpack2(t, lo[c], hi[c]);
                    cm[c] = min3(cm[c], lo[c], hi[c]);
                }
                best[2 * k] = min3(min3(best[2 * k], lo[0], lo[1]), lo[2], lo[3]);
                best[2 * k + 1] = min3(min3(best[2 * k + 1], hi[0], hi[1]), hi[2], hi[3]);
            }
            if (MODE == 3) {
                const float m0 = warp_min_f32(cm[0]), m1 = warp_min_f32(cm[1]), m2 = warp_min_f32(cm[2]), m3 = warp_min_f32(cm[3]);
                const unsigned h0 = __ballot_sync(0xffffffffu, cm[0] == m0), h1 = __ballot_sync(0xffffffffu, cm[1] == m1);
                const unsigned h2 = __ballot_sync(0xffffffffu, cm[2] == m2), h3 = __ballot_sync(0xffffffffu, cm[3] == m3);
                if (lane == 0) {
                    res[w][4 * step + 0] = ((u64)__float_as_uint(m0) << 32) | h0; res[w][4 * step + 1] = ((u64)__float_as_uint(m1) << 32) | h1;
                    res[w][4 * step + 2] = ((u64)__float_as_uint(m2) << 32) | h2; res[w][4 * step + 3] = ((u64)__float_as_uint(m3) << 32) | h3;
                }
            } else {
                best[0] = min3(best[0], cm[0], cm[1]); best[1] = min3(best[1], cm[2], cm[3]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) s += best[k];
    __syncthreads();
    s += (float)(res[w][lane] & 0xff);
    out[blockIdx.x * 128 + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// MODE 4/5: expanded-form FILTER loop (round 2): e = fma(a'x,bx, fma(a'y,by, fma(a'z,bz, bb))) + aa with a' = -2a: 3 FFMA + 1 FADD
// per evaluation; row direction min3 into a piece minimum; column direction thread minimum, CREDUX, and a ballot of the lanes
// within a tolerance of the warp minimum (one FADD + FSETP per column). RR rows per thread (16 or 8).
template <int RR, int CTAS>
__global__ void __launch_bounds__(128, CTAS) filter_kernel(const float* __restrict__ rows, const float* __restrict__ cols, float* out, long long* cycles, float tol) {
    __shared__ __align__(16) float tile[TILE * 4];  // (x, y, z, bb) per column
    __shared__ u64 res[4][TILE];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < TILE; i += 128) {
        const float x = cols[3 * i], y = cols[3 * i + 1], z = cols[3 * i + 2];
        tile[4 * i] = x; tile[4 * i + 1] = y; tile[4 * i + 2] = z; tile[4 * i + 3] = x * x + y * y + z * z;
    }
    float ax[RR], ay[RR], az[RR], aa[RR], pm[RR];
#pragma unroll
    for (int k = 0; k < RR; ++k) {
        const float x = rows[(tid * RR + k) * 3], y = rows[(tid * RR + k) * 3 + 1], z = rows[(tid * RR + k) * 3 + 2];
        ax[k] = -2.f * x; ay[k] = -2.f * y; az[k] = -2.f * z; aa[k] = x * x + y * y + z * z; pm[k] = CUDART_INF_F;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int s = 0; s < STEPS; ++s) {
        const int step = s & 15;
        const float4 q0 = *reinterpret_cast<const float4*>(tile + 16 * step);
        const float4 q1 = *reinterpret_cast<const float4*>(tile + 16 * step + 4);
        const float4 q2 = *reinterpret_cast<const float4*>(tile + 16 * step + 8);
        const float4 q3 = *reinterpret_cast<const float4*>(tile + 16 * step + 12);
        float c0m = CUDART_INF_F, c1m = CUDART_INF_F, c2m = CUDART_INF_F, c3m = CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < RR; k += 2) {
            const float a0 = __fadd_rn(__fmaf_rn(ax[k], q0.x, __fmaf_rn(ay[k], q0.y, __fmaf_rn(az[k], q0.z, q0.w))), aa[k]);
            const float a1 = __fadd_rn(__fmaf_rn(ax[k], q1.x, __fmaf_rn(ay[k], q1.y, __fmaf_rn(az[k], q1.z, q1.w))), aa[k]);
            const float a2 = __fadd_rn(__fmaf_rn(ax[k], q2.x, __fmaf_rn(ay[k], q2.y, __fmaf_rn(az[k], q2.z, q2.w))), aa[k]);
            const float a3 = __fadd_rn(__fmaf_rn(ax[k], q3.x, __fmaf_rn(ay[k], q3.y, __fmaf_rn(az[k], q3.z, q3.w))), aa[k]);
            const float b0 = __fadd_rn(__fmaf_rn(ax[k + 1], q0.x, __fmaf_rn(ay[k + 1], q0.y, __fmaf_rn(az[k + 1], q0.z, q0.w))), aa[k + 1]);
            const float b1 = __fadd_rn(__fmaf_rn(ax[k + 1], q1.x, __fmaf_rn(ay[k + 1], q1.y, __fmaf_rn(az[k + 1], q1.z, q1.w))), aa[k + 1]);
            const float b2 = __fadd_rn(__fmaf_rn(ax[k + 1], q2.x, __fmaf_rn(ay[k + 1], q2.y, __fmaf_rn(az[k + 1], q2.z, q2.w))), aa[k + 1]);
            const float b3 = __fadd_rn(__fmaf_rn(ax[k + 1], q3.x, __fmaf_rn(ay[k + 1], q3.y, __fmaf_rn(az[k + 1], q3.z, q3.w))), aa[k + 1]);
            pm[k] = min3(min3(pm[k], a0, a1), a2, a3);
            pm[k + 1] = min3(min3(pm[k + 1], b0, b1), b2, b3);
            c0m = min3(c0m, a0, b0); c1m = min3(c1m, a1, b1); c2m = min3(c2m, a2, b2); c3m = min3(c3m, a3, b3);
        }
        const float m0 = warp_min_f32(c0m), m1 = warp_min_f32(c1m), m2 = warp_min_f32(c2m), m3 = warp_min_f32(c3m);
        const unsigned h0 = __ballot_sync(0xffffffffu, c0m <= m0 + tol), h1 = __ballot_sync(0xffffffffu, c1m <= m1 + tol);
        const unsigned h2 = __ballot_sync(0xffffffffu, c2m <= m2 + tol), h3 = __ballot_sync(0xffffffffu, c3m <= m3 + tol);
        if (lane == 0) {
            res[w][4 * step + 0] = ((u64)__float_as_uint(m0) << 32) | h0; res[w][4 * step + 1] = ((u64)__float_as_uint(m1) << 32) | h1;
            res[w][4 * step + 2] = ((u64)__float_as_uint(m2) << 32) | h2; res[w][4 * step + 3] = ((u64)__float_as_uint(m3) << 32) | h3;
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < RR; ++k) s += pm[k];
    __syncthreads();
    s += (float)(res[w][lane] & 0xff);
    out[blockIdx.x * 128 + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int RR, int CTAS>
void run_filter(const char* name, int sms, const float* rows, const float* cols, float* out, long long* cyc) {
    const int grid = sms * CTAS;
    filter_kernel<RR, CTAS><<<grid, 128>>>(rows, cols, out, cyc, 1e-6f);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    filter_kernel<RR, CTAS><<<grid, 128>>>(rows, cols, out, cyc, 1e-6f);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double evals = (double)grid * 128 * RR * 4 * STEPS;
    printf("{\"loop\": \"%s\", \"rows_per_thread\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tera_evals_per_s\": %.3f}\n", name, RR, CTAS, ms, evals / (ms * 1e-3) / 1e12);
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int sms, const float* rows, const float* cols, float* out, long long* cyc) {
    const int grid = sms * ctas_per_sm;
    loop_kernel<MODE><<<grid, 128>>>(rows, cols, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    loop_kernel<MODE><<<grid, 128>>>(rows, cols, out, cyc);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(grid * sizeof(long long));
    CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0; for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; } avg /= grid;
    free(h);
    // all ctas_per_sm CTAs (1 warp per SMSP each) run concurrently: cycles per step per SMSP-resident warp set
    const double evals = (double)grid * 128 * R * 4 * STEPS;
    printf("{\"loop\": \"%s\", \"ctas_per_sm\": %d, \"cycles_per_step_avg\": %.1f, \"cycles_per_step_max\": %.1f, \"issue_cycles_per_eval\": %.3f, \"ms\": %.4f, \"tflops_16flop_per_pair\": %.2f}\n",
           name, ctas_per_sm, avg / STEPS, (double)mx / STEPS, (avg / STEPS) / (R * 4.0 * ctas_per_sm), ms, evals * 8.0 / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *rows, *cols, *out; long long* cyc;
    CK(cudaMalloc(&rows, 128 * R * 3 * sizeof(float))); CK(cudaMalloc(&cols, TILE * 3 * sizeof(float)));
    CK(cudaMalloc(&out, sms * 8 * 128 * sizeof(float))); CK(cudaMalloc(&cyc, sms * 8 * sizeof(long long)));
    float* h = (float*)malloc(128 * R * 3 * sizeof(float));
    for (int i = 0; i < 128 * R * 3; ++i) h[i] = (float)rand() / RAND_MAX;
    CK(cudaMemcpy(rows, h, 128 * R * 3 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cols, h, TILE * 3 * sizeof(float), cudaMemcpyHostToDevice));
    for (int c = 1; c <= 4; ++c) {
        run<0>("scalar", c, sms, rows, cols, out, cyc);
        run<1>("scalar+colreduce", c, sms, rows, cols, out, cyc);
        if (c <= 3) { run<2>("packed", c, sms, rows, cols, out, cyc); run<3>("packed+colreduce", c, sms, rows, cols, out, cyc); }
    }
    // reference point in the same units: the shipped loop at 4 CTAs/SM
    {
        const int grid = sms * 4;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        loop_kernel<1><<<grid, 128>>>(rows, cols, out, cyc); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); loop_kernel<1><<<grid, 128>>>(rows, cols, out, cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"loop\": \"exact (shipped)\", \"rows_per_thread\": 16, \"ctas_per_sm\": 4, \"ms\": %.4f, \"tera_evals_per_s\": %.3f}\n", ms, (double)grid * 128 * R * 4 * STEPS / (ms * 1e-3) / 1e12);
    }
    run_filter<16, 3>("filter", sms, rows, cols, out, cyc);
    run_filter<16, 4>("filter", sms, rows, cols, out, cyc);
    run_filter<8, 4>("filter", sms, rows, cols, out, cyc);
    run_filter<8, 5>("filter", sms, rows, cols, out, cyc);
    run_filter<8, 6>("filter", sms, rows, cols, out, cyc);
    return 0;
}
