// K3: kNN graph for the DGCNN encoder, sm_100a.
//
// Replaces `knn(x, k)` (reference src/dgcnn/model.py:13-20): a batched matmul, two broadcast passes that
// materialise three (B,N,N) FP32 temporaries, and a generic radix top-k. Here nothing of size N x N ever reaches
// memory: a warp owns a query point, scores 32 candidates per step straight from a shared-memory tile of the
// cloud, and keeps the running top-k as a sorted list spread over its lanes (entry e lives in lane e % 32,
// register e / 32). A candidate enters the list only when it beats the current k-th value, which after the first
// few steps is rare, so the common step is score + compare + one ballot.
//
// Arithmetic follows the reference expression bit for bit (oracle/fpsg_oracle.c oracle_knn_f32):
//   dot  = fma chain over ascending c from 0            (model.py:14, the matmul)
//   xx_j = sum over ascending c of rn(x_cj^2)           (model.py:15)
//   pd   = rn( rn(-xx_j + 2*dot) - xx_i )               (model.py:16; -2*dot is exact, so the first
//                                                        subtraction is one rounding = fma(2, dot, -xx_j))
// Ranking: larger pd first; equal pd ordered by ascending index (torch.topk leaves ties unspecified).
#include "common.cuh"

namespace fpsg {

constexpr int kKnnThreads = 256;
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnRowsPerWarp = 8;
constexpr int kKnnRows = kKnnWarps * kKnnRowsPerWarp;  // query points per CTA
constexpr size_t kKnnTileBytes = 64 * 1024;            // shared-memory budget for the candidate tile

// squared norms, one thread per point: xx[b,j] = sum_c rn(x^2), ascending c
__global__ void __launch_bounds__(256) knn_sqnorm_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ xx) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= N) return;
    const float* xb = x + (size_t)b * C * N;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = __ldg(xb + (size_t)c * N + j);
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    xx[(size_t)b * N + j] = s;
}

// KR = list registers per lane (k <= 32*KR); C3 = channel count is exactly 3 (query point kept in registers)
template <int KR, bool C3>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int k, int J, int64_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sTile = reinterpret_cast<float*>(smem_raw);           // [C][J] candidate coordinates
    float* sXX = sTile + (size_t)C * J;                          // [J]    candidate squared norms
    float* sRow = sXX + J;                                       // [kKnnWarps][C] the warp's query point (generic C)
    float* sListV = sRow + (C3 ? 0 : kKnnWarps * C);             // [kKnnRows][32*KR] list values
    int* sListI = reinterpret_cast<int*>(sListV + kKnnRows * 32 * KR);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y;
    const int rowbase = blockIdx.x * kKnnRows + w * kKnnRowsPerWarp;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    const int kth_lane = (k - 1) & 31, kth_reg = (k - 1) >> 5;

    for (int j0 = 0; j0 < N; j0 += J) {
        const int jn = min(J, N - j0);
        __syncthreads();  // previous tile fully consumed
        for (int c = 0; c < C; ++c)
            for (int jj = tid; jj < jn; jj += kKnnThreads) sTile[(size_t)c * J + jj] = __ldg(xb + (size_t)c * N + j0 + jj);
        for (int jj = tid; jj < jn; jj += kKnnThreads) sXX[jj] = __ldg(xxb + j0 + jj);
        __syncthreads();

        for (int rr = 0; rr < kKnnRowsPerWarp; ++rr) {
            const int i = rowbase + rr;
            if (i >= N) break;  // warp-uniform
            const int slot = (w * kKnnRowsPerWarp + rr) * 32 * KR;
            float lv[KR];
            int li[KR];
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                if (j0 == 0) {
                    lv[r] = -CUDART_INF_F;
                    li[r] = 0;
                } else {
                    lv[r] = sListV[slot + r * 32 + lane];
                    li[r] = sListI[slot + r * 32 + lane];
                }
            }
            float tau = -CUDART_INF_F;
#pragma unroll
            for (int r = 0; r < KR; ++r)
                if (r == kth_reg) tau = __shfl_sync(0xffffffffu, lv[r], kth_lane);

            float q0 = 0.f, q1 = 0.f, q2 = 0.f;
            if (C3) {
                q0 = __ldg(xb + i);
                q1 = __ldg(xb + (size_t)N + i);
                q2 = __ldg(xb + 2 * (size_t)N + i);
            } else {
                __syncwarp();
                for (int c = lane; c < C; c += 32) sRow[w * C + c] = __ldg(xb + (size_t)c * N + i);
                __syncwarp();
            }
            const float xxi = __ldg(xxb + i);

            for (int jb = 0; jb < jn; jb += 32) {
                const int jj = jb + lane;
                float pd = -CUDART_INF_F;
                if (jj < jn) {
                    float dot;
                    if (C3) {
                        dot = __fmul_rn(q0, sTile[jj]);
                        dot = __fmaf_rn(q1, sTile[J + jj], dot);
                        dot = __fmaf_rn(q2, sTile[2 * J + jj], dot);
                    } else {
                        dot = 0.f;
                        for (int c = 0; c < C; ++c) dot = __fmaf_rn(sRow[w * C + c], sTile[(size_t)c * J + jj], dot);
                    }
                    pd = __fsub_rn(__fmaf_rn(2.0f, dot, -sXX[jj]), xxi);
                }
                unsigned mask = __ballot_sync(0xffffffffu, pd > tau);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float v = __shfl_sync(0xffffffffu, pd, src);
                    if (!(v > tau)) continue;  // tau rose since the ballot
                    const int vj = j0 + jb + src;
                    // position = number of entries that stay ahead (value >= v: all of them have lower index)
                    int pos = 0;
#pragma unroll
                    for (int r = 0; r < KR; ++r) pos += __popc(__ballot_sync(0xffffffffu, lv[r] >= v));
#pragma unroll
                    for (int r = KR - 1; r >= 0; --r) {
                        float pv = __shfl_up_sync(0xffffffffu, lv[r], 1);
                        int pi = __shfl_up_sync(0xffffffffu, li[r], 1);
                        if (r > 0) {
                            const float cv = __shfl_sync(0xffffffffu, lv[r - 1], 31);
                            const int ci = __shfl_sync(0xffffffffu, li[r - 1], 31);
                            if (lane == 0) {
                                pv = cv;
                                pi = ci;
                            }
                        }
                        const int e = r * 32 + lane;
                        if (e > pos) {
                            lv[r] = pv;
                            li[r] = pi;
                        } else if (e == pos) {
                            lv[r] = v;
                            li[r] = vj;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < KR; ++r)
                        if (r == kth_reg) tau = __shfl_sync(0xffffffffu, lv[r], kth_lane);
                }
            }

            if (j0 + J >= N) {
#pragma unroll
                for (int r = 0; r < KR; ++r) {
                    const int e = r * 32 + lane;
                    if (e < k) idx[((size_t)b * N + i) * k + e] = li[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < KR; ++r) {
                    sListV[slot + r * 32 + lane] = lv[r];
                    sListI[slot + r * 32 + lane] = li[r];
                }
            }
        }
    }
}

struct KnnPlan {
    int KR;
    int J;
    size_t smem;
    size_t ws_bytes;
};

static KnnPlan make_knn_plan(int B, int C, int N, int k) {
    KnnPlan p;
    p.KR = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    long long j = (long long)(kKnnTileBytes / sizeof(float)) / (C + 1);
    j = j / 32 * 32;
    if (j < 32) j = 32;
    const int npad = (N + 31) / 32 * 32;
    if (j > npad) j = npad;
    p.J = (int)j;
    p.smem = ((size_t)C * p.J + p.J + (C == 3 ? 0 : (size_t)kKnnWarps * C)) * sizeof(float) + (size_t)kKnnRows * 32 * p.KR * (sizeof(float) + sizeof(int));
    p.ws_bytes = align_up((size_t)B * N * sizeof(float), 256);
    return p;
}

size_t knn_workspace(int B, int C, int N, int k) {
    if (B <= 0 || C < 1 || N < 1 || k < 1) return 0;
    return make_knn_plan(B, C, N, k).ws_bytes;
}

template <int KR, bool C3>
static int knn_launch_t(const float* x, const float* xx, int B, int C, int N, int k, const KnnPlan& p, int64_t* idx, cudaStream_t stream) {
    auto kern = knn_kernel<KR, C3>;
    if (p.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const dim3 grid((unsigned)cdiv(N, kKnnRows), (unsigned)B);
    kern<<<grid, kKnnThreads, p.smem, stream>>>(x, xx, C, N, k, p.J, idx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

int knn_forward_launch(const float* x, int B, int C, int N, int k, int64_t* idx, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B < 0 || C < 1 || N < 1) return fail(FPSG_EINVAL, "knn: need B >= 0, C >= 1, N >= 1 (got B=%d C=%d N=%d)", B, C, N);
    if (k < 1 || k > N) return fail(FPSG_EINVAL, "knn: k=%d out of range for N=%d", k, N);
    if (k > FPSG_KNN_MAX_K) return fail(FPSG_EINVAL, "knn: k=%d exceeds FPSG_KNN_MAX_K=%d", k, FPSG_KNN_MAX_K);
    if (B == 0) return FPSG_OK;
    if (!x || !idx) return fail(FPSG_EINVAL, "knn: null pointer");
    if (B > 65535) return fail(FPSG_EINVAL, "knn: B > 65535 not supported in one call");
    if ((size_t)C * 32 * sizeof(float) > 160 * 1024) return fail(FPSG_EINVAL, "knn: C=%d too large", C);
    const KnnPlan p = make_knn_plan(B, C, N, k);
    if (!ws || ws_bytes < p.ws_bytes) return fail(FPSG_EWORKSPACE, "knn: workspace of %zu bytes required, %zu given", p.ws_bytes, ws ? ws_bytes : (size_t)0);
    float* xx = static_cast<float*>(ws);
    knn_sqnorm_kernel<<<dim3((unsigned)cdiv(N, 256), (unsigned)B), 256, 0, stream>>>(x, C, N, xx);
    FPSG_CUDA_OK(cudaGetLastError());
    const bool c3 = (C == 3);
    switch (p.KR) {
        case 1:
            return c3 ? knn_launch_t<1, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_t<1, false>(x, xx, B, C, N, k, p, idx, stream);
        case 2:
            return c3 ? knn_launch_t<2, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_t<2, false>(x, xx, B, C, N, k, p, idx, stream);
        default:
            return c3 ? knn_launch_t<4, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_t<4, false>(x, xx, B, C, N, k, p, idx, stream);
    }
}

}  // namespace fpsg
