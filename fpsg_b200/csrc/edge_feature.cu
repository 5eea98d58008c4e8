// K4: EdgeConv graph feature for the DGCNN encoder, sm_100a.
//
// Replaces the tail of `get_graph_feature` (reference src/dgcnn/model.py:31-41): after `idx = knn(x, k)` the
// reference adds a per-cloud base to the indices, flat-gathers the neighbours from a transposed copy of x, tiles x
// with `repeat`, concatenates and `permute(...).contiguous()`s -- five passes over tensors the size of the output.
// Here the (B, 2C, N, k) tensor is written once, straight from x and idx:
//     out[b,     c, n, j] = x[b, c, idx[b,n,j]] - x[b, c, n]
//     out[b, C + c, n, j] = x[b, c, n]
// and the backward gathers through an inverted index of idx (built once per call with integer atomics, shared by all
// channels) instead of scattering with floating-point atomics:
//     gx[b,c,n] = sum_j (g[b,C+c,n,j] - g[b,c,n,j]) + sum_{(n',j): idx[b,n',j] = n} g[b,c,n',j]      (double accumulation)
// The forward is HBM-write-bound (8*C*N*k bytes per cloud); the backward reads as much but is bound by the random
// gathers (one L1 wavefront per gathered element) -- shared-memory staging of the plane and shared-memory double
// atomics were both tried in round 1 and were slower (uncoalesced list reads; 64-bit CAS loops).
#include "common.cuh"

namespace fpsg {

constexpr int kEdgeThreads = 256;

// one thread per group of 4 consecutive (n, j) slots of one cloud, looping over channels; float4 stores
template <bool VEC4>
__global__ void __launch_bounds__(kEdgeThreads)
edge_feature_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int C, int N, int k, float* __restrict__ out) {
    constexpr int W = VEC4 ? 4 : 1;
    const int b = blockIdx.y;
    const long long nk = (long long)N * k;
    const long long e0 = ((long long)blockIdx.x * kEdgeThreads + threadIdx.x) * W;
    if (e0 >= nk) return;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const int64_t* __restrict__ ib = idx + (size_t)b * nk;
    float* __restrict__ ob = out + (size_t)b * 2 * C * nk;

    int nb[W], ctr_n[W];
#pragma unroll
    for (int t = 0; t < W; ++t) {
        nb[t] = (int)ib[e0 + t];
        ctr_n[t] = (int)((e0 + t) / k);
    }
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
        const float* __restrict__ row = xb + (size_t)c * N;
        float d[W], m[W];
#pragma unroll
        for (int t = 0; t < W; ++t) {
            m[t] = __ldg(row + ctr_n[t]);
            d[t] = __fsub_rn(__ldg(row + nb[t]), m[t]);
        }
        float* od = ob + (size_t)c * nk + e0;
        float* om = ob + (size_t)(C + c) * nk + e0;
        if (VEC4) {
            __stcs(reinterpret_cast<float4*>(od), make_float4(d[0], d[1], d[2], d[3]));
            __stcs(reinterpret_cast<float4*>(om), make_float4(m[0], m[1], m[2], m[3]));
        } else {
            od[0] = d[0];
            om[0] = m[0];
        }
    }
}

// inverted index of idx for one cloud: for every point n the list of slots e = n'*k + j with idx[e] == n.
// start (N+1 ints) and list (N*k ints) live in the workspace. One CTA per cloud; counts in shared memory.
__global__ void __launch_bounds__(1024)
edge_invert_kernel(const int64_t* __restrict__ idx, int N, int k, int* __restrict__ start_all, int* __restrict__ list_all) {
    extern __shared__ int sCnt[];  // N + 1 counters, then reused as cursors
    __shared__ int sWarp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long nk = (long long)N * k;
    const int64_t* __restrict__ ib = idx + (size_t)b * nk;
    int* __restrict__ start = start_all + (size_t)b * (N + 1);
    int* __restrict__ list = list_all + (size_t)b * nk;

    for (int n = tid; n <= N; n += 1024) sCnt[n] = 0;
    __syncthreads();
#pragma unroll 8
    for (long long e = tid; e < nk; e += 1024) atomicAdd(&sCnt[(int)ib[e]], 1);
    __syncthreads();
    // exclusive scan over N counters, 1024 threads, chunked
    int carry = 0;
    for (int base = 0; base < N; base += 1024) {
        const int n = base + tid;
        const int c = n < N ? sCnt[n] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) sWarp[w] = incl;
        __syncthreads();
        if (w == 0) {
            int v = sWarp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            sWarp[lane] = v;  // inclusive over warps
        }
        __syncthreads();
        const int excl = carry + (w > 0 ? sWarp[w - 1] : 0) + incl - c;
        const int total = sWarp[31];
        __syncthreads();
        if (n < N) {
            start[n] = excl;
            sCnt[n] = excl;  // cursor
        }
        carry += total;
    }
    if (tid == 0) start[N] = carry;
    __syncthreads();
#pragma unroll 8
    for (long long e = tid; e < nk; e += 1024) list[atomicAdd(&sCnt[(int)ib[e]], 1)] = (int)e;
}

// one thread per (cloud, channel, point): direct term over its own k slots + gathered term over its bucket
__global__ void __launch_bounds__(kEdgeThreads)
edge_feature_bwd_kernel(const float* __restrict__ g, const int* __restrict__ start_all, const int* __restrict__ list_all, int C, int N,
                        int k, float* __restrict__ gx) {
    const int n = blockIdx.x * kEdgeThreads + threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    if (n >= N) return;
    const long long nk = (long long)N * k;
    const float* __restrict__ gd = g + ((size_t)b * 2 * C + c) * nk;      // gradient of the difference half
    const float* __restrict__ gm = g + ((size_t)b * 2 * C + C + c) * nk;  // gradient of the centre half
    const int* __restrict__ start = start_all + (size_t)b * (N + 1);
    const int* __restrict__ list = list_all + (size_t)b * nk;
    double acc = 0.0;
    const long long e0 = (long long)n * k;
    if ((k & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        // the point's k slots as 16-byte pieces: a quarter of the load instructions and L1 wavefronts
        const float4* gm4 = reinterpret_cast<const float4*>(gm + e0);
        const float4* gd4 = reinterpret_cast<const float4*>(gd + e0);
        for (int j = 0; j < (k >> 2); ++j) {
            const float4 m = __ldcs(gm4 + j), d = __ldg(gd4 + j);
            acc += ((double)m.x - (double)d.x) + ((double)m.y - (double)d.y) + ((double)m.z - (double)d.z) + ((double)m.w - (double)d.w);
        }
    } else {
        for (int j = 0; j < k; ++j) acc += (double)__ldcs(gm + e0 + j) - (double)__ldg(gd + e0 + j);
    }
    const int s0 = start[n], s1 = start[n + 1];
    for (int s = s0; s < s1; s += 4) {
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = s + t < s1 ? __ldg(gd + list[s + t]) : 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) acc += (double)v[t];
    }
    gx[((size_t)b * C + c) * N + n] = (float)acc;
}

int edge_feature_forward_launch(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out, cudaStream_t stream) {
    if (B < 0 || C < 1 || N < 1 || k < 1) return fail(FPSG_EINVAL, "edge_feature: need B >= 0, C, N, k >= 1 (got B=%d C=%d N=%d k=%d)", B, C, N, k);
    if (B == 0) return FPSG_OK;
    if (!x || !idx || !out) return fail(FPSG_EINVAL, "edge_feature: null pointer");
    if (B > 65535) return fail(FPSG_EINVAL, "edge_feature: B > 65535 not supported in one call");
    const long long nk = (long long)N * k;
    if (nk > 0x7fffffffLL) return fail(FPSG_EINVAL, "edge_feature: N*k too large");
    const bool vec4 = (nk % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec4) {
        const dim3 grid((unsigned)((nk / 4 + kEdgeThreads - 1) / kEdgeThreads), (unsigned)B);
        edge_feature_fwd_kernel<true><<<grid, kEdgeThreads, 0, stream>>>(x, idx, C, N, k, out);
    } else {
        const dim3 grid((unsigned)((nk + kEdgeThreads - 1) / kEdgeThreads), (unsigned)B);
        edge_feature_fwd_kernel<false><<<grid, kEdgeThreads, 0, stream>>>(x, idx, C, N, k, out);
    }
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

size_t edge_feature_backward_workspace(int B, int C, int N, int k) {
    (void)C;
    if (B <= 0 || N < 1 || k < 1) return 0;
    return align_up((size_t)B * (N + 1) * sizeof(int), 256) + align_up((size_t)B * N * k * sizeof(int), 256);
}

int edge_feature_backward_launch(const float* g, const int64_t* idx, int B, int C, int N, int k, float* gx, void* ws, size_t ws_bytes,
                                 cudaStream_t stream) {
    if (B < 0 || C < 1 || N < 1 || k < 1) return fail(FPSG_EINVAL, "edge_feature_bwd: need B >= 0, C, N, k >= 1");
    if (B == 0) return FPSG_OK;
    if (!g || !idx || !gx) return fail(FPSG_EINVAL, "edge_feature_bwd: null pointer");
    if (B > 65535 || C > 65535) return fail(FPSG_EINVAL, "edge_feature_bwd: B, C > 65535 not supported in one call");
    if ((long long)N * k > 0x7fffffffLL) return fail(FPSG_EINVAL, "edge_feature_bwd: N*k too large");
    if ((size_t)(N + 1) * sizeof(int) > 200 * 1024) return fail(FPSG_EINVAL, "edge_feature_bwd: N=%d too large (shared-memory counters)", N);
    const size_t need = edge_feature_backward_workspace(B, C, N, k);
    if (!ws || ws_bytes < need) return fail(FPSG_EWORKSPACE, "edge_feature_bwd: workspace of %zu bytes required, %zu given", need, ws ? ws_bytes : (size_t)0);
    int* start = static_cast<int*>(ws);
    int* list = reinterpret_cast<int*>(static_cast<char*>(ws) + align_up((size_t)B * (N + 1) * sizeof(int), 256));
    const size_t smem = (size_t)(N + 1) * sizeof(int);
    if (smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(edge_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_invert_kernel<<<B, 1024, smem, stream>>>(idx, N, k, start, list);
    FPSG_CUDA_OK(cudaGetLastError());
    const dim3 grid((unsigned)cdiv(N, kEdgeThreads), (unsigned)C, (unsigned)B);
    edge_feature_bwd_kernel<<<grid, kEdgeThreads, 0, stream>>>(g, start, list, C, N, k, gx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

}  // namespace fpsg
