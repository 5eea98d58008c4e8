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
// The forward is HBM-write-bound (8*C*N*k bytes per cloud). The backward reads as much. Gathering straight from
// global memory costs one 32-byte sector per 4-byte element, so when a (cloud, channel) plane of the difference half
// (N*k floats) fits shared memory twice, a persistent grid (one 1024-thread CTA per SM) streams the planes through a
// double buffer with 16-byte cp.async -- the next plane arrives while the current one is gathered -- with the cloud's
// inverted index held in shared memory as 16-bit slot numbers. Larger planes use the direct global gather.
#include <cstdlib>

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

// The same forward with the channel's row of x staged in shared memory (two rows in flight, 16-byte cp.async): gathering a row
// straight from global memory costs an L1 wavefront per 128-byte line a warp touches (about 20 per gather instruction for a random
// kNN graph) -- 3.9 TB/s written, bound by the load/store unit, not by HBM; from shared memory a gather is 3-4 wavefronts.
// Needs k % 4 == 0 (the four slots of a thread then belong to one centre point), N % 4 == 0 and 16-byte aligned x.
__global__ void __launch_bounds__(kEdgeThreads)
edge_feature_fwd_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int C, int N, int k, float* __restrict__ out) {
    extern __shared__ __align__(16) float sRows[];  // [2][N]
    const int b = blockIdx.y, tid = threadIdx.x;
    const long long nk = (long long)N * k;
    const long long e0 = ((long long)blockIdx.x * kEdgeThreads + tid) * 4;
    const bool live = e0 < nk;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    float* __restrict__ ob = out + (size_t)b * 2 * C * nk;
    int nb[4] = {0, 0, 0, 0};
    int ctr = 0;
    if (live) {
        const int64_t* __restrict__ ib = idx + (size_t)b * nk + e0;
#pragma unroll
        for (int t = 0; t < 4; ++t) nb[t] = (int)ib[t];
        ctr = (int)(e0 / k);
    }
    auto prefetch = [&](int c) {
        const float4* src = reinterpret_cast<const float4*>(xb + (size_t)c * N);
        float4* dst = reinterpret_cast<float4*>(sRows + (size_t)(c & 1) * N);
        for (int q = tid; q < (N >> 2); q += kEdgeThreads) {
            const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst + q));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0);
    for (int c = 0; c < C; ++c) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // row c is complete for everyone, and everyone is done with row c - 1 (whose buffer the next prefetch overwrites)
        if (c + 1 < C) prefetch(c + 1);
        if (live) {
            const float* __restrict__ row = sRows + (size_t)(c & 1) * N;
            const float m = row[ctr];
            const float4 d = make_float4(__fsub_rn(row[nb[0]], m), __fsub_rn(row[nb[1]], m), __fsub_rn(row[nb[2]], m), __fsub_rn(row[nb[3]], m));
            __stcs(reinterpret_cast<float4*>(ob + (size_t)c * nk + e0), d);
            __stcs(reinterpret_cast<float4*>(ob + (size_t)(C + c) * nk + e0), make_float4(m, m, m, m));
        }
    }
}

// inverted index of idx for one cloud: for every point n the list of slots e = n'*k + j with idx[e] == n.
// start (N+1 ints) and list (N*k ints) live in the workspace. One CTA per cloud; counts in shared memory.
// For the plane-staged backward the buckets are also cut into SUB-BUCKETS of at most kPlaneSub entries (kNN graphs of
// high-dimensional features have hubs: a fifth of the points of a 64-channel cloud have more than 32 in-edges, some several
// hundred): vfirst[n] = number of sub-buckets of the points before n (N+1 entries, 16-bit), sub[v] = first list position of
// sub-bucket v | its length << 16. On that path every bucket starts on a multiple of 4 list entries -- 8 bytes of the 16-bit copy the
// kernel keeps in shared memory, so four entries arrive per load -- which needs N*k + 3N <= 65535 (16-bit positions; the padding
// entries are zeroed by the caller). list_stride = list entries per cloud.
constexpr int kPlaneSub = 16;

__global__ void __launch_bounds__(1024)
edge_invert_kernel(const int64_t* __restrict__ idx, int N, int k, int* __restrict__ start_all, int* __restrict__ list_all,
                   long long list_stride, unsigned short* __restrict__ vfirst_all, unsigned* __restrict__ sub_all, int vstride) {
    extern __shared__ int sCnt[];  // N + 1 counters, then reused as cursors
    __shared__ int sWarp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long nk = (long long)N * k;
    const int64_t* __restrict__ ib = idx + (size_t)b * nk;
    int* __restrict__ start = start_all + (size_t)b * (N + 1);
    int* __restrict__ list = list_all + (size_t)b * list_stride;
    const bool padded = vfirst_all != nullptr;

    for (int n = tid; n <= N; n += 1024) sCnt[n] = 0;
    __syncthreads();
#pragma unroll 8
    for (long long e = tid; e < nk; e += 1024) atomicAdd(&sCnt[(int)ib[e]], 1);
    __syncthreads();
    // exclusive scan over N values, 1024 threads, chunked; value(n) is given by f, the result handed to out(n, exclusive prefix)
    auto scan = [&](auto f, auto out) {
        int carry = 0;
        for (int base = 0; base < N; base += 1024) {
            const int n = base + tid;
            const int c = n < N ? f(n) : 0;
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
            if (n < N) out(n, excl, c);
            carry += total;
        }
        return carry;
    };
    if (padded) {  // sub-bucket counts (before the counters turn into cursors)
        unsigned short* __restrict__ vfirst = vfirst_all + (size_t)b * (N + 1);
        const int V = scan([&](int n) { return (sCnt[n] + kPlaneSub - 1) / kPlaneSub; }, [&](int n, int excl, int) { vfirst[n] = (unsigned short)excl; });
        if (tid == 0) vfirst[N] = (unsigned short)V;
    }
    const int total = scan([&](int n) { return padded ? (sCnt[n] + 3) & ~3 : sCnt[n]; },
                           [&](int n, int excl, int) {
                               start[n] = excl;
                               if (padded) {
                                   const int c = sCnt[n];
                                   unsigned* __restrict__ sub = sub_all + (size_t)b * vstride + vfirst_all[(size_t)b * (N + 1) + n];
                                   for (int j = 0; j * kPlaneSub < c; ++j)
                                       sub[j] = (unsigned)(excl + j * kPlaneSub) | ((unsigned)min(kPlaneSub, c - j * kPlaneSub) << 16);
                               }
                               sCnt[n] = excl;  // cursor
                           });
    if (tid == 0) start[N] = total;
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

// Plane-staged backward: CTA g owns the planes [g*P/G, (g+1)*P/G) of the flattened (cloud, channel) range.
// Per plane: (A) the sub-buckets of the cloud's inverted index are dealt round-robin to the 1024 threads -- at most kPlaneSub
// gathers each, whatever the in-degree of their point, so a hub costs its point's thread nothing extra -- and their sums go to
// shared memory as doubles; (B) every point's thread adds its own k slots, the centre half and its sub-bucket sums in index
// order (a fixed order: the result does not depend on the schedule).
constexpr int kPlaneThreads = 1024;
constexpr int kPlaneK4 = 5;  // k <= 20 on this path (DGCNN uses 20)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

static int edge_plane_vmax(int N, int k) { return (int)(((long long)N * k) / kPlaneSub) + N; }  // sub-buckets of a cloud, at most
static int edge_plane_list(int N, int k) { return N * k + 3 * N; }                                // list entries of a cloud with 4-entry aligned buckets, at most
constexpr int kPlaneSubPerThread = 3;  // sub-buckets a thread takes (v = tid, tid + 1024, tid + 2048): vmax <= 3072

__global__ void __launch_bounds__(kPlaneThreads, 1)
edge_feature_bwd_plane_kernel(const float* __restrict__ g, const int* __restrict__ list_all, const unsigned short* __restrict__ vfirst_all,
                              const unsigned* __restrict__ sub_all, int vmax, int nlist, int B, int C, int N, int k, float* __restrict__ gx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nk = N * k;
    float* sPlane = reinterpret_cast<float*>(smem_raw);                            // [2][nk] difference-half planes
    double* sPart = reinterpret_cast<double*>(sPlane + 2 * (size_t)nk);             // [vmax] sub-bucket sums of the current plane
    unsigned short* sList = reinterpret_cast<unsigned short*>(sPart + vmax);       // [nlist] slots grouped by target point, buckets 8-byte aligned
    const int tid = threadIdx.x;
    const long long planes = (long long)B * C;
    const int p0 = (int)(planes * blockIdx.x / gridDim.x), p1 = (int)(planes * (blockIdx.x + 1) / gridDim.x);

    auto prefetch = [&](int p, int buf) {
        const int b = p / C, c = p - b * C;
        const float4* src = reinterpret_cast<const float4*>(g + ((size_t)b * 2 * C + c) * nk);
        float4* dst = reinterpret_cast<float4*>(sPlane + (size_t)buf * nk);
        for (int q = tid; q < (nk >> 2); q += kPlaneThreads) cp_async16(dst + q, src + q);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // The centre half is read exactly once and only its per-point sum is needed: a thread loads the k values of its
    // point (N <= 1024: one point per thread) for the NEXT plane at the top of an iteration and folds them into one
    // double at the end of it, so their HBM latency hides behind the current plane's gathers.
    const int k4 = k >> 2;
    float4 mreg[kPlaneK4];
    auto load_centre = [&](int p) {
        const int b = p / C, c = p - b * C;
        const float4* gm4 = reinterpret_cast<const float4*>(g + ((size_t)b * 2 * C + C + c) * nk + (size_t)min(tid, N - 1) * k);
#pragma unroll
        for (int j = 0; j < kPlaneK4; ++j)
            if (j < k4) mreg[j] = __ldcs(gm4 + j);
    };
    auto fold_centre = [&]() {
        double s_ = 0.0;
#pragma unroll
        for (int j = 0; j < kPlaneK4; ++j)
            if (j < k4) s_ += ((double)mreg[j].x + (double)mreg[j].y) + ((double)mreg[j].z + (double)mreg[j].w);
        return s_;
    };
    double msum = 0.0;

    // per cloud, in registers: this thread's sub-buckets (first list position | length << 16; 0 = none) and its point's range of sums
    unsigned mysub[kPlaneSubPerThread];
#pragma unroll
    for (int r = 0; r < kPlaneSubPerThread; ++r) mysub[r] = 0u;
    int v0 = 0, v1 = 0;
    int cur_b = -1;
    if (p0 < p1) {
        prefetch(p0, 0);
        load_centre(p0);
        msum = fold_centre();
    }
    for (int p = p0; p < p1; ++p) {
        const int b = p / C, c = p - b * C;
        const int buf = (p - p0) & 1;
        if (p + 1 < p1) prefetch(p + 1, buf ^ 1);
        if (b != cur_b) {  // CTA-uniform: the index of a new cloud (the previous plane's readers passed the barrier at the loop's end)
            cur_b = b;
            const int* __restrict__ list = list_all + (size_t)b * nlist;
            const unsigned short* __restrict__ vfirst = vfirst_all + (size_t)b * (N + 1);
            const unsigned* __restrict__ sub = sub_all + (size_t)b * vmax;
            const int V = __ldg(vfirst + N);
            for (int e = tid; e < nlist; e += kPlaneThreads) sList[e] = (unsigned short)__ldg(list + e);
#pragma unroll
            for (int r = 0; r < kPlaneSubPerThread; ++r) {
                const int v = tid + r * kPlaneThreads;
                mysub[r] = v < V ? __ldg(sub + v) : 0u;
            }
            v0 = __ldg(vfirst + min(tid, N - 1));
            v1 = tid < N ? (int)__ldg(vfirst + tid + 1) : v0;
        }
        if (p + 1 < p1)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // plane p and the index are in shared memory for everyone

        const float* __restrict__ plane = sPlane + (size_t)buf * nk;
        if (p + 1 < p1) load_centre(p + 1);  // consumed by fold_centre() below, after this plane's gathers
        // (A) sub-bucket sums. Fully unrolled and predicated: the (up to) 16 list entries are loaded back to back, then the 16
        // gathers, then one fixed summation tree -- as a loop of dependent round trips this phase ran at 30 % issue utilisation.
#pragma unroll
        for (int r = 0; r < kPlaneSubPerThread; ++r) {
            const int len = (int)(mysub[r] >> 16);
            if (len != 0) {
                const uint2* lp = reinterpret_cast<const uint2*>(sList + (mysub[r] & 0xffffu));  // 8-byte aligned: four entries per load
                unsigned e[kPlaneSub];
#pragma unroll
                for (int q = 0; q < kPlaneSub / 4; ++q) {
                    const uint2 w4 = 4 * q < len ? lp[q] : make_uint2(0u, 0u);
                    e[4 * q] = w4.x & 0xffffu, e[4 * q + 1] = w4.x >> 16, e[4 * q + 2] = w4.y & 0xffffu, e[4 * q + 3] = w4.y >> 16;
                }
                float val[kPlaneSub];
#pragma unroll
                for (int t = 0; t < kPlaneSub; ++t) val[t] = t < len ? plane[e[t]] : 0.f;
                double d4[kPlaneSub / 4];
#pragma unroll
                for (int t = 0; t < kPlaneSub / 4; ++t)
                    d4[t] = ((double)val[4 * t] + (double)val[4 * t + 1]) + ((double)val[4 * t + 2] + (double)val[4 * t + 3]);
                sPart[tid + r * kPlaneThreads] = (d4[0] + d4[1]) + (d4[2] + d4[3]);
            }
        }
        // the point's own k slots of the difference half (while the others finish their sub-buckets)
        double acc = msum;
        if (tid < N) {
            const float4* gd4 = reinterpret_cast<const float4*>(plane + tid * k);
#pragma unroll
            for (int j = 0; j < kPlaneK4; ++j) {
                if (j < k4) {
                    const float4 d = gd4[j];
                    acc -= ((double)d.x + (double)d.y) + ((double)d.z + (double)d.w);
                }
            }
        }
        __syncthreads();
        // (B) the point's sub-bucket sums, in index order
        if (tid < N) {
            for (int v = v0; v < v1; ++v) acc += sPart[v];
            gx[((size_t)b * C + c) * N + tid] = (float)acc;
        }
        if (p + 1 < p1) msum = fold_centre();
        __syncthreads();  // everyone is done with this buffer and with sPart before the next plane overwrites them
    }
}

static size_t edge_plane_smem(int N, int k) {
    const size_t nk = (size_t)N * k, vmax = (size_t)edge_plane_vmax(N, k);
    return 2 * nk * sizeof(float) + vmax * sizeof(double) + (size_t)edge_plane_list(N, k) * sizeof(unsigned short);
}

int edge_feature_forward_launch(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out, cudaStream_t stream) {
    if (B < 0 || C < 1 || N < 1 || k < 1) return fail(FPSG_EINVAL, "edge_feature: need B >= 0, C, N, k >= 1 (got B=%d C=%d N=%d k=%d)", B, C, N, k);
    if (B == 0) return FPSG_OK;
    if (!x || !idx || !out) return fail(FPSG_EINVAL, "edge_feature: null pointer");
    if (B > 65535) return fail(FPSG_EINVAL, "edge_feature: B > 65535 not supported in one call");
    const long long nk = (long long)N * k;
    if (nk > 0x7fffffffLL) return fail(FPSG_EINVAL, "edge_feature: N*k too large");
    const bool vec4 = (nk % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const size_t rows_smem = 2 * (size_t)N * sizeof(float);
    if (vec4 && (k & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && rows_smem <= 96 * 1024 && C >= 8 && !tune_env("FPSG_EDGE_NOROWS")) {
        const dim3 grid((unsigned)((nk / 4 + kEdgeThreads - 1) / kEdgeThreads), (unsigned)B);
        if (rows_smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(edge_feature_fwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
        edge_feature_fwd_rows_kernel<<<grid, kEdgeThreads, rows_smem, stream>>>(x, idx, C, N, k, out);
    } else if (vec4) {
        const dim3 grid((unsigned)((nk / 4 + kEdgeThreads - 1) / kEdgeThreads), (unsigned)B);
        edge_feature_fwd_kernel<true><<<grid, kEdgeThreads, 0, stream>>>(x, idx, C, N, k, out);
    } else {
        const dim3 grid((unsigned)((nk + kEdgeThreads - 1) / kEdgeThreads), (unsigned)B);
        edge_feature_fwd_kernel<false><<<grid, kEdgeThreads, 0, stream>>>(x, idx, C, N, k, out);
    }
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

static bool edge_plane_path(int N, int k) {
    const long long nk = (long long)N * k;
    return nk + 3LL * N <= 65535 && (k & 3) == 0 && k <= 4 * kPlaneK4 && N <= kPlaneThreads && edge_plane_vmax(N, k) <= kPlaneSubPerThread * kPlaneThreads &&
           edge_plane_smem(N, k) <= 226 * 1024;
}

size_t edge_feature_backward_workspace(int B, int C, int N, int k) {
    (void)C;
    if (B <= 0 || N < 1 || k < 1) return 0;
    // (the list is sized for the padded layout of the plane path whether or not the call takes it: that also depends on the alignment of g)
    size_t need = align_up((size_t)B * (N + 1) * sizeof(int), 256) + align_up((size_t)B * ((size_t)N * k + 3 * (size_t)N) * sizeof(int), 256);
    if (edge_plane_path(N, k))  // the sub-bucket tables
        need += align_up((size_t)B * (N + 1) * sizeof(unsigned short), 256) + align_up((size_t)B * edge_plane_vmax(N, k) * sizeof(unsigned), 256);
    return need;
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
    char* wp = static_cast<char*>(ws);
    int* start = reinterpret_cast<int*>(wp);
    wp += align_up((size_t)B * (N + 1) * sizeof(int), 256);
    int* list = reinterpret_cast<int*>(wp);
    wp += align_up((size_t)B * ((size_t)N * k + 3 * (size_t)N) * sizeof(int), 256);
    const bool plane = edge_plane_path(N, k) && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && !tune_env("FPSG_EDGE_NOPLANE");
    const int vmax = edge_plane_vmax(N, k), nlist = edge_plane_list(N, k);
    unsigned short* vfirst = nullptr;
    unsigned* vbeg = nullptr;
    if (plane) {
        vfirst = reinterpret_cast<unsigned short*>(wp);
        wp += align_up((size_t)B * (N + 1) * sizeof(unsigned short), 256);
        vbeg = reinterpret_cast<unsigned*>(wp);
        FPSG_CUDA_OK(cudaMemsetAsync(list, 0, (size_t)B * nlist * sizeof(int), stream));  // the padding entries of the list: slot 0
    }
    const size_t smem = (size_t)(N + 1) * sizeof(int);
    if (smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(edge_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_invert_kernel<<<B, 1024, smem, stream>>>(idx, N, k, start, list, plane ? (long long)nlist : (long long)N * k, vfirst, vbeg, vmax);
    FPSG_CUDA_OK(cudaGetLastError());
    if (plane) {
        const size_t psmem = edge_plane_smem(N, k);
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long planes = (long long)B * C;
        const unsigned grid = (unsigned)(planes < sms ? planes : sms);
        FPSG_CUDA_OK(cudaFuncSetAttribute(edge_feature_bwd_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        edge_feature_bwd_plane_kernel<<<grid, kPlaneThreads, psmem, stream>>>(g, list, vfirst, vbeg, vmax, nlist, B, C, N, k, gx);
        FPSG_CUDA_OK(cudaGetLastError());
        return FPSG_OK;
    }
    const dim3 grid((unsigned)cdiv(N, kEdgeThreads), (unsigned)C, (unsigned)B);
    edge_feature_bwd_kernel<<<grid, kEdgeThreads, 0, stream>>>(g, start, list, C, N, k, gx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

}  // namespace fpsg
