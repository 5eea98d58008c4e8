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
#include "knn_common.cuh"

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

// Asynchronous global->shared copy of one float (LDGSTS): a whole tile is issued without waiting for any element,
// so a tile load costs one memory latency instead of one per row of the tile.
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void cp_async_f32x4(float* smem_dst, const float* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// loads the [C][jn] block x[:, j0:j0+jn] and the matching squared norms into shared memory (row stride J)
__device__ __forceinline__ void knn_issue_tile(float* sTile, float* sXX, const float* __restrict__ xb, const float* __restrict__ xxb,
                                               int C, int N, int J, int j0, int jn, int tid, int nthreads = kKnnThreads) {
    const int lane = tid & 31, w = tid >> 5, nwarps = nthreads >> 5;
    // 16-byte copies when every row segment is 16-byte aligned on both sides (J is always a multiple of 32)
    const bool vec = ((N | j0 | jn) & 3) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0 && (reinterpret_cast<uintptr_t>(xxb) & 15) == 0;
    if (vec) {
        const int jq = jn >> 2;
        for (int c = w; c < C; c += nwarps) {
            const float* src = xb + (size_t)c * N + j0;
            float* dst = sTile + (size_t)c * J;
            for (int jj = lane; jj < jq; jj += 32) cp_async_f32x4(dst + 4 * jj, src + 4 * jj);
        }
        for (int jj = tid; jj < jq; jj += nthreads) cp_async_f32x4(sXX + 4 * jj, xxb + j0 + 4 * jj);
    } else {
        for (int c = w; c < C; c += nwarps) {  // a warp copies whole channel rows: coalesced, no index arithmetic
            const float* src = xb + (size_t)c * N + j0;
            float* dst = sTile + (size_t)c * J;
            for (int jj = lane; jj < jn; jj += 32) cp_async_f32(dst + jj, src + jj);
        }
        for (int jj = tid; jj < jn; jj += nthreads) cp_async_f32(sXX + jj, xxb + j0 + jj);
    }
}
__device__ __forceinline__ void knn_load_tile(float* sTile, float* sXX, const float* __restrict__ xb, const float* __restrict__ xxb,
                                              int C, int N, int J, int j0, int jn, int tid, int nthreads = kKnnThreads) {
    knn_issue_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid, nthreads);
    cp_async_wait_all();
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
        knn_load_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid);
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

// Generic channel count: every warp owns QB query points for the whole kernel (their top-k lists never leave
// registers) and scores 32 candidates per step against all QB of them: per channel one candidate value per lane
// (LDS), the QB query values by broadcast (LDS.128) and QB FFMAs -- a register-blocked FP32 inner product. The
// candidate tiles ([C][J] floats) are shared by the CTA's warps.
constexpr int kKnnQB = 8;

template <int KR>
__global__ void __launch_bounds__(kKnnThreads)
knn_blocked_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int k, int J, int64_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sTile = reinterpret_cast<float*>(smem_raw);  // [C][J]
    float* sXX = sTile + (size_t)C * J;                 // [J]
    float* sQ = sXX + J;                                // [kKnnWarps][C][kKnnQB] query values, channel-major

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y;
    const int i0 = (blockIdx.x * kKnnWarps + w) * kKnnQB;  // first query point of this warp
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    const int kth_lane = (k - 1) & 31, kth_reg = (k - 1) >> 5;
    float* myQ = sQ + (size_t)w * C * kKnnQB;

    // query values and norms (rows past N are clamped: they compute garbage that is never written)
    for (int e = lane; e < C * kKnnQB; e += 32) {
        const int c = e / kKnnQB, q = e - c * kKnnQB;
        myQ[e] = __ldg(xb + (size_t)c * N + min(i0 + q, N - 1));
    }
    float xxi[kKnnQB], tau[kKnnQB];
    float lv[kKnnQB][KR];
    int li[kKnnQB][KR];
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        xxi[q] = __ldg(xxb + min(i0 + q, N - 1));
        tau[q] = -CUDART_INF_F;
#pragma unroll
        for (int r = 0; r < KR; ++r) {
            lv[q][r] = -CUDART_INF_F;
            li[q][r] = 0;
        }
    }
    __syncwarp();

    for (int j0 = 0; j0 < N; j0 += J) {
        const int jn = min(J, N - j0);
        __syncthreads();  // previous tile fully consumed
        knn_load_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid);
        __syncthreads();
        if (i0 >= N) continue;  // warp-uniform; the warp still takes part in the barriers

        for (int jb = 0; jb < jn; jb += 32) {
            const int jj = min(jb + lane, jn - 1);  // clamped lanes are masked out below
            float dot[kKnnQB];
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q) dot[q] = 0.f;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const float cv = sTile[(size_t)c * J + jj];
                const float4 qa = *reinterpret_cast<const float4*>(myQ + c * kKnnQB);
                const float4 qb = *reinterpret_cast<const float4*>(myQ + c * kKnnQB + 4);
                dot[0] = __fmaf_rn(qa.x, cv, dot[0]);
                dot[1] = __fmaf_rn(qa.y, cv, dot[1]);
                dot[2] = __fmaf_rn(qa.z, cv, dot[2]);
                dot[3] = __fmaf_rn(qa.w, cv, dot[3]);
                dot[4] = __fmaf_rn(qb.x, cv, dot[4]);
                dot[5] = __fmaf_rn(qb.y, cv, dot[5]);
                dot[6] = __fmaf_rn(qb.z, cv, dot[6]);
                dot[7] = __fmaf_rn(qb.w, cv, dot[7]);
            }
            const float nxx = -sXX[jj];
            const bool inb = jb + lane < jn;
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q) {
                const float pd = inb ? __fsub_rn(__fmaf_rn(2.0f, dot[q], nxx), xxi[q]) : -CUDART_INF_F;
                unsigned mask = __ballot_sync(0xffffffffu, pd > tau[q]);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float v = __shfl_sync(0xffffffffu, pd, src);
                    if (!(v > tau[q])) continue;  // tau rose since the ballot
                    const int vj = j0 + jb + src;
                    int pos = 0;
#pragma unroll
                    for (int r = 0; r < KR; ++r) pos += __popc(__ballot_sync(0xffffffffu, lv[q][r] >= v));
#pragma unroll
                    for (int r = KR - 1; r >= 0; --r) {
                        float pv = __shfl_up_sync(0xffffffffu, lv[q][r], 1);
                        int pi = __shfl_up_sync(0xffffffffu, li[q][r], 1);
                        if (r > 0) {
                            const float cv2 = __shfl_sync(0xffffffffu, lv[q][r - 1], 31);
                            const int ci = __shfl_sync(0xffffffffu, li[q][r - 1], 31);
                            if (lane == 0) {
                                pv = cv2;
                                pi = ci;
                            }
                        }
                        const int e = r * 32 + lane;
                        if (e > pos) {
                            lv[q][r] = pv;
                            li[q][r] = pi;
                        } else if (e == pos) {
                            lv[q][r] = v;
                            li[q][r] = vj;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < KR; ++r)
                        if (r == kth_reg) tau[q] = __shfl_sync(0xffffffffu, lv[q][r], kth_lane);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        if (i0 + q < N) {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int e = r * 32 + lane;
                if (e < k) idx[((size_t)b * N + i0 + q) * k + e] = li[q][r];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k <= 32: bulk top-k. The sorted list of a query is one 64-bit key per lane,
//   key = (order-preserving bits of pd) << 32 | (0xFFFFFFFF - j)      larger key = better, lower j wins ties,
// and candidates that beat the current k-th value are not inserted one by one but parked in a 64-slot
// shared-memory buffer (ballot + rank). Only when 32 have piled up -- about five times per query at N = 1024,
// because every flush raises the threshold -- are they sorted by a warp bitonic network and merged into the list
// (reverse, element-wise max, bitonic merge). Cost per buffered candidate ~5 instructions instead of ~35 for a
// shuffle insertion.
// ---------------------------------------------------------------------------------------------------------
// Sorts up to 64 buffered candidate keys and merges them into the sorted list `key` (one entry per lane). Kept out
// of line on purpose: eight inlined copies (one per query of the warp) overflow the instruction cache.
__device__ __noinline__ u64 knn_flush_list(u64 key, const u64* buf, int cnt) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        if (cnt > 32 * h) {  // warp-uniform
            u64 kb = 32 * h + lane < cnt ? buf[32 * h + lane] : 0ull;
            kb = warp_sort_desc(kb, lane);
            const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
            key = warp_merge_desc(key > rev ? key : rev, lane);
        }
    }
    __syncwarp();
    return key;
}

template <bool C3>
__global__ void __launch_bounds__(kKnnThreads)
knn_bulk_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int k, int J, int64_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sTile = reinterpret_cast<float*>(smem_raw);  // [C][J]
    float* sXX = sTile + (size_t)C * J;                 // [J]
    float* sQ = sXX + J;                                // [kKnnWarps][C][kKnnQB] query values (generic C only)
    u64* sBuf = reinterpret_cast<u64*>(sQ + (C3 ? 0 : (size_t)kKnnWarps * C * kKnnQB));  // [kKnnWarps][kKnnQB][64]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int b = blockIdx.y;
    const int i0 = (blockIdx.x * kKnnWarps + w) * kKnnQB;  // first query point of this warp
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    float* myQ = sQ + (size_t)w * C * kKnnQB;
    u64* myBuf = sBuf + (size_t)w * kKnnQB * 64;

    float q0[kKnnQB], q1[kKnnQB], q2[kKnnQB];  // C3: the query points themselves
    float xxi[kKnnQB], tau[kKnnQB];
    u64 key[kKnnQB];
    int cnt[kKnnQB];
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        const int i = min(i0 + q, N - 1);  // rows past N are clamped: computed, never written
        if (C3) {
            q0[q] = __ldg(xb + i);
            q1[q] = __ldg(xb + (size_t)N + i);
            q2[q] = __ldg(xb + 2 * (size_t)N + i);
        }
        xxi[q] = __ldg(xxb + i);
        tau[q] = -CUDART_INF_F;
        key[q] = 0ull;
        cnt[q] = 0;
    }
    if (!C3) {
        for (int e = lane; e < C * kKnnQB; e += 32) {
            const int c = e / kKnnQB, q = e - c * kKnnQB;
            myQ[e] = __ldg(xb + (size_t)c * N + min(i0 + q, N - 1));
        }
        __syncwarp();
    }

    // sorts the buffered candidates of query q and merges them into its list; refreshes the threshold
    auto flush = [&](int q) {
        key[q] = knn_flush_list(key[q], myBuf + q * 64, cnt[q]);
        cnt[q] = 0;
        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(key[q] >> 32), k - 1);
        tau[q] = hi ? ord_to_f32(hi) : -CUDART_INF_F;  // fewer than k entries so far: everything passes
    };

    for (int j0 = 0; j0 < N; j0 += J) {
        const int jn = min(J, N - j0);
        __syncthreads();  // previous tile fully consumed
        knn_load_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid);
        __syncthreads();
        if (i0 >= N) continue;  // warp-uniform; the warp still takes part in the barriers

        for (int jb = 0; jb < jn; jb += 32) {
            const int jj = min(jb + lane, jn - 1);  // clamped lanes are masked out below
            const bool inb = jb + lane < jn;
            float dot[kKnnQB];
            if (C3) {
                const float c0 = sTile[jj], c1 = sTile[J + jj], c2 = sTile[2 * J + jj];
#pragma unroll
                for (int q = 0; q < kKnnQB; ++q) {
                    float d = __fmul_rn(q0[q], c0);
                    d = __fmaf_rn(q1[q], c1, d);
                    dot[q] = __fmaf_rn(q2[q], c2, d);
                }
            } else {
#pragma unroll
                for (int q = 0; q < kKnnQB; ++q) dot[q] = 0.f;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    const float cv = sTile[(size_t)c * J + jj];
                    const float4 qa = *reinterpret_cast<const float4*>(myQ + c * kKnnQB);
                    const float4 qb = *reinterpret_cast<const float4*>(myQ + c * kKnnQB + 4);
                    dot[0] = __fmaf_rn(qa.x, cv, dot[0]);
                    dot[1] = __fmaf_rn(qa.y, cv, dot[1]);
                    dot[2] = __fmaf_rn(qa.z, cv, dot[2]);
                    dot[3] = __fmaf_rn(qa.w, cv, dot[3]);
                    dot[4] = __fmaf_rn(qb.x, cv, dot[4]);
                    dot[5] = __fmaf_rn(qb.y, cv, dot[5]);
                    dot[6] = __fmaf_rn(qb.z, cv, dot[6]);
                    dot[7] = __fmaf_rn(qb.w, cv, dot[7]);
                }
            }
            const float nxx = -sXX[jj];
            const unsigned jkey = 0xffffffffu - (unsigned)(j0 + jb + lane);
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q) {
                const float pd = __fsub_rn(__fmaf_rn(2.0f, dot[q], nxx), xxi[q]);
                const bool pass = inb && pd > tau[q];
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                if (mask) {  // warp-uniform
                    if (pass) myBuf[q * 64 + cnt[q] + __popc(mask & lt_mask)] = ((u64)f32_to_ord(pd) << 32) | jkey;
                    cnt[q] += __popc(mask);
                    if (cnt[q] >= 32) flush(q);
                }
            }
        }
    }
    if (i0 >= N) return;
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        if (cnt[q] > 0) flush(q);
        if (i0 + q < N && lane < k) idx[((size_t)b * N + i0 + q) * k + lane] = (int64_t)(0xffffffffu - (unsigned)key[q]);
    }
}

// Generic channel count, k <= 32: the scoring is a register-tiled FP32 inner product. A warp owns 8 query points and
// scores 128 candidates per step, 4 consecutive ones per lane: per channel one LDS.128 of candidates, two broadcast
// LDS.128 of query values and 32 FFMAs -- six shared-memory wavefronts per 32 warp-FFMAs (the 32-candidate step of
// knn_bulk_kernel needs three per 8). Selection is the bulk top-k above, one ballot per candidate slot.
constexpr int kKnnTC = 4;  // candidates per lane per step
constexpr int kGemmThreads = 512, kGemmWarps = kGemmThreads / 32;  // 16 warps share one candidate tile

__global__ void __launch_bounds__(kGemmThreads, 1)
knn_gemm_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int k, int J, int64_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sTile = reinterpret_cast<float*>(smem_raw);                 // [C][J], J a multiple of 128
    float* sXX = sTile + (size_t)C * J;                                // [J]
    float* sQ = sXX + J;                                               // [kGemmWarps][C][kKnnQB]
    u64* sBuf = reinterpret_cast<u64*>(sQ + (size_t)kGemmWarps * C * kKnnQB);  // [kGemmWarps][kKnnQB][64]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int b = blockIdx.y;
    const int i0 = (blockIdx.x * kGemmWarps + w) * kKnnQB;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    float* myQ = sQ + (size_t)w * C * kKnnQB;
    u64* myBuf = sBuf + (size_t)w * kKnnQB * 64;

    float xxi[kKnnQB], tau[kKnnQB];
    u64 key[kKnnQB];
    int cnt[kKnnQB];
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        xxi[q] = __ldg(xxb + min(i0 + q, N - 1));
        tau[q] = -CUDART_INF_F;
        key[q] = 0ull;
        cnt[q] = 0;
    }
    for (int e = lane; e < C * kKnnQB; e += 32) {
        const int c = e / kKnnQB, q = e - c * kKnnQB;
        myQ[e] = __ldg(xb + (size_t)c * N + min(i0 + q, N - 1));
    }
    __syncwarp();

    auto flush = [&](int q) {
        key[q] = knn_flush_list(key[q], myBuf + q * 64, cnt[q]);
        cnt[q] = 0;
        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(key[q] >> 32), k - 1);
        tau[q] = hi ? ord_to_f32(hi) : -CUDART_INF_F;
    };

    for (int j0 = 0; j0 < N; j0 += J) {
        const int jn = min(J, N - j0);
        __syncthreads();  // previous tile fully consumed
        knn_load_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid, kGemmThreads);
        __syncthreads();
        if (i0 >= N) continue;  // warp-uniform; the warp still takes part in the barriers

        for (int jb = 0; jb < jn; jb += 32 * kKnnTC) {
            const int jl = jb + kKnnTC * lane;  // this lane's first candidate inside the tile
            float dot[kKnnQB][kKnnTC];
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q)
#pragma unroll
                for (int t = 0; t < kKnnTC; ++t) dot[q][t] = 0.f;
            const float* tp = sTile + jl;
#pragma unroll 2
            for (int c = 0; c < C; ++c) {
                const float4 cv = *reinterpret_cast<const float4*>(tp + (size_t)c * J);
                const float4 qa = *reinterpret_cast<const float4*>(myQ + c * kKnnQB);
                const float4 qb = *reinterpret_cast<const float4*>(myQ + c * kKnnQB + 4);
                const float qv[kKnnQB] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                const float cc[kKnnTC] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int q = 0; q < kKnnQB; ++q)
#pragma unroll
                    for (int t = 0; t < kKnnTC; ++t) dot[q][t] = __fmaf_rn(qv[q], cc[t], dot[q][t]);
            }
            const float4 nx = *reinterpret_cast<const float4*>(sXX + jl);
            const float nxx[kKnnTC] = {-nx.x, -nx.y, -nx.z, -nx.w};
#pragma unroll
            for (int t = 0; t < kKnnTC; ++t) {
                const bool inb = jl + t < jn;
                const unsigned jkey = 0xffffffffu - (unsigned)(j0 + jl + t);
#pragma unroll
                for (int q = 0; q < kKnnQB; ++q) {
                    const float pd = __fsub_rn(__fmaf_rn(2.0f, dot[q][t], nxx[t]), xxi[q]);
                    const bool pass = inb && pd > tau[q];
                    const unsigned mask = __ballot_sync(0xffffffffu, pass);
                    if (mask) {  // warp-uniform
                        if (pass) myBuf[q * 64 + cnt[q] + __popc(mask & lt_mask)] = ((u64)f32_to_ord(pd) << 32) | jkey;
                        cnt[q] += __popc(mask);
                        if (cnt[q] >= 32) flush(q);
                    }
                }
            }
        }
    }
    if (i0 >= N) return;
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) {
        if (cnt[q] > 0) flush(q);
        if (i0 + q < N && lane < k) idx[((size_t)b * N + i0 + q) * k + lane] = (int64_t)(0xffffffffu - (unsigned)key[q]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// C = 3, k <= 32: threshold selection (the DGCNN xyz layer, BASELINE config c3).
//
// The whole cloud sits in shared memory as (x, y, z, -xx).
//   pass 1  one query per thread, every lane of a warp reads the same candidate (broadcast LDS.128): no shuffles,
//           no ballots, no divergence. Only the maximum score of each of <= 128 contiguous "fine" candidate
//           groups is kept (one float per group, in shared memory). Four warps share a group of 32 queries and
//           scan a quarter of the candidates each, so that small batches still fill the chip.
//   thresh  64 coarse group maxima (consecutive fine groups, kept as 16-bit lower bounds) per query; their k-th largest value T is a
//           lower bound of the k-th largest score, because k different candidates reach it: every thread sorts its 16 with
//           a bitonic network in registers, the query's first thread merges the four runs. About 24 candidates reach T
//           for k = 20, N = 1024, and more than 32 do for one query in a thousand.
//   lists   every thread lists the fine groups of its quarter whose maximum reaches T (about 22 of 128 per query).
//   rescore the four lists of a query are dealt round-robin to its four threads; each re-scores its groups' candidates and
//           appends the indices of those with pd >= T to its own shared-memory column.
//   final   warp w takes the CTA's queries 8w .. 8w+7, four lanes per query: 8 keys of (order-preserving pd bits, ~index)
//           per lane, a 32-key bitonic network spread over the four lanes' registers (shuffles only), the first k indices
//           staged per warp and written as contiguous stores.
// A query with more than 32 candidates >= T, a column overflow or fewer than k candidates (NaN scores) is sorted by its whole
// warp: two bitonic rounds up to 64 candidates, else (heavily duplicated points: everything ties with T) a streaming
// sort-and-merge over all N candidates, so the result is exact for any input.
// pd is computed by the same instruction sequence everywhere (see the file header), so the stages agree bit for
// bit. Padded candidates carry NaN in .w: their pd is NaN, max.f32 ignores it and `>=` rejects it.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSelS = 4;                        // warps that share 32 queries in pass 1
// A CTA holds QW groups of 32 queries (QW = 1 or 2: 128 or 256 threads). c3 (N = 1024) takes QW = 1: 1024 CTAs of 31 KB, seven per
// SM, fill 148 SMs almost evenly (6.9 per SM), where 512 CTAs of 50 KB leave 68 SMs with four and 80 with three.
template <int QW>
struct SelCfg {
    static constexpr int Threads = 128 * QW;
    static constexpr int Q = 32 * QW;                                   // queries per CTA
    static constexpr int ColCap = QW == 1 ? 40 : 48;                    // candidate slots per thread column (one of the four shares of a query)
    static constexpr size_t RegionA = (size_t)ColCap * Threads * 2;     // fine maxima + coarse maxima, later the candidate columns
    static constexpr size_t RegionB = (size_t)Q * 66 * 2;               // (tail of the coarse maxima,) group lists, sorted indices
    static constexpr size_t FmBytes = (size_t)Q * 130 * 2;              // fine maxima, at the head of region A
};
constexpr int kSelGroups = 64;                  // coarse groups (>= k): their k-th largest maximum is the threshold
constexpr int kSelGpp = kSelGroups / kSelS;     // coarse groups per part (sorted in that thread's registers)
constexpr int kSelFine = 128;                   // most fine groups per cloud
constexpr int kSelFmStride = kSelFine + 2;      // 16-bit elements per query row of the fine maxima (65 words: conflict-free)
constexpr int kSelCap = 64;                     // candidates per query
constexpr int kSelMaxN = 8192;                  // the cloud must fit shared memory as float4

__device__ __forceinline__ float knn_pd3(const float4 q, const float xxi, const float4 c) {
    float dot = __fmul_rn(q.x, c.x);
    dot = __fmaf_rn(q.y, c.y, dot);
    dot = __fmaf_rn(q.z, c.z, dot);
    return __fsub_rn(__fmaf_rn(2.0f, dot, c.w), xxi);
}


// rn(2 dot - xx_j): pd before the query's own norm is subtracted. x -> rn(x - xx_i) is monotone, so maxima may be taken on this
// value and the subtraction applied once per group (same bits as the maximum of the pd values; NaN paddings are ignored by max)
__device__ __forceinline__ float knn_t3(const float4 q, const float4 c) {
    float dot = __fmul_rn(q.x, c.x);
    dot = __fmaf_rn(q.y, c.y, dot);
    dot = __fmaf_rn(q.z, c.z, dot);
    return __fmaf_rn(2.0f, dot, c.w);
}

// bitonic networks on a thread's registers (fully unrolled: static indices only)
template <int NV>
__device__ __forceinline__ void regs_merge_desc(float (&v)[NV]) {  // bitonic sequence -> descending
#pragma unroll
    for (int d = NV >> 1; d > 0; d >>= 1) {
#pragma unroll
        for (int e = 0; e < NV; ++e) {
            if ((e & d) == 0) {
                const float hi = fmaxf(v[e], v[e | d]), lo = fminf(v[e], v[e | d]);
                v[e] = hi;
                v[e | d] = lo;
            }
        }
    }
}
template <int NV>
__device__ __forceinline__ void regs_sort_desc(float (&v)[NV]) {
#pragma unroll
    for (int s = 2; s <= NV; s <<= 1) {
#pragma unroll
        for (int d = s >> 1; d > 0; d >>= 1) {
#pragma unroll
            for (int e = 0; e < NV; ++e) {
                const int p = e ^ d;
                if (p > e) {
                    const bool desc = (s == NV) || ((e & s) == 0);
                    const float hi = fmaxf(v[e], v[p]), lo = fminf(v[e], v[p]);
                    v[e] = desc ? hi : lo;
                    v[p] = desc ? lo : hi;
                }
            }
        }
    }
}

__device__ __forceinline__ void u64_cmpx(u64& a, u64& b) {  // a = larger, b = smaller
    const bool sw = a < b;
    const u64 hi = sw ? b : a, lo = sw ? a : b;
    a = hi;
    b = lo;
}
template <int NV>
__device__ __forceinline__ void regs_merge_desc(u64 (&v)[NV]) {  // bitonic sequence -> descending
#pragma unroll
    for (int d = NV >> 1; d > 0; d >>= 1) {
#pragma unroll
        for (int e = 0; e < NV; ++e)
            if ((e & d) == 0) u64_cmpx(v[e], v[e | d]);
    }
}
template <int NV>
__device__ __forceinline__ void regs_sort_desc(u64 (&v)[NV]) {
#pragma unroll
    for (int s = 2; s <= NV; s <<= 1) {
#pragma unroll
        for (int d = s >> 1; d > 0; d >>= 1) {
#pragma unroll
            for (int e = 0; e < NV; ++e) {
                const int p = e ^ d;
                if (p > e) {
                    if ((s == NV) || ((e & s) == 0))
                        u64_cmpx(v[e], v[p]);
                    else
                        u64_cmpx(v[p], v[e]);
                }
            }
        }
    }
}

// sorts two independent key sets at once (the two shuffle chains overlap: the stage is latency-bound otherwise)
__device__ __forceinline__ void warp_sort_desc2(u64& a, u64& b, int lane) {
#pragma unroll
    for (int s = 2; s <= 32; s <<= 1) {
        const bool desc = s == 32 ? true : ((lane & s) == 0);
#pragma unroll
        for (int d = s >> 1; d > 0; d >>= 1) {
            const bool tm = ((lane & d) == 0) == desc;
            a = key_cmpx(a, d, tm);
            b = key_cmpx(b, d, tm);
        }
    }
}
__device__ __forceinline__ void warp_merge_desc2(u64& a, u64& b, int lane) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a = key_cmpx(a, d, (lane & d) == 0);
        b = key_cmpx(b, d, (lane & d) == 0);
    }
}

struct KnnSelPlan {
    int fsh;     // log2(candidates per fine group), >= 3
    int nfine;   // fine groups (<= kSelFine), all holding at least one real candidate
    int npad;    // nfine << fsh
    int qw;      // groups of 32 queries per CTA
    size_t smem;
};
template <int QW>
constexpr bool sel_cfg_ok() {
    using Cf = SelCfg<QW>;
    return Cf::FmBytes % 16 == 0 && Cf::FmBytes + (size_t)kSelGroups * Cf::Q * sizeof(unsigned short) <= Cf::RegionA + Cf::RegionB  // coarse maxima behind the fine ones
           && Cf::RegionB >= (size_t)(kSelFine / kSelS) * Cf::Threads                                                              // the per-thread group lists
           && (size_t)(Cf::Threads / 32) * 256 * sizeof(unsigned short) <= Cf::RegionB;                                            // the warps' sorted indices
}
static_assert(sel_cfg_ok<1>() && sel_cfg_ok<2>(), "shared-memory regions of knn_select_kernel");
template <int QW>
static size_t knn_select_smem(int ntile) {
    using Cf = SelCfg<QW>;
    return (size_t)ntile * sizeof(float4) + Cf::RegionA + Cf::RegionB + (size_t)Cf::Q * sizeof(float) + 2 * (size_t)Cf::Threads;
}
static KnnSelPlan knn_select_plan(int N) {
    KnnSelPlan p;
    p.fsh = 3;
    while (((long long)kSelFine << p.fsh) < N) ++p.fsh;
    p.nfine = (N + (1 << p.fsh) - 1) >> p.fsh;
    p.npad = p.nfine << p.fsh;
    const int ntile = p.npad + (1 << p.fsh);
    // small CTAs when seven of them fit an SM (N <= 1024: 31 KB each), else the 256-thread shape (N = 2048: 66 KB, three per SM)
    p.qw = (knn_select_smem<1>(ntile) + 1024) * 7 <= 228 * 1024 ? 1 : 2;
    p.smem = p.qw == 1 ? knn_select_smem<1>(ntile) : knn_select_smem<2>(ntile);
    return p;
}

template <int QW>
__global__ void __launch_bounds__(SelCfg<QW>::Threads)
knn_select_kernel(const float* __restrict__ x, int N, int k, int nfine, int fsh, int64_t* __restrict__ idx) {
    constexpr int kSelThreads = SelCfg<QW>::Threads, kSelQ = SelCfg<QW>::Q, kSelColCap = SelCfg<QW>::ColCap;
    constexpr size_t kSelRegionA = SelCfg<QW>::RegionA, kSelRegionB = SelCfg<QW>::RegionB, kSelFmBytes = SelCfg<QW>::FmBytes;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npad = nfine << fsh;
    const int ntile = npad + (1 << fsh);  // one more fine group, all padding: where idle lanes point
    float4* sTile = reinterpret_cast<float4*>(smem_raw);                              // [ntile] (x, y, z, -xx)
    unsigned char* regA = reinterpret_cast<unsigned char*>(sTile + ntile);
    unsigned char* regB = regA + kSelRegionA;
    unsigned short* sFm = reinterpret_cast<unsigned short*>(regA);                    // [Q][130] fine group maxima, rounded UP to 16 bits ...
    unsigned short* sBuf = reinterpret_cast<unsigned short*>(regA);                   // ... then [48][T] candidate columns
    unsigned short* sScr = reinterpret_cast<unsigned short*>(regA + kSelFmBytes);     // [64][Q] coarse maxima rounded DOWN to 16 bits (then sorted): tail of A (+ B) ...
    unsigned char* sGl = regB;                                                        // ... then [32][T] fine groups to re-score ...
    float* sThr = reinterpret_cast<float*>(regB + kSelRegionB);                       // [Q]
    signed char* sCntP = reinterpret_cast<signed char*>(sThr + kSelQ);                // [T] candidates found by a thread, -1 = overflow
    unsigned char* sNgl = reinterpret_cast<unsigned char*>(sCntP + kSelThreads);      // [T] fine groups a thread has listed

    // the warp index through a shuffle: the compiler then knows it is warp-uniform (and with it every loop bound derived from it)
    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = w / kSelS, part = w - wq * kSelS;  // warps wq*4 .. wq*4+3 share 32 queries
    const int ql = wq * 32 + lane;                    // query of this thread inside the CTA
    const int b = blockIdx.y;
    const float* __restrict__ xb = x + (size_t)b * 3 * N;

    // the cloud and its squared norms (xx = rn(rn(x0^2 + x1^2) + x2^2), model.py:15)
    for (int j = tid; j < ntile; j += kSelThreads) {
        float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
        if (j < N) {
            v.x = __ldg(xb + j);
            v.y = __ldg(xb + (size_t)N + j);
            v.z = __ldg(xb + 2 * (size_t)N + j);
            v.w = -__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
        }
        sTile[j] = v;
    }
    __syncthreads();

    const int i = blockIdx.x * kSelQ + ql;
    const float4 q = sTile[min(i, N - 1)];  // rows past N are clamped: computed, never written
    const float xxi = -q.w;
    const int f0 = (part * nfine) / kSelS, f1 = ((part + 1) * nfine) / kSelS;  // this thread's share of the fine groups
    const int octs = 1 << (fsh - 3);

    // ---- pass 1: maxima of this warp's quarter of the fine groups, one query per thread ----
    // The fine maxima are kept rounded UP to 16 bits (they only decide which groups are re-scored: a superset is fine);
    // the threshold needs exact values and gets the exact maxima of the 64 coarse groups (consecutive fine groups).
    {
        unsigned short* fm = sFm + ql * kSelFmStride;
        auto up16 = [](float v) {  // smallest 16-bit-truncated float >= v (truncation moves negative values up already)
            const unsigned u = __float_as_uint(v);
            return (unsigned short)(((u >> 31) ? u : u + 0xffffu) >> 16);
        };
        // largest 16-bit-truncated float <= v: the threshold may be any lower bound of the k-th largest score (a lower one lets a few
        // more candidates through: 2^-8 of |T| is a tenth of a candidate at c3), and 16-bit coarse maxima halve their shared memory
        auto down16 = [](float v) {
            const unsigned u = __float_as_uint(v);
            return (unsigned short)(((u >> 31) ? min(u + 0xffffu, 0xff800000u) : u) >> 16);  // (negative: magnitude up, saturating at -INF)
        };
        if (fsh == 3 && nfine == 2 * kSelGroups) {  // 896 < N <= 1024 (c3): every coarse group is exactly two octets, no loop control at all
#pragma unroll 1
            for (int c = part * kSelGpp; c < (part + 1) * kSelGpp; ++c) {
                const float4* tp = sTile + (c << 4);
                float m2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float p[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) p[u] = knn_t3(q, tp[8 * h + u]);
                    m2[h] = max3(max3(max3(p[0], p[1], p[2]), p[3], p[4]), max3(p[5], p[6], p[7]), p[0]);
                }
                *reinterpret_cast<unsigned*>(fm + 2 * c) = (unsigned)up16(__fsub_rn(m2[0], xxi)) | ((unsigned)up16(__fsub_rn(m2[1], xxi)) << 16);
                sScr[c * kSelQ + ql] = down16(__fsub_rn(fmaxf(m2[0], m2[1]), xxi));
            }
        } else
#pragma unroll 1
        for (int c = part * kSelGpp; c < (part + 1) * kSelGpp; ++c) {
            const int fb = (c * nfine) / kSelGroups, fe = ((c + 1) * nfine) / kSelGroups;  // at most 2 fine groups
            float cmax = -CUDART_INF_F;  // maxima are taken before the query's norm is subtracted (knn_t3)
            if (fsh == 3) {  // one octet per fine group (N <= 1024): no inner loop
#pragma unroll 2
                for (int f = fb; f < fe; ++f) {
                    const float4* tp = sTile + (f << 3);
                    float p[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) p[u] = knn_t3(q, tp[u]);
                    const float m = max3(max3(max3(p[0], p[1], p[2]), p[3], p[4]), max3(p[5], p[6], p[7]), p[0]);
                    fm[f] = up16(__fsub_rn(m, xxi));
                    cmax = fmaxf(cmax, m);
                }
            } else {
                for (int f = fb; f < fe; ++f) {
                    const float4* tp = sTile + (f << fsh);
                    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll 2
                    for (int o = 0; o < octs; ++o, tp += 8) {
                        float p[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) p[u] = knn_t3(q, tp[u]);
                        m0 = max3(max3(m0, p[0], p[1]), p[2], p[3]);
                        m1 = max3(max3(m1, p[4], p[5]), p[6], p[7]);
                    }
                    const float m = max3(m0, m0, m1);
                    fm[f] = up16(__fsub_rn(m, xxi));
                    cmax = fmaxf(cmax, m);
                }
            }
            sScr[c * kSelQ + ql] = down16(__fsub_rn(cmax, xxi));
        }
    }

    // ---- threshold: k-th largest of the 64 coarse group maxima. Every thread sorts the 16 maxima it has just produced (no
    // barrier needed: its own values); the first thread of a query then merges the four sorted runs, keeping the 32 largest ----
    {
        float v[kSelGpp];
#pragma unroll
        for (int r = 0; r < kSelGpp; ++r) v[r] = __uint_as_float((unsigned)sScr[(part * kSelGpp + r) * kSelQ + ql] << 16);
        regs_sort_desc<kSelGpp>(v);
        if (part != 0) {
#pragma unroll
            for (int r = 0; r < kSelGpp; ++r) sScr[(part * kSelGpp + r) * kSelQ + ql] = (unsigned short)(__float_as_uint(v[r]) >> 16);
        }
        __syncthreads();
        if (part == 0) {
            float t[32];
            // run 0 (registers) and run 1: the half-cleaner leaves the 16 largest of the 32 in t[0..15], all >= t[16..31], both bitonic
#pragma unroll
            for (int r = 0; r < kSelGpp; ++r) {
                const float o = __uint_as_float((unsigned)sScr[(kSelGpp + (kSelGpp - 1 - r)) * kSelQ + ql] << 16);
                t[r] = fmaxf(v[r], o);
                t[kSelGpp + r] = fminf(v[r], o);
            }
            {
                float hi[kSelGpp], lo[kSelGpp];
#pragma unroll
                for (int r = 0; r < kSelGpp; ++r) hi[r] = t[r], lo[r] = t[kSelGpp + r];
                regs_merge_desc<kSelGpp>(hi);
                regs_merge_desc<kSelGpp>(lo);
#pragma unroll
                for (int r = 0; r < kSelGpp; ++r) t[r] = hi[r], t[kSelGpp + r] = lo[r];
            }
            // runs 2 and 3: the 32 largest of (t, run) are max(t[i], run[31 - i]) with the run padded by -inf: a bitonic sequence again
#pragma unroll
            for (int run = 2; run < kSelS; ++run) {
#pragma unroll
                for (int r = 0; r < kSelGpp; ++r) t[31 - r] = fmaxf(t[31 - r], __uint_as_float((unsigned)sScr[(run * kSelGpp + r) * kSelQ + ql] << 16));
                regs_merge_desc<32>(t);
            }
            // t[k-1] without dynamic register indexing: through shared memory (this thread's own slots)
#pragma unroll
            for (int c = 0; c < 32; ++c) sScr[c * kSelQ + ql] = (unsigned short)(__float_as_uint(t[c]) >> 16);  // (all 16-bit values: lossless)
            sThr[ql] = __uint_as_float((unsigned)sScr[(k - 1) * kSelQ + ql] << 16);
        }
    }
    __syncthreads();
    const float thr = sThr[ql];

    // ---- this thread's fine groups whose maximum reaches the threshold: a private list (branch-free append) ----
    int ngl = 0;
    {
        const unsigned short* fm = sFm + ql * kSelFmStride;
        const unsigned gl0 = static_cast<unsigned>(__cvta_generic_to_shared(sGl + tid));
        unsigned gp = gl0;  // an unconditional byte store and a predicated cursor bump per group (inline PTX, as in the re-scoring below)
        auto push = [&](int f, unsigned bits) {
            asm volatile(
                "{\n\t.reg .pred pp;\n\t"
                "st.shared.u8 [%0], %1;\n\t"
                "setp.ge.f32 pp, %2, %3;\n\t"
                "@pp add.u32 %0, %0, %4;\n\t}"
                : "+r"(gp)
                : "r"(f), "f"(__uint_as_float(bits)), "f"(thr), "n"(kSelThreads)
                : "memory");
        };
        if (((f0 | f1) & 1) == 0) {  // two maxima per load (rows are 4-byte aligned: kSelFmStride is even)
#pragma unroll 4
            for (int f = f0; f < f1; f += 2) {
                const unsigned two = *reinterpret_cast<const unsigned*>(fm + f);
                push(f, two << 16);
                push(f + 1, two & 0xffff0000u);
            }
        } else {
            for (int f = f0; f < f1; ++f) push(f, (unsigned)fm[f] << 16);
        }
        ngl = (int)((gp - gl0) / kSelThreads);
    }
    sNgl[tid] = (unsigned char)ngl;
    __syncthreads();  // every read of the fine maxima is done: their region becomes the candidate columns

    // ---- re-score the listed groups: indices of the candidates with pd >= thr go to this thread's column ----
    // The four lists of a query are taken as one (positions 0 .. total-1) and dealt round-robin to its four threads: a warp walks
    // max over its lanes of ceil(total / 4) groups instead of the longest private list (7.5 instead of 10.5 at c3).
    {
        const int qb = wq * kSelS * 32 + lane;
        const int b1 = sNgl[qb], b2 = b1 + sNgl[qb + 32], b3 = b2 + sNgl[qb + 64], total = b3 + sNgl[qb + 96];
        const int wmax = __reduce_max_sync(0xffffffffu, (total - part + kSelS - 1) >> 2);
        int rot[8];  // (u + lane) & 7, pinned: the compiler otherwise recomputes all eight in every iteration of the group loop
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("and.b32 %0, %1, 7;" : "=r"(rot[u]) : "r"(u + lane));
        const unsigned col = static_cast<unsigned>(__cvta_generic_to_shared(sBuf + tid));
        const unsigned stepb = 2u * kSelThreads;
        const unsigned limit = col + (kSelColCap - 8) * stepb;
        unsigned p = col;
        bool ovf = false;
#pragma unroll 1
        for (int g = 0; g < wmax; ++g) {
            const int pos = g * kSelS + part;
            int f = nfine;  // idle lanes score the all-padding group
            if (pos < total) {
                const int pp = (pos >= b1 ? 1 : 0) + (pos >= b2 ? 1 : 0) + (pos >= b3 ? 1 : 0);  // whose list, and where in it
                const int li = pos - (pos >= b3 ? b3 : pos >= b2 ? b2 : pos >= b1 ? b1 : 0);
                f = (int)sGl[li * kSelThreads + qb + pp * 32];
            }
            const int j0 = f << fsh;
#pragma unroll 1
            for (int o = 0; o < octs; ++o) {
                // fine groups start on 128-byte boundaries, so lanes reading "their" candidate u of different groups would
                // all hit the same four banks: every lane walks its octet rotated by its lane number instead
                const float4* tp = sTile + j0 + 8 * o;
                float pd[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pd[u] = knn_pd3(q, xxi, tp[rot[u]]);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned short jv = (unsigned short)(j0 + 8 * o + rot[u]);
                    asm volatile(
                        "{\n\t.reg .pred pp;\n\t"
                        "st.shared.u16 [%0], %1;\n\t"
                        "setp.ge.f32 pp, %2, %3;\n\t"
                        "@pp add.u32 %0, %0, %4;\n\t}"
                        : "+r"(p)
                        : "h"(jv), "f"(pd[u]), "f"(thr), "r"(stepb)
                        : "memory");
                }
                if (p > limit) {
                    ovf = true;
                    p = limit;
                }
            }
        }
        sCntP[tid] = ovf ? (signed char)-1 : (signed char)((p - col) / stepb);
    }
    __syncthreads();

    // ---- final: the candidates of a query (its four columns taken as one list) sorted, the first k indices written ----
    // Common case, k <= candidates <= 32: warp w takes the CTA's queries 8w .. 8w+7, four lanes per query. The 32 keys of a query are
    // a bitonic network spread over the registers of its four lanes -- 8 keys each, sorted locally, then three exchange steps with
    // the partner lane (its keys read reversed or straight, the larger or the smaller half kept) and local 8-key merges in between:
    // one instruction serves 32 (query, quarter) pairs, where a warp-wide network serves one query (8 M -> 2 M warp-instructions at
    // c3), and nothing but shuffles passes between lanes -- no barrier, no shared-memory round trip.
    {
        const unsigned short* sCol = sBuf;
        const int ql2 = w * 8 + (lane >> 2), pr = lane & 3;
        const int i2 = blockIdx.x * kSelQ + ql2;
        const int colbase = (ql2 >> 5) * (kSelS * 32) + (ql2 & 31);  // column of the query's first share; the others follow 32 apart
        const int c0 = sCntP[colbase], c1 = sCntP[colbase + 32], c2 = sCntP[colbase + 64], c3 = sCntP[colbase + 96];
        const bool ovf = (c0 | c1 | c2 | c3) < 0;
        const int b1 = c0, b2 = b1 + c1, b3 = b2 + c2, total = b3 + c3;
        const bool fast = i2 < N && !ovf && total >= k && total <= 32;
        const float4 q2 = sTile[min(i2, N - 1)];
        // entry e of a query's list: which column, and where in it
        auto cand = [&](int e, int cb, int e1, int e2, int e3) {
            const int pp = (e >= e1 ? 1 : 0) + (e >= e2 ? 1 : 0) + (e >= e3 ? 1 : 0);
            const int li = e - (e >= e3 ? e3 : e >= e2 ? e2 : e >= e1 ? e1 : 0);
            return (int)sCol[li * kSelThreads + cb + pp * 32];
        };
        u64 key[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            key[r] = 0ull;
            if (fast && pr * 8 + r < total) {
                const int j = cand(pr * 8 + r, colbase, b1, b2, b3);
                key[r] = knn_key_of(knn_pd3(q2, -q2.w, sTile[j]), j);
            }
        }
        regs_sort_desc<8>(key);
        auto exchange = [&](const int mask, const bool reversed, const bool keep_max) {
            u64 o[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) o[r] = __shfl_xor_sync(0xffffffffu, key[reversed ? 7 - r : r], mask);
#pragma unroll
            for (int r = 0; r < 8; ++r) key[r] = ((key[r] > o[r]) == keep_max) ? key[r] : o[r];
        };
        exchange(1, true, (pr & 1) == 0);  // 2 x 8 -> 16 (two bitonic halves), sorted by the local merge
        regs_merge_desc<8>(key);
        exchange(3, true, pr < 2);         // 2 x 16 -> the upper 16 in lanes 0, 1 of the query, the lower 16 in lanes 2, 3
        exchange(1, false, (pr & 1) == 0);
        regs_merge_desc<8>(key);
        // The sorted indices go through (warp-private) shared memory so that the warp's 8 x k indices leave as contiguous 256-byte
        // stores (written from the registers, every store instruction would touch 24 different 32-byte sectors for 8 bytes each).
        unsigned short* so = reinterpret_cast<unsigned short*>(regB) + w * 256;  // [8 queries][32 ranks]; the group lists are dead
        {
            unsigned pk[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                pk[r] = ((0xffffffffu - (unsigned)key[2 * r]) & 0xffffu) | ((0xffffffffu - (unsigned)key[2 * r + 1]) << 16);
            reinterpret_cast<uint4*>(so)[lane] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        const unsigned fmask = __ballot_sync(0xffffffffu, fast);
        __syncwarp();
        {
            const unsigned inv = (65536u + (unsigned)k - 1u) / (unsigned)k;  // e / k for e < 256, k <= 32: exact
            int64_t* __restrict__ out = idx + ((size_t)b * N + (size_t)blockIdx.x * kSelQ + w * 8) * k;
            for (int e = lane; e < 8 * k; e += 32) {
                const int qq = (int)(((unsigned)e * inv) >> 16), r = e - qq * k;
                if ((fmask >> (qq * 4)) & 1u) out[e] = (int64_t)so[qq * 32 + r];
            }
        }

        // The rest -- more than 32 candidates, an overflow of a column, or (NaN scores) fewer than k candidates: the whole warp
        // sorts one such query at a time.
        unsigned todo = __ballot_sync(0xffffffffu, pr == 0 && i2 < N && !fast);
        while (todo != 0u) {  // warp-uniform
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int tot = __shfl_sync(0xffffffffu, ovf ? -1 : total, src);
            const int cb = __shfl_sync(0xffffffffu, colbase, src);
            const int e1 = __shfl_sync(0xffffffffu, b1, src), e2 = __shfl_sync(0xffffffffu, b2, src), e3 = __shfl_sync(0xffffffffu, b3, src);
            const int iq = blockIdx.x * kSelQ + w * 8 + (src >> 2);
            const float4 qp = sTile[iq];
            u64 kk = 0ull;
            if (tot >= k && tot <= kSelCap) {
#pragma unroll 1
                for (int h = 0; h * 32 < tot; ++h) {
                    const int e = h * 32 + lane;
                    u64 kb = 0ull;
                    if (e < tot) {
                        const int j = cand(e, cb, e1, e2, e3);
                        kb = knn_key_of(knn_pd3(qp, -qp.w, sTile[j]), j);
                    }
                    kb = warp_sort_desc(kb, lane);
                    if (h == 0) {
                        kk = kb;
                    } else {
                        const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                        kk = warp_merge_desc(kk > rev ? kk : rev, lane);
                    }
                }
            } else {
                // more than 64 candidates tie with the threshold, a column overflowed, or fewer than k candidates reach a NaN-ridden
                // threshold: streaming sort-and-merge over the whole cloud
                u64 tau = 0ull;
                for (int j0 = 0; j0 < N; j0 += 32) {
                    const int j = j0 + lane;
                    u64 kb = j < N ? knn_key_of(knn_pd3(qp, -qp.w, sTile[j]), j) : 0ull;
                    if (!(kb > tau)) kb = 0ull;
                    if (__any_sync(0xffffffffu, kb != 0ull)) {
                        kb = warp_sort_desc(kb, lane);
                        const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                        kk = warp_merge_desc(kk > rev ? kk : rev, lane);
                        tau = __shfl_sync(0xffffffffu, kk, k - 1);
                    }
                }
            }
            if (lane < k) idx[((size_t)b * N + iq) * k + lane] = (int64_t)(0xffffffffu - (unsigned)kk);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Generic channel count, k <= 32: threshold selection around an undisturbed FP32 contraction (DGCNN layers 2-4:
// C = 64, 64, 128), two kernels per batch chunk.
//   knn_rows_kernel    the register-tiled inner product of knn_gemm_kernel (8 queries x 128 candidates per warp
//                      step, software-pipelined over channels, 16-byte cp.async tiles, two 256-thread CTAs per SM)
//                      with NO selection at all: a lane streams the four scores of each query to a scratch row in
//                      the workspace (one coalesced 512-byte store per query and step). The loop runs at the FP32
//                      pipe's pace (92 % of peak while it runs, ncu).
//   knn_rows_select_kernel  a warp per query, eight warps per CTA, as many CTAs per SM as fit -- the selection is a
//                      chain of shuffles and ballots, so it wants thread-level parallelism, not the 16 warps per
//                      SM the contraction leaves. The warp reads the query's row (coalesced float4, from L2), the
//                      32 lane maxima are the maxima of 32 disjoint candidate groups, so their k-th largest value T
//                      (warp bitonic sort) is a lower bound of the k-th largest score; the scores >= T (about 31
//                      for k = 20, N = 1024) are marked in a 64-bit mask per lane, a warp prefix sum of the counts
//                      gives every lane its output position, the keys are sorted and the first k indices written.
//                      More than 64 candidates >= T (massive ties) -> streaming sort-and-merge over the row.
// Nothing is computed twice; the scratch traffic (4 N^2 bytes per cloud, written once and read once) replaces the
// reference's three (B,N,N) temporaries and its radix top-k, and stays far below the FP32 time of the contraction.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sort_desc_f32(float v, int lane) {
#pragma unroll
    for (int s = 2; s <= 32; s <<= 1) {
        const bool desc = s == 32 ? true : ((lane & s) == 0);
#pragma unroll
        for (int d = s >> 1; d > 0; d >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v, d);
            v = (((lane & d) == 0) == desc) ? fmaxf(v, o) : fminf(v, o);
        }
    }
    return v;
}

constexpr int kRowsMaxN = 2048;  // one bit per score of a lane in a 64-bit mask; 16 float4 of scores per lane
constexpr int kRowsThreads = 256, kRowsWarps = kRowsThreads / 32;  // two CTAs per SM: one computes while the other loads a tile

__global__ void __launch_bounds__(kRowsThreads, 2)
knn_rows_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int J, int npad, int nbuf, float* __restrict__ rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // nbuf (1 or 2) candidate tiles of [C][J] coordinates + [J] norms each (J a multiple of 128): with two, the next
    // tile is in flight (cp.async) while the current one is multiplied
    const size_t tile_floats = (size_t)C * J + J;
    float* sTiles = reinterpret_cast<float*>(smem_raw);
    float* sQ = sTiles + nbuf * tile_floats;            // [kRowsWarps][C][kKnnQB] (+ 8 floats: the pipelined loop reads one channel ahead)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y;
    const int i0 = (blockIdx.x * kRowsWarps + w) * kKnnQB;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    float* myQ = sQ + (size_t)w * C * kKnnQB;
    float* myRows = rows + ((size_t)b * N + min(i0, N - 1)) * npad + 4 * lane;  // + q * npad + step * 128

    float xxi[kKnnQB];
#pragma unroll
    for (int q = 0; q < kKnnQB; ++q) xxi[q] = __ldg(xxb + min(i0 + q, N - 1));
    for (int e = lane; e < C * kKnnQB; e += 32) {
        const int c = e / kKnnQB, q = e - c * kKnnQB;
        myQ[e] = __ldg(xb + (size_t)c * N + min(i0 + q, N - 1));
    }
    __syncwarp();

    if (nbuf == 2) knn_issue_tile(sTiles, sTiles + (size_t)C * J, xb, xxb, C, N, J, 0, min(J, N), tid, kRowsThreads);
    for (int j0 = 0, it = 0; j0 < N; j0 += J, ++it) {
        const int jn = min(J, N - j0);
        float* sTile = sTiles + (nbuf == 2 ? (it & 1) * tile_floats : 0);
        float* sXX = sTile + (size_t)C * J;
        if (nbuf == 2) {
            cp_async_wait_all();  // this thread's share of tile `it` has landed ...
            __syncthreads();      // ... everybody's has, and the other buffer is no longer read
            if (j0 + J < N) {
                float* nt = sTiles + ((it + 1) & 1) * tile_floats;
                knn_issue_tile(nt, nt + (size_t)C * J, xb, xxb, C, N, J, j0 + J, min(J, N - j0 - J), tid, kRowsThreads);
            }
        } else {
            __syncthreads();  // previous tile fully consumed
            knn_load_tile(sTile, sXX, xb, xxb, C, N, J, j0, jn, tid, kRowsThreads);
            __syncthreads();
        }
        if (i0 >= N) continue;  // warp-uniform; the warp still takes part in the barriers

        for (int jb = 0; jb < jn; jb += 32 * kKnnTC) {
            const int jl = jb + kKnnTC * lane;  // this lane's first candidate inside the tile
            float dot[kKnnQB][kKnnTC];
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q)
#pragma unroll
                for (int t = 0; t < kKnnTC; ++t) dot[q][t] = 0.f;
            // software-pipelined over channels: the operands of channel c+1 are loaded while channel c is multiplied
            // (the load past the last channel reads the norms row / the next warp's queries: valid memory, never used)
            const float* tp = sTile + jl;
            const float* qp = myQ;
            float4 cv = *reinterpret_cast<const float4*>(tp);
            float4 qa = *reinterpret_cast<const float4*>(qp), qb = *reinterpret_cast<const float4*>(qp + 4);
#pragma unroll 2
            for (int c = 0; c < C; ++c) {
                tp += J;
                qp += kKnnQB;
                const float4 ncv = *reinterpret_cast<const float4*>(tp);
                const float4 nqa = *reinterpret_cast<const float4*>(qp), nqb = *reinterpret_cast<const float4*>(qp + 4);
                const float qv[kKnnQB] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                const float cc[kKnnTC] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int q = 0; q < kKnnQB; ++q)
#pragma unroll
                    for (int t = 0; t < kKnnTC; ++t) dot[q][t] = __fmaf_rn(qv[q], cc[t], dot[q][t]);
                cv = ncv;
                qa = nqa;
                qb = nqb;
            }
            const float4 nx = *reinterpret_cast<const float4*>(sXX + jl);
            const float nxx[kKnnTC] = {-nx.x, -nx.y, -nx.z, -nx.w};
            float* out = myRows + (j0 + jb);
#pragma unroll
            for (int q = 0; q < kKnnQB; ++q) {
                float pd[kKnnTC];
#pragma unroll
                for (int t = 0; t < kKnnTC; ++t)
                    pd[t] = jl + t < jn ? __fsub_rn(__fmaf_rn(2.0f, dot[q][t], nxx[t]), xxi[q]) : -CUDART_INF_F;  // padding never wins
                if (i0 + q < N) __stcg(reinterpret_cast<float4*>(out + (size_t)q * npad), make_float4(pd[0], pd[1], pd[2], pd[3]));
            }
        }
    }
}

constexpr int kRowsSelWarps = 8;

// kRowsSelV = float4 of scores per lane (npad / 128 <= kRowsSelV)
template <int kRowsSelV>
__global__ void __launch_bounds__(kRowsSelWarps * 32)
knn_rows_select_kernel(const float* __restrict__ rows, int N, int k, int npad, int64_t* __restrict__ idx) {
    __shared__ u64 sKey[kRowsSelWarps][kSelCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int i = blockIdx.x * kRowsSelWarps + w;
    if (i >= N) return;  // warp-uniform; no CTA-wide barrier below
    const float* __restrict__ row = rows + ((size_t)b * N + i) * npad;
    u64* myKey = sKey[w];

    // the row: lane l holds scores 128 s + 4 l .. + 3 of every step s (the layout the scoring kernel wrote)
    float4 p[kRowsSelV];
    float cm = -CUDART_INF_F;
#pragma unroll
    for (int s_ = 0; s_ < kRowsSelV; ++s_) {
        if (128 * s_ < npad) {  // warp-uniform
            p[s_] = __ldcg(reinterpret_cast<const float4*>(row + 128 * s_ + 4 * lane));
            cm = max3(max3(cm, p[s_].x, p[s_].y), p[s_].z, p[s_].w);
        }
    }
    const float thr = __shfl_sync(0xffffffffu, warp_sort_desc_f32(cm, lane), k - 1);

    // scores >= thr as a bit mask (bit 4 s + t), positions by a warp prefix sum of the counts
    u64 mask = 0ull;
#pragma unroll
    for (int s_ = 0; s_ < kRowsSelV; ++s_) {
        if (128 * s_ < npad) {
            const int j = 128 * s_ + 4 * lane;
            const unsigned m4 = (j < N && p[s_].x >= thr ? 1u : 0u) | (j + 1 < N && p[s_].y >= thr ? 2u : 0u) |
                                (j + 2 < N && p[s_].z >= thr ? 4u : 0u) | (j + 3 < N && p[s_].w >= thr ? 8u : 0u);
            mask |= (u64)m4 << (4 * s_);
        }
    }
    const int mine_n = __popcll(mask);
    int incl = mine_n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    const int cnt = __shfl_sync(0xffffffffu, incl, 31);

    u64 key = 0ull;
    if (cnt <= kSelCap) {
        int pos = incl - mine_n;
        while (mask) {
            const int bit = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const int off = (bit >> 2) * 128 + 4 * lane + (bit & 3);
            myKey[pos++] = knn_key_of(__ldcg(row + off), off);
        }
        __syncwarp();
        for (int h = 0; h * 32 < cnt; ++h) {
            const int e = h * 32 + lane;
            const u64 kb = warp_sort_desc(e < cnt ? myKey[e] : 0ull, lane);
            if (h == 0) {
                key = kb;
            } else {
                const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                key = warp_merge_desc(key > rev ? key : rev, lane);
            }
        }
    } else {
        // more than 64 candidates tie with the threshold: streaming sort-and-merge over the scratch row
        u64 tau = 0ull;
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int j = j0 + lane;
            u64 kb = j < N ? knn_key_of(__ldcg(row + j), j) : 0ull;
            if (!(kb > tau)) kb = 0ull;
            if (__any_sync(0xffffffffu, kb != 0ull)) {
                kb = warp_sort_desc(kb, lane);
                const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                key = warp_merge_desc(key > rev ? key : rev, lane);
                tau = __shfl_sync(0xffffffffu, key, k - 1);
            }
        }
    }
    if (lane < k) idx[((size_t)b * N + i) * k + lane] = (int64_t)(0xffffffffu - (unsigned)key);
}

struct KnnPlan {
    int KR;
    bool rows;   // generic C, k <= 32: knn_rows_kernel (scores streamed to scratch rows in the workspace)
    int npad;    // floats per scratch row
    int chunk;   // clouds per launch of knn_rows_kernel: the scratch (chunk * N rows) is reused by the next launch
    int nbuf;    // candidate tiles in shared memory (2 = double-buffered, when two CTAs per SM still fit)
    bool gemm;  // generic C, k <= 32: the register-tiled scoring kernel
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
    p.gemm = false;
    p.rows = false;
    p.npad = 0;
    p.chunk = 0;
    p.nbuf = 1;
    if (p.KR == 1 && C != 3 && !tune_env("FPSG_KNN_INSERT") && !tune_env("FPSG_KNN_NOGEMM")) {
        long long jg = (long long)(kKnnTileBytes / sizeof(float)) / (C + 1) / 128 * 128;
        if (jg < 128) jg = 128;
        const long long npad128 = (N + 127) / 128 * 128;
        if (jg > npad128) jg = npad128;
        const size_t smem = ((size_t)C * jg + jg + (size_t)kGemmWarps * C * kKnnQB) * sizeof(float) + (size_t)kGemmWarps * kKnnQB * 64 * sizeof(u64);
        if (smem <= 220 * 1024) {
            p.gemm = true;
            p.J = (int)jg;
            p.smem = smem;
            p.ws_bytes = align_up((size_t)B * N * sizeof(float), 256);
            const size_t npad = (size_t)(N + 127) / 128 * 128;
            // scratch rows for at most ~256 MB worth of clouds at a time (at least one cloud): launches on the same stream
            // run back to back, so the next chunk can reuse the rows of the previous one
            const size_t cloud_bytes = (size_t)N * npad * sizeof(float);
            size_t chunk = ((size_t)256 << 20) / cloud_bytes;
            if (chunk < 1) chunk = 1;
            if (chunk > (size_t)B) chunk = (size_t)B;
            const size_t row_bytes = chunk * cloud_bytes;
            if (N <= kRowsMaxN && !tune_env("FPSG_KNN_NOROWS")) {
                p.rows = true;
                p.npad = (int)npad;
                p.chunk = (int)chunk;
                const size_t tile_b = ((size_t)C * jg + jg) * sizeof(float), q_b = ((size_t)kRowsWarps * C * kKnnQB + kKnnQB) * sizeof(float);
                p.nbuf = 2 * tile_b + q_b <= 110 * 1024 ? 2 : 1;
                p.smem = p.nbuf * tile_b + q_b;
                p.ws_bytes += align_up(row_bytes, 256);
            }
            return p;
        }
    }
    if (p.KR == 1 && !tune_env("FPSG_KNN_INSERT"))  // bulk top-k kernel: tile + (generic C) query values + candidate buffers
        p.smem = ((size_t)C * p.J + p.J + (C == 3 ? 0 : (size_t)kKnnWarps * C * kKnnQB)) * sizeof(float) + (size_t)kKnnWarps * kKnnQB * 64 * sizeof(u64);
    else if (C == 3)
        p.smem = ((size_t)C * p.J + p.J) * sizeof(float) + (size_t)kKnnRows * 32 * p.KR * (sizeof(float) + sizeof(int));
    else if (p.KR == 4)  // generic C, 64 < k <= 128: knn_kernel<4,false> (tile + norms + one query row per warp + the lists)
        p.smem = ((size_t)C * p.J + p.J + (size_t)kKnnWarps * C) * sizeof(float) + (size_t)kKnnRows * 32 * p.KR * (sizeof(float) + sizeof(int));
    else
        p.smem = ((size_t)C * p.J + p.J + (size_t)kKnnWarps * C * kKnnQB) * sizeof(float);
    p.ws_bytes = align_up((size_t)B * N * sizeof(float), 256);
    return p;
}

// knn_tc.cu: tensor-core pre-filter + exact re-scoring (generic C, k <= 32, 256 <= N <= 8192)
bool knn_tc_eligible(int B, int C, int N, int k);
size_t knn_tc_workspace(int B, int C, int N, int k);
int knn_tc_launch(const float* x, int B, int C, int N, int k, int64_t* idx, void* ws, cudaStream_t stream);

size_t knn_workspace(int B, int C, int N, int k) {
    if (B <= 0 || C < 1 || N < 1 || k < 1) return 0;
    if (C != 3 && knn_tc_eligible(B, C, N, k)) return knn_tc_workspace(B, C, N, k);
    return make_knn_plan(B, C, N, k).ws_bytes;
}

static int knn_launch_rows(const float* x, const float* xx, float* rows, int B, int C, int N, int k, const KnnPlan& p, int64_t* idx, cudaStream_t stream) {
    if (p.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    for (int b0 = 0; b0 < B; b0 += p.chunk) {
        const int nb = min(p.chunk, B - b0);
        knn_rows_kernel<<<dim3((unsigned)cdiv(N, kRowsWarps * kKnnQB), (unsigned)nb), kRowsThreads, p.smem, stream>>>(
            x + (size_t)b0 * C * N, xx + (size_t)b0 * N, C, N, p.J, p.npad, p.nbuf, rows);
        FPSG_CUDA_OK(cudaGetLastError());
        const dim3 sgrid((unsigned)cdiv(N, kRowsSelWarps), (unsigned)nb);
        int64_t* out = idx + (size_t)b0 * N * k;
        if (p.npad <= 512)
            knn_rows_select_kernel<4><<<sgrid, kRowsSelWarps * 32, 0, stream>>>(rows, N, k, p.npad, out);
        else if (p.npad <= 1024)
            knn_rows_select_kernel<8><<<sgrid, kRowsSelWarps * 32, 0, stream>>>(rows, N, k, p.npad, out);
        else
            knn_rows_select_kernel<16><<<sgrid, kRowsSelWarps * 32, 0, stream>>>(rows, N, k, p.npad, out);
        FPSG_CUDA_OK(cudaGetLastError());
    }
    return FPSG_OK;
}

static int knn_launch_gemm(const float* x, const float* xx, int B, int C, int N, int k, const KnnPlan& p, int64_t* idx, cudaStream_t stream) {
    if (p.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const dim3 grid((unsigned)cdiv(N, kGemmWarps * kKnnQB), (unsigned)B);
    knn_gemm_kernel<<<grid, kGemmThreads, p.smem, stream>>>(x, xx, C, N, k, p.J, idx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

template <bool C3>
static int knn_launch_bulk(const float* x, const float* xx, int B, int C, int N, int k, const KnnPlan& p, int64_t* idx, cudaStream_t stream) {
    auto kern = knn_bulk_kernel<C3>;
    if (p.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const dim3 grid((unsigned)cdiv(N, kKnnWarps * kKnnQB), (unsigned)B);
    kern<<<grid, kKnnThreads, p.smem, stream>>>(x, xx, C, N, k, p.J, idx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

template <int KR>
static int knn_launch_blocked(const float* x, const float* xx, int B, int C, int N, int k, const KnnPlan& p, int64_t* idx, cudaStream_t stream) {
    auto kern = knn_blocked_kernel<KR>;
    if (p.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const dim3 grid((unsigned)cdiv(N, kKnnWarps * kKnnQB), (unsigned)B);
    kern<<<grid, kKnnThreads, p.smem, stream>>>(x, xx, C, N, k, p.J, idx);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
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
    if (C != 3 && knn_tc_eligible(B, C, N, k)) {
        const size_t need = knn_tc_workspace(B, C, N, k);
        if (!ws || ws_bytes < need) return fail(FPSG_EWORKSPACE, "knn: workspace of %zu bytes required, %zu given", need, ws ? ws_bytes : (size_t)0);
        if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(FPSG_EINVAL, "knn: the workspace must be 16-byte aligned");
        return knn_tc_launch(x, B, C, N, k, idx, ws, stream);
    }
    if ((size_t)C * 32 * sizeof(float) > 160 * 1024) return fail(FPSG_EINVAL, "knn: C=%d too large", C);
    const KnnPlan p = make_knn_plan(B, C, N, k);
    if (p.smem > 227 * 1024) return fail(FPSG_EINVAL, "knn: C=%d with k=%d needs %zu bytes of shared memory per CTA (limit 227 KB)", C, k, p.smem);
    if (!ws || ws_bytes < p.ws_bytes) return fail(FPSG_EWORKSPACE, "knn: workspace of %zu bytes required, %zu given", p.ws_bytes, ws ? ws_bytes : (size_t)0);
    if (C == 3 && k <= 32 && N <= kSelMaxN && !tune_env("FPSG_KNN_BULK")) {  // threshold-selection kernel (needs no workspace)
        const KnnSelPlan sp = knn_select_plan(N);
        if (sp.qw == 1) {
            if (sp.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_select_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
            knn_select_kernel<1><<<dim3((unsigned)cdiv(N, SelCfg<1>::Q), (unsigned)B), SelCfg<1>::Threads, sp.smem, stream>>>(x, N, k, sp.nfine, sp.fsh, idx);
        } else {
            if (sp.smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_select_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
            knn_select_kernel<2><<<dim3((unsigned)cdiv(N, SelCfg<2>::Q), (unsigned)B), SelCfg<2>::Threads, sp.smem, stream>>>(x, N, k, sp.nfine, sp.fsh, idx);
        }
        FPSG_CUDA_OK(cudaGetLastError());
        return FPSG_OK;
    }
    float* xx = static_cast<float*>(ws);
    knn_sqnorm_kernel<<<dim3((unsigned)cdiv(N, 256), (unsigned)B), 256, 0, stream>>>(x, C, N, xx);
    FPSG_CUDA_OK(cudaGetLastError());
    const bool c3 = (C == 3);
    switch (p.KR) {
        case 1:
            if (tune_env("FPSG_KNN_INSERT"))  // tuning override: the shuffle-insertion kernels also serve k <= 32
                return c3 ? knn_launch_t<1, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_blocked<1>(x, xx, B, C, N, k, p, idx, stream);
            if (p.rows)
                return knn_launch_rows(x, xx, reinterpret_cast<float*>(static_cast<char*>(ws) + align_up((size_t)B * N * sizeof(float), 256)), B, C, N, k,
                                       p, idx, stream);
            if (p.gemm) return knn_launch_gemm(x, xx, B, C, N, k, p, idx, stream);
            return c3 ? knn_launch_bulk<true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_bulk<false>(x, xx, B, C, N, k, p, idx, stream);
        case 2:
            return c3 ? knn_launch_t<2, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_blocked<2>(x, xx, B, C, N, k, p, idx, stream);
        default:
            return c3 ? knn_launch_t<4, true>(x, xx, B, C, N, k, p, idx, stream) : knn_launch_t<4, false>(x, xx, B, C, N, k, p, idx, stream);
    }
}

}  // namespace fpsg
