// K2: Chamfer backward for sm_100a -- gradients gathered through an inverted index, no floating-point atomics.
//
// Replaces Kaolin 0.9.0's sided_distance_backward_cuda_kernel as reached from `.backward()` on the reference's
// loss (reference src/trainNetwork.py:144 -> src/models/few_shot.py:110,117), which
// scatters with float atomicAdd into both gradient tensors from a 16-CTA grid.
//
//   grad_A[b,i] = sum over every nearest-neighbour edge (i <-> x) incident to A_i of fl(2*g_edge * fl(A_i - Q_x))
//     * one "direct" edge  x = idxA[i]          with g_edge = gA[b,i]
//     * "scattered" edges  j : idxQ[b,j] == i   with g_edge = gQ[b,j]
//
// Here every CTA owns 512 consecutive points of one cloud (one thread per point; 512 rather than 256 threads halves the
// redundant index scans and staged copies per pair: c1 8.7 -> 8.1 us). It scans the other cloud's
// arg-min indices in chunks, builds a CSR inverted index of the chunk in shared memory with *integer* atomics
// (counts -> exclusive scan -> cursor fill), and then each thread walks its own bucket and accumulates its
// gradient privately in double. The sum is therefore independent of any floating-point atomic ordering (the
// bucket order is arbitrary, but a double sum of FP32 terms is exact to far below FP32 resolution), matches the
// oracle's double accumulation, and the gradient is written exactly once, coalesced, with no pre-zeroing.
#include "common.cuh"

namespace fpsg {

constexpr int kBwdThreads = 512;
constexpr int kBwdChunk = 2048;  // arg-min indices of the other cloud scanned per CSR build
constexpr size_t kBwdStageMaxBytes = 64 * 1024;  // other clouds up to this size are staged in shared memory

struct BwdSide {
    bool first;            // A = the p1 side of the segment table (gradient to gp1), else the p2 side
    const int64_t* idxA;   // (B,Na)   nearest Q point of every A point
    const int64_t* idxQ;   // (B,Nq)   nearest A point of every Q point
    const float* gA;       // upstream gradient of distA: (B,Na), or (B,) when uniform
    const float* gQ;       // upstream gradient of distQ: (B,Nq), or (B,) when uniform
    float scaleA, scaleQ;  // uniform mode: g = (g[b] * w) / count  -> scale carries w, divisor carries count
    float divA, divQ;
    int Na, Nq;
    int blocks;            // CTAs per pair for this side
};

// uniform mode: the gradient that `.sum()` / kaolin's mean sends to every element of a pair, (gloss * w) / count
__device__ __forceinline__ float uniform_weight(float gloss, float scale, float div) {
    return __fdiv_rn(__fmul_rn(gloss, scale), div);
}

// seg: where the clouds and the gradient outputs live. uniform: the upstream gradient is one value per pair, gA = gQ =
// gloss (B,), nullptr = ones. multi (multi-segment loss, implies uniform): gtotal is a device scalar (nullptr = 1), the
// upstream gradient of the weighted total; the pairs of segment s then receive gloss = gtotal * weight_s.
__global__ void __launch_bounds__(kBwdThreads)
chamfer_bwd_kernel(const SegTable seg, BwdSide s1, BwdSide s2, bool uniform, bool multi, const float* __restrict__ gtotal, bool stage) {
    extern __shared__ __align__(16) float sQ[];  // stage: the other cloud of this pair (every gathered edge ends there)
    __shared__ int sCount[kBwdThreads];
    __shared__ int sCursor[kBwdThreads];
    __shared__ int sWarp[kBwdThreads / 32];
    __shared__ int sList[kBwdChunk];

    cudaTriggerProgrammaticLaunchCompletion();  // a following kernel launched with PDL may start to launch
    cudaGridDependencySynchronize();            // launched with PDL: the forward's outputs are complete from here on
    const bool first = (int)blockIdx.x < s1.blocks;
    const BwdSide& S = first ? s1 : s2;
    const int blk = first ? blockIdx.x : blockIdx.x - s1.blocks;
    const int b = blockIdx.y;
    // (the warp index through a shuffle: the compiler then treats it, and the branches on it, as warp-uniform)
    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int i0 = blk * kBwdThreads;
    const int i = i0 + tid;
    const bool live = i < S.Na;

    const int sg = seg.n == 1 ? 0 : seg_find(seg, b);
    const int bl = b - seg.b0[sg];  // pair index inside its segment
    float* const outb = (S.first ? seg.gp1[sg] : seg.gp2[sg]);
    if (outb == nullptr) return;  // this segment does not want this side's gradient (CTA-uniform)
    const float* __restrict__ Ab = (S.first ? seg.p1[sg] : seg.p2[sg]) + (size_t)bl * S.Na * 3;
    const float* __restrict__ Qb = (S.first ? seg.p2[sg] : seg.p1[sg]) + (size_t)bl * S.Nq * 3;
    const int64_t* __restrict__ iQ = S.idxQ + (size_t)b * S.Nq;

    // 2*g of the two edge kinds when the upstream gradient is uniform over the pair
    float uA = 0.f, uQ = 0.f;
    if (uniform) {
        const float glA = multi ? __fmul_rn(gtotal ? __ldg(gtotal) : 1.0f, seg.weight[sg]) : (S.gA ? __ldg(S.gA + b) : 1.0f);
        const float glQ = multi ? glA : (S.gQ ? __ldg(S.gQ + b) : 1.0f);
        uA = __fmul_rn(2.0f, uniform_weight(glA, S.scaleA, S.divA));
        uQ = __fmul_rn(2.0f, uniform_weight(glQ, S.scaleQ, S.divQ));
    }

    constexpr int kPer = kBwdChunk / kBwdThreads;  // arg-min indices of the other cloud per thread per chunk

    // Staging (clouds up to 64 KB): the other cloud is copied to shared memory asynchronously while the indices are loaded
    // and the inverted index is built, so the gathers below are shared-memory reads instead of dependent L2 round trips.
    // (Issued after the dependency wait: unlike in the forward, the previous kernel in the stream may be the clouds' producer.)
    if (stage) {
        const int qfl = 3 * S.Nq;
        const int q4 = (reinterpret_cast<uintptr_t>(Qb) & 15) == 0 ? qfl >> 2 : 0;
        const unsigned sdst = (unsigned)__cvta_generic_to_shared(sQ);
        for (int q = tid; q < q4; q += kBwdThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 16u * q), "l"(reinterpret_cast<const float4*>(Qb) + q) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int q = 4 * q4 + tid; q < qfl; q += kBwdThreads) sQ[q] = __ldg(Qb + q);
    }

    // issue order matters more than arithmetic here (the kernel is a chain of L2 latencies): own point and its
    // direct edge first, then the first chunk of the other cloud's indices, all in flight together
    float ax = 0.f, ay = 0.f, az = 0.f;
    long long x = 0;
    if (live) {
        // clamped: an out-of-range index in a caller's idx tensor must not become an out-of-bounds read (the forward
        // never produces one); scattered edges that point outside [0, Na) fall in no live thread's bucket below
        x = min(max((long long)S.idxA[(size_t)b * S.Na + i], 0LL), (long long)S.Nq - 1);
        ax = __ldg(Ab + 3 * i);
        ay = __ldg(Ab + 3 * i + 1);
        az = __ldg(Ab + 3 * i + 2);
    }
    long long t[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int jj = tid + q * kBwdThreads;
        t[q] = jj < S.Nq ? iQ[jj] - i0 : -1;
    }
    // the direct edge's other end: its load depends on x, so it is issued here and consumed only after the first
    // inverted-index build below, which hides its latency
    float dqx = 0.f, dqy = 0.f, dqz = 0.f, gdir = 0.f;
    if (live) {
        gdir = uniform ? uA : __fmul_rn(2.0f, __ldg(S.gA + (size_t)b * S.Na + i));
        if (!stage) {
            dqx = __ldg(Qb + 3 * x);
            dqy = __ldg(Qb + 3 * x + 1);
            dqz = __ldg(Qb + 3 * x + 2);
        }
    }
    double gx = 0.0, gy = 0.0, gz = 0.0;
    if (stage) asm volatile("cp.async.wait_group 0;" ::: "memory");  // published by the first barrier of the build below
    const float* __restrict__ Qg = stage ? sQ : Qb;                   // where the gathers read

    for (int j0 = 0; j0 < S.Nq; j0 += kBwdChunk) {
        const int jn = min(kBwdChunk, S.Nq - j0);
        if (j0 > 0) {
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const int jj = tid + q * kBwdThreads;
                t[q] = jj < jn ? iQ[j0 + jj] - i0 : -1;
            }
        }
        sCount[tid] = 0;
        __syncthreads();
        // pass 1: bucket sizes
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (t[q] >= 0 && t[q] < kBwdThreads) atomicAdd(&sCount[(int)t[q]], 1);
        __syncthreads();
        // exclusive scan of the kBwdThreads counts
        const int c = sCount[tid];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) sWarp[w] = incl;
        __syncthreads();
        int base = 0;
        for (int q = 0; q < w; ++q) base += sWarp[q];
        const int start = base + incl - c;
        sCursor[tid] = start;
        __syncthreads();
        // pass 2: fill
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (t[q] >= 0 && t[q] < kBwdThreads) sList[atomicAdd(&sCursor[(int)t[q]], 1)] = j0 + tid + q * kBwdThreads;
        __syncthreads();
        // gather: this thread's bucket, four edges in flight at a time (each is a dependent L2 round trip)
        if (live) {
            if (j0 == 0) {
                if (stage) {
                    dqx = sQ[3 * x];
                    dqy = sQ[3 * x + 1];
                    dqz = sQ[3 * x + 2];
                }
                gx = (double)__fmul_rn(gdir, __fsub_rn(ax, dqx));
                gy = (double)__fmul_rn(gdir, __fsub_rn(ay, dqy));
                gz = (double)__fmul_rn(gdir, __fsub_rn(az, dqz));
            }
            for (int e = start; e < start + c; e += 4) {
                float qx[4], qy[4], qz[4], gg[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const bool on = e + v < start + c;
                    const int j = on ? sList[e + v] : 0;
                    gg[v] = !on ? 0.f : (uniform ? uQ : __fmul_rn(2.0f, __ldg(S.gQ + (size_t)b * S.Nq + j)));
                    qx[v] = on ? Qg[3 * j] : ax;
                    qy[v] = on ? Qg[3 * j + 1] : ay;
                    qz[v] = on ? Qg[3 * j + 2] : az;
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (e + v < start + c) {
                        gx += (double)__fmul_rn(gg[v], __fsub_rn(ax, qx[v]));
                        gy += (double)__fmul_rn(gg[v], __fsub_rn(ay, qy[v]));
                        gz += (double)__fmul_rn(gg[v], __fsub_rn(az, qz[v]));
                    }
                }
            }
        }
        if (j0 + kBwdChunk >= S.Nq) break;  // uniform: no further chunk reuses the shared arrays
        __syncthreads();
    }
    if (live) {
        float* o = outb + ((size_t)bl * S.Na + i) * 3;
        o[0] = (float)gx;
        o[1] = (float)gy;
        o[2] = (float)gz;
    }
}

int chamfer_backward_launch(const SegTable& seg, const int64_t* idx1, const int64_t* idx2, const float* g1, const float* g2,
                            bool uniform, bool multi, const float* gtotal, float w1, float w2, int B, int N, int M, cudaStream_t stream) {
    if (B < 0 || N < 1 || M < 1) return fail(FPSG_EINVAL, "chamfer_bwd: need B >= 0, N >= 1, M >= 1 (got B=%d N=%d M=%d)", B, N, M);
    bool want1 = false, want2 = false;
    for (int sg = 0; sg < seg.n; ++sg) {
        if (seg.b0[sg + 1] == seg.b0[sg]) continue;
        want1 |= seg.gp1[sg] != nullptr;
        want2 |= seg.gp2[sg] != nullptr;
        if (!seg.p1[sg] || !seg.p2[sg]) return fail(FPSG_EINVAL, "chamfer_bwd: null input pointer");
    }
    if (B == 0 || (!want1 && !want2)) return FPSG_OK;
    if (!idx1 || !idx2 || (!uniform && (!g1 || !g2))) return fail(FPSG_EINVAL, "chamfer_bwd: null input pointer");
    if (B > 65535) return fail(FPSG_EINVAL, "chamfer_bwd: B > 65535 not supported in one call");
    BwdSide s1{true, idx1, idx2, g1, g2, w1, w2, (float)N, (float)M, N, M, want1 ? cdiv(N, kBwdThreads) : 0};
    BwdSide s2{false, idx2, idx1, g2, g1, w2, w1, (float)M, (float)N, M, N, want2 ? cdiv(M, kBwdThreads) : 0};
    const dim3 grid((unsigned)(s1.blocks + s2.blocks), (unsigned)B);
    size_t smem = align_up((size_t)(N > M ? N : M) * 12, 16);
    static const bool no_stage = tune_env("FPSG_BWD_NOSTAGE") != nullptr;  // tuning build only
    bool stage = smem <= kBwdStageMaxBytes && !no_stage;
    if (!stage) smem = 0;
    if (smem > 32 * 1024)  // static + dynamic beyond 48 KB needs the (per-device) opt-in
        FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdStageMaxBytes));
    FPSG_CUDA_OK(launch_pdl(chamfer_bwd_kernel, grid, dim3(kBwdThreads), smem, stream, seg, s1, s2, uniform, multi, gtotal, stage));
    return FPSG_OK;
}

}  // namespace fpsg
