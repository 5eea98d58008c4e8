// K1b + K2 fused: the "finish" kernel -- everything that follows the sweep of a Chamfer loss step whose upstream
// gradient is known when the step is launched (the reference's `pc_metric(a, b).sum()` ... `ttl_loss.backward()`,
// src/models/few_shot.py:110,117,121 and src/trainNetwork.py:144: the gradient of the loss is a constant).
//
// One thread-block CLUSTER of 8 CTAs per pair. A CTA owns an eighth of the pair's rows (p1 points) and an eighth of
// its columns (p2 points), one row and one column per thread:
//
//   0. (before the dependency wait) both clouds of the pair are copied to shared memory with cp.async -- they are
//      inputs of the sweep, complete before either kernel started; bucket counters are zeroed; split cluster barrier
//      #1 is signalled.
//   1. merge of the sweep's partial minima (rows: over the CTAs that covered the strip; columns: strips x warps),
//      exactly as the finalize kernel does (chamfer_plan.cuh).
//   2. arg-min recovery out of the staged clouds (thread-private, bit-equal re-evaluation), outputs dist/idx.
//   3. inverted index THROUGH DISTRIBUTED SHARED MEMORY: column j pushes itself into the bucket of the row it is
//      nearest to, in the shared memory of the CTA that owns that row (atom.shared::cluster on the count; four
//      entries per bucket, longer buckets spill to the owner's overflow list), and symmetrically rows into column
//      buckets when a gradient for p2 is wanted. Per-warp loss sums travel the same way to CTA 0. One cluster
//      barrier (#2) publishes all of it: no global round trip, no grid-wide dependency between the pair's
//      finalize and its backward.
//   4. backward: every thread gathers its direct edge and its bucket from the staged clouds and accumulates in
//      double (no floating-point atomics; the bucket order is arbitrary, a double sum of FP32 terms is exact far
//      below FP32 resolution), writes its gradient once. Meanwhile the CTA's extra warp (rank 0 only) adds the loss
//      partials in a fixed order, writes loss[b] and -- for the multi-segment / sharded form -- takes part in the
//      batch-level reduction.
//
// Results are identical to the finalize + backward kernels (same merge, same re-evaluation, same edge terms).
#include <cooperative_groups.h>

#include "chamfer_plan.cuh"

namespace cg = cooperative_groups;

namespace fpsg {

constexpr int kFinishClusterMax = 8;  // CTAs per pair (= cluster size)
constexpr int kBucket = 4;           // inverted-index entries held per point; longer buckets spill to the overflow list
constexpr int kFinishMaxPoints = 4096;  // per cloud: one row and one column per thread, up to 512 threads per CTA

struct FinishArgs {
    SegTable seg;
    int N, M, R, strips, ups, q, r, maxslots, Mpad;
    const float* rowmin;
    const int* rowsub;
    const u64* colpart;
    float* dist1;   // optional outputs
    float* dist2;
    int64_t* idx1;
    int64_t* idx2;
    unsigned* counters;
    float* loss;    // (B,)
    float w1, w2;
    float* totals;  // [1 + segments] or nullptr
    const float* gloss;   // (B,) upstream gradient of loss, or nullptr = ones
    const float* gtotal;  // multi-segment form: device scalar (nullptr = 1), gloss = gtotal * weight_s
    int multi;
    int B;
    int rpcN, rpcM;  // rows / columns per CTA (multiples of 32)
    int N32, M32;    // cloud sizes rounded up to whole 32-point windows
    const unsigned* progress;   // the sweep CTAs' progress words (chamfer_plan.cuh)
    unsigned long long* trace;  // timeline stamps (tuning builds), else nullptr
    PeerExchange px;            // multi-GPU scalar exchange in the batch-level epilogue (world == 0: none)
};

struct FinishSmem {
    size_t p2, p1, cnt1, bkt1, ovf1, cnt2, bkt2, ovf2, head1, head2, misc, loss, total;
};

static FinishSmem finish_smem_layout(int N, int M, int rpcN, int rpcM, bool want2, int nwarps) {
    FinishSmem L;
    size_t off = 0;
    const size_t N32 = align_up((size_t)N, 32), M32 = align_up((size_t)M, 32);
    L.p2 = off;
    off += align_up(M32 * 12, 16);
    L.p1 = off;
    off += align_up(N32 * 12, 16);
    L.cnt1 = off;
    off += align_up((size_t)rpcN * 4, 16);
    L.bkt1 = off;
    off += align_up((size_t)rpcN * kBucket * 2, 16);
    L.ovf1 = off;
    off += align_up((size_t)M * 4, 16);
    L.cnt2 = off;
    if (want2) off += align_up((size_t)rpcM * 4, 16);
    L.bkt2 = off;
    if (want2) off += align_up((size_t)rpcM * kBucket * 2, 16);
    L.ovf2 = off;
    if (want2) off += align_up((size_t)N * 4, 16);
    L.head1 = off;  // per point: head of its chain in the overflow list (0xFFFF = none)
    off += align_up((size_t)rpcN * 4, 16);
    L.head2 = off;
    if (want2) off += align_up((size_t)rpcM * 4, 16);
    L.misc = off;  // [0] overflow count side 1, [1] overflow count side 2
    off += 16;
    L.loss = off;  // [cluster rank][warp][2] doubles (used in rank 0)
    off += (size_t)kFinishClusterMax * nwarps * 2 * sizeof(double);
    L.total = off;
    return L;
}

struct FinishSmemOff {
    unsigned p2, p1, cnt1, bkt1, ovf1, cnt2, bkt2, ovf2, head1, head2, misc, loss;
};

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <bool kWant2>
__global__ void __launch_bounds__(kFinishMaxPoints / kFinishClusterMax + 64, 2) chamfer_finish_kernel(const FinishArgs a, const FinishSmemOff so) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* sP2 = reinterpret_cast<float*>(smem + so.p2);
    float* sP1 = reinterpret_cast<float*>(smem + so.p1);
    unsigned* sCnt1 = reinterpret_cast<unsigned*>(smem + so.cnt1);
    unsigned short* sBkt1 = reinterpret_cast<unsigned short*>(smem + so.bkt1);
    unsigned* sOvf1 = reinterpret_cast<unsigned*>(smem + so.ovf1);
    unsigned* sCnt2 = reinterpret_cast<unsigned*>(smem + so.cnt2);
    unsigned short* sBkt2 = reinterpret_cast<unsigned short*>(smem + so.bkt2);
    unsigned* sOvf2 = reinterpret_cast<unsigned*>(smem + so.ovf2);
    unsigned* sHead1 = reinterpret_cast<unsigned*>(smem + so.head1);
    unsigned* sHead2 = reinterpret_cast<unsigned*>(smem + so.head2);
    unsigned* sMisc = reinterpret_cast<unsigned*>(smem + so.misc);
    double* sLoss = reinterpret_cast<double*>(smem + so.loss);

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = blockIdx.x, kFinishCluster = gridDim.x;  // cluster dimension = (gridDim.x, 1, 1)
    const int b = blockIdx.y;
    // Register budget: 64 per thread, so that a 256-thread CTA (c1: 2048-point clouds) needs 2 warps x 2048 registers per SM
    // sub-partition -- exactly what ONE retiring sweep CTA (4 warps x 128 threads x 128 registers) frees there. The finish
    // CTAs then become resident, stage their clouds and start polling while the rest of the sweep is still running.
    const int nthreads = blockDim.x;
    // (the warp index through a shuffle: the compiler then treats it, and the branches on it, as warp-uniform)
    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0), nwarps = nthreads >> 5;
    constexpr bool worker = true;

    const int sg = a.seg.n == 1 ? 0 : seg_find(a.seg, b);
    const int bl = b - a.seg.b0[sg];
    const float* __restrict__ p1b = a.seg.p1[sg] + (size_t)bl * a.N * 3;
    const float* __restrict__ p2b = a.seg.p2[sg] + (size_t)bl * a.M * 3;
    float* const gp1b = a.seg.gp1[sg] ? a.seg.gp1[sg] + (size_t)bl * a.N * 3 : nullptr;
    float* const gp2b = (kWant2 && a.seg.gp2[sg]) ? a.seg.gp2[sg] + (size_t)bl * a.M * 3 : nullptr;

    cudaTriggerProgrammaticLaunchCompletion();  // the next kernel of the stream may start to launch

    // ---- 0. staging + bucket initialisation, ahead of the dependency wait ----
    __shared__ __align__(8) unsigned long long sStageBar;
    bool bulk;
    {
        const unsigned d2 = (unsigned)__cvta_generic_to_shared(sP2), d1 = (unsigned)__cvta_generic_to_shared(sP1);
        const int f2 = 3 * a.M, f1 = 3 * a.N;
        int v2 = (reinterpret_cast<uintptr_t>(p2b) & 15) == 0 ? f2 >> 2 : 0;
        int v1 = (reinterpret_cast<uintptr_t>(p1b) & 15) == 0 ? f1 >> 2 : 0;
        // Both clouds aligned and a whole number of 16-byte units (the usual case: 2048 points = 24576 bytes): ONE thread hands each
        // cloud to the bulk-copy engine (cp.async.bulk, completion on an mbarrier) instead of every thread issuing 16-byte copies
        // (the copy loops were 9 % of this kernel's instructions).
        bulk = v2 > 0 && v1 > 0 && (f2 & 3) == 0 && (f1 & 3) == 0;
        if (bulk) {
            if (tid == 0) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&sStageBar);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * (unsigned)(f2 + f1)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d2), "l"(p2b),
                             "r"(4u * (unsigned)f2), "r"(bar)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d1), "l"(p1b),
                             "r"(4u * (unsigned)f1), "r"(bar)
                             : "memory");
            }
        } else {
            for (int q = tid; q < v2; q += nthreads)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d2 + 16u * q), "l"(reinterpret_cast<const float4*>(p2b) + q) : "memory");
            for (int q = tid; q < v1; q += nthreads)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1 + 16u * q), "l"(reinterpret_cast<const float4*>(p1b) + q) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // what the 16-byte copies do not cover (unaligned cloud, last 1-3 floats), and +INF up to whole 32-point windows:
        // a candidate window may run past the end of the cloud, and a padded point never equals a finite minimum
        for (int q = 4 * v2 + tid; q < f2; q += nthreads) sP2[q] = __ldg(p2b + q);
        for (int q = f2 + tid; q < 3 * a.M32; q += nthreads) sP2[q] = CUDART_INF_F;
        for (int q = 4 * v1 + tid; q < f1; q += nthreads) sP1[q] = __ldg(p1b + q);
        for (int q = f1 + tid; q < 3 * a.N32; q += nthreads) sP1[q] = CUDART_INF_F;
        for (int q = tid; q < a.rpcN; q += nthreads) {
            sCnt1[q] = 0u;
            sHead1[q] = 0xffffu;
        }
        if (kWant2)
            for (int q = tid; q < a.rpcM; q += nthreads) {
                sCnt2[q] = 0u;
                sHead2[q] = 0xffffu;
            }
        if (tid < 4) sMisc[tid] = 0u;
    }
    cluster_arrive_release();  // barrier #1 (split): this CTA's counters are zero; waited for just before the pushes

    // this thread's row and column
    const int i = rank * a.rpcN + tid, j = rank * a.rpcM + tid;
    const bool rvalid = worker && tid < a.rpcN && i < a.N;
    const bool cvalid = worker && tid < a.rpcM && j < a.M;
    float ax = 0.f, ay = 0.f, az = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
    if (rvalid) {  // inputs: may be read ahead of the dependency wait
        ax = __ldg(p1b + 3 * i);
        ay = __ldg(p1b + 3 * i + 1);
        az = __ldg(p1b + 3 * i + 2);
    }
    if (cvalid) {
        qx = __ldg(p2b + 3 * j);
        qy = __ldg(p2b + 3 * j + 1);
        qz = __ldg(p2b + 3 * j + 2);
    }
    // upstream gradient of this pair's loss -> the two edge weights 2*g1, 2*g2 (g1 = gloss*w1/N, g2 = gloss*w2/M, rounded
    // as chamfer_bwd.cu rounds them)
    float gl;
    if (a.multi)
        gl = __fmul_rn(a.gtotal ? __ldg(a.gtotal) : 1.0f, a.seg.weight[sg]);
    else
        gl = a.gloss ? __ldg(a.gloss + b) : 1.0f;
    const float u1 = __fmul_rn(2.0f, __fdiv_rn(__fmul_rn(gl, a.w1), (float)a.N));
    const float u2 = __fmul_rn(2.0f, __fdiv_rn(__fmul_rn(gl, a.w2), (float)a.M));

    // Not cudaGridDependencySynchronize(): the CTA waits for the sweep CTAs that cover ITS pair only (their progress
    // words, zeroed before this grid could launch). Units are pair-major over the sweep CTAs and an SM retires its
    // resident sweep CTAs one after the other, so the early pairs are finished -- merge, arg-min, loss, backward -- while
    // the sweep is still working on the late ones, in the issue slots the sweep leaves free.
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 0);
    wait_pair_swept(a.progress, b, a.strips * a.ups, a.q, a.r);
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 1);

    // ---- 1. merges: the column records of strip 0 are put in flight first, then the row minima (24 per round trip), so
    // that both directions share one L2 latency ----
    const int TN = kSweepThreads * a.R;
    float mr = CUDART_INF_F, mc = CUDART_INF_F;
    int sub = 0, r0 = 0;
    u64 rec[kSweepWarps];
#pragma unroll
    for (int ww = 0; ww < kSweepWarps; ++ww)
        rec[ww] = cvalid ? __ldcg(a.colpart + (((size_t)b * a.strips) * kSweepWarps + ww) * a.Mpad + j) : ~0ull;
    if (rvalid) {
        const int strip = i / TN;
        merge_row_partials(a.rowmin, a.rowsub, b * a.strips + strip, i - strip * TN, TN, a.ups, a.q, a.r, a.maxslots, mr, sub);
    }
    if (cvalid) {
        for (int s = 0; s < a.strips; ++s) {  // ascending (strip, warp) = ascending rows, strict '<'
            if (s > 0) {
#pragma unroll
                for (int ww = 0; ww < kSweepWarps; ++ww) rec[ww] = __ldcg(a.colpart + (((size_t)b * a.strips + s) * kSweepWarps + ww) * a.Mpad + j);
            }
#pragma unroll
            for (int ww = 0; ww < kSweepWarps; ++ww) {
                const float v = __uint_as_float((unsigned)(rec[ww] >> 32));
                if (v < mc) {
                    mc = v;
                    r0 = (s * kSweepThreads + ww * 32 + ((unsigned)rec[ww] ? __ffs((unsigned)rec[ww]) - 1 : 0)) * a.R;
                }
            }
        }
    }

    // loss partials of this warp (fixed-order shuffle tree in double); pushed to cluster rank 0 in step 3
    double ls1 = rvalid ? (double)mr : 0.0, ls2 = cvalid ? (double)mc : 0.0;
    __syncwarp();  // tells the compiler the warp is converged here: the 20 shuffles below need no convergence wrappers (~16 instructions each)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ls1 += __shfl_down_sync(0xffffffffu, ls1, o);
        ls2 += __shfl_down_sync(0xffffffffu, ls2, o);
    }

    if (bulk) {  // CTA-uniform. The barrier was initialised by thread 0 ahead of the CTA barrier inside wait_pair_swept()
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&sStageBar);
        unsigned done = 0;
        for (long long spins = 0; !done && spins < (1ll << 26); ++spins)  // bounded: a lost transaction must not hang the GPU
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar)
                : "memory");
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // staged clouds visible to the whole CTA
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 2);

    // ---- 2. arg-min recovery: thread-private scan of the candidates' window, visited in a lane-dependent rotation
    // (windows start on multiples of 32 points = 96 floats: un-rotated all lanes would hit the same bank); the lowest
    // index whose distance equals the minimum bit-for-bit is the arg-min whatever the visiting order ----
    int x = 0, y = 0;  // idx1[i], idx2[j]
    if (rvalid) {
        const float* win = sP2 + 3 * (sub * kSub);
        int f = 0x7fffffff;
#pragma unroll 8
        for (int c = 0; c < kSub; ++c) {
            const int o = lane ^ c;
            const float d = sqdist(ax, ay, az, win[3 * o], win[3 * o + 1], win[3 * o + 2]);
            if (d == mr) f = min(f, o);
        }
        // no finite distance at all (NaN / Inf coordinates): index 0, like a strict-'<' scan that starts from +INF
        if (mr == CUDART_INF_F) x = 0;
        else if (f == 0x7fffffff) {
            atomicAdd(a.counters + a.B + 1, 1u);  // cannot happen; counted so that the tests can assert it
            x = min(sub * kSub, a.M - 1);
        } else x = sub * kSub + f;
        if (a.dist1) a.dist1[(size_t)b * a.N + i] = mr;
        if (a.idx1) a.idx1[(size_t)b * a.N + i] = x;
    }
    if (cvalid) {
        const float* win = sP1 + 3 * r0;
        const int R = a.R;
        int f = 0x7fffffff;
#pragma unroll 4
        for (int c = 0; c < R; ++c) {
            const int o = (lane ^ c) & (R - 1);
            const float d = sqdist(qx, qy, qz, win[3 * o], win[3 * o + 1], win[3 * o + 2]);
            if (d == mc) f = min(f, o);
        }
        if (mc == CUDART_INF_F) y = 0;
        else if (f == 0x7fffffff) {
            atomicAdd(a.counters + a.B + 1, 1u);
            y = min(r0, a.N - 1);
        } else y = r0 + f;
        if (a.dist2) a.dist2[(size_t)b * a.M + j] = mc;
        if (a.idx2) a.idx2[(size_t)b * a.M + j] = y;
    }

    // ---- 3. inverted index through distributed shared memory ----
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 3);
    cluster_wait_acquire();  // barrier #1: every CTA of the cluster is running and has zeroed its counters (long ago)
    if (lane == 0 && worker) {
        double* dst = cluster.map_shared_rank(sLoss, 0) + ((size_t)rank * nwarps + w) * 2;
        dst[0] = ls1;
        dst[1] = ls2;
    }
    if (cvalid && gp1b) {
        const int ro = y / a.rpcN, loc = y - ro * a.rpcN;
        const unsigned pos = atomicAdd(cluster.map_shared_rank(sCnt1, ro) + loc, 1u);
        if (pos < (unsigned)kBucket)
            cluster.map_shared_rank(sBkt1, ro)[loc * kBucket + pos] = (unsigned short)j;
        else {
            // a longer bucket continues as a chain through the owner's overflow list: entry = (previous head << 16) | j
            const unsigned o = atomicAdd(cluster.map_shared_rank(sMisc, ro), 1u);
            const unsigned prev = atomicExch(cluster.map_shared_rank(sHead1, ro) + loc, o);
            cluster.map_shared_rank(sOvf1, ro)[o] = (prev << 16) | (unsigned)j;
        }
    }
    if (kWant2 && rvalid && gp2b) {
        const int ro = x / a.rpcM, loc = x - ro * a.rpcM;
        const unsigned pos = atomicAdd(cluster.map_shared_rank(sCnt2, ro) + loc, 1u);
        if (pos < (unsigned)kBucket)
            cluster.map_shared_rank(sBkt2, ro)[loc * kBucket + pos] = (unsigned short)i;
        else {
            const unsigned o = atomicAdd(cluster.map_shared_rank(sMisc, ro) + 1, 1u);
            const unsigned prev = atomicExch(cluster.map_shared_rank(sHead2, ro) + loc, o);
            cluster.map_shared_rank(sOvf2, ro)[o] = (prev << 16) | (unsigned)i;
        }
    }
    cluster_arrive_release();  // barrier #2: buckets and loss partials of the whole pair are complete
    cluster_wait_acquire();
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 4);

    // ---- 4a. warp 0 of rank 0: loss of the pair (fixed order: independent of arrival order); for the batch-level sums it
    // only ISSUES its arrival here -- the atomic's round trip hides behind the backward, its result is looked at in 4c ----
    unsigned tot_ticket = 0u;
    if (w == 0 && rank == 0) {
        double s1 = 0.0, s2 = 0.0;
        const int nent = kFinishCluster * nwarps;
        for (int e = lane; e < nent; e += 32) {
            s1 += sLoss[2 * e];
            s2 += sLoss[2 * e + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            a.loss[b] = (float)((double)a.w1 * (s1 / a.N) + (double)a.w2 * (s2 / a.M));
            if (a.totals) tot_ticket = arrive_acq_rel(a.counters + a.B);  // publishes loss[b]
        }
    }

    // ---- 4b. backward: direct edge + bucket, gathered from the staged clouds, accumulated in double ----
    if (gp1b && tid < a.rpcN) {
        const unsigned c = rvalid ? sCnt1[tid] : 0u;
        double gx = 0.0, gy = 0.0, gz = 0.0;
        if (rvalid) {
            const float* d = sP2 + 3 * x;
            gx = (double)__fmul_rn(u1, __fsub_rn(ax, d[0]));
            gy = (double)__fmul_rn(u1, __fsub_rn(ay, d[1]));
            gz = (double)__fmul_rn(u1, __fsub_rn(az, d[2]));
            const uint2 bk = *reinterpret_cast<const uint2*>(sBkt1 + tid * kBucket);
            const unsigned e[4] = {bk.x & 0xffffu, bk.x >> 16, bk.y & 0xffffu, bk.y >> 16};
#pragma unroll
            for (int v = 0; v < kBucket; ++v) {
                if ((unsigned)v < c) {
                    const float* s = sP2 + 3 * e[v];
                    gx += (double)__fmul_rn(u2, __fsub_rn(ax, s[0]));
                    gy += (double)__fmul_rn(u2, __fsub_rn(ay, s[1]));
                    gz += (double)__fmul_rn(u2, __fsub_rn(az, s[2]));
                }
            }
        }
        if (c > (unsigned)kBucket) {  // a hub: the rest of its bucket is a chain in the CTA's overflow list (only this thread walks it;
                                      // scanning the whole list per thread was a fifth of the kernel's instructions on the bench's clouds)
            for (unsigned o = sHead1[tid]; o != 0xffffu;) {
                const unsigned v = sOvf1[o];
                const float* s = sP2 + 3 * (v & 0xffffu);
                gx += (double)__fmul_rn(u2, __fsub_rn(ax, s[0]));
                gy += (double)__fmul_rn(u2, __fsub_rn(ay, s[1]));
                gz += (double)__fmul_rn(u2, __fsub_rn(az, s[2]));
                o = v >> 16;
            }
        }
        if (rvalid) {
            float* o = gp1b + 3 * (size_t)i;
            o[0] = (float)gx;
            o[1] = (float)gy;
            o[2] = (float)gz;
        }
    }
    if (kWant2 && gp2b && tid < a.rpcM) {
        const unsigned c = cvalid ? sCnt2[tid] : 0u;
        double gx = 0.0, gy = 0.0, gz = 0.0;
        if (cvalid) {
            const float* d = sP1 + 3 * y;
            gx = (double)__fmul_rn(u2, __fsub_rn(qx, d[0]));
            gy = (double)__fmul_rn(u2, __fsub_rn(qy, d[1]));
            gz = (double)__fmul_rn(u2, __fsub_rn(qz, d[2]));
            const uint2 bk = *reinterpret_cast<const uint2*>(sBkt2 + tid * kBucket);
            const unsigned e[4] = {bk.x & 0xffffu, bk.x >> 16, bk.y & 0xffffu, bk.y >> 16};
#pragma unroll
            for (int v = 0; v < kBucket; ++v) {
                if ((unsigned)v < c) {
                    const float* s = sP1 + 3 * e[v];
                    gx += (double)__fmul_rn(u1, __fsub_rn(qx, s[0]));
                    gy += (double)__fmul_rn(u1, __fsub_rn(qy, s[1]));
                    gz += (double)__fmul_rn(u1, __fsub_rn(qz, s[2]));
                }
            }
        }
        if (c > (unsigned)kBucket) {
            for (unsigned o = sHead2[tid]; o != 0xffffu;) {
                const unsigned v = sOvf2[o];
                const float* s = sP1 + 3 * (v & 0xffffu);
                gx += (double)__fmul_rn(u1, __fsub_rn(qx, s[0]));
                gy += (double)__fmul_rn(u1, __fsub_rn(qy, s[1]));
                gz += (double)__fmul_rn(u1, __fsub_rn(qz, s[2]));
                o = v >> 16;
            }
        }
        if (cvalid) {
            float* o = gp2b + 3 * (size_t)j;
            o[0] = (float)gx;
            o[1] = (float)gy;
            o[2] = (float)gz;
        }
    }
    // ---- 4c. the reduction on top (few_shot.py:110,117,121): whichever pair arrived last adds the per-pair losses in index
    // order, per segment, in double -- independent of arrival order ----
    if (w == 0 && rank == 0 && a.totals) {
        if (__shfl_sync(0xffffffffu, tot_ticket, 0) == (unsigned)(a.B - 1)) {
            __syncwarp();
            const volatile float* lv = a.loss;
            double total = 0.0;
            for (int s = 0; s < a.seg.n; ++s) {
                double acc = 0.0;
                for (int q = a.seg.b0[s] + lane; q < a.seg.b0[s + 1]; q += 32) acc += (double)lv[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
                acc = __shfl_sync(0xffffffffu, acc, 0);
                if (lane == 0) a.totals[1 + s] = (float)acc;
                total += (double)a.seg.weight[s] * (double)(float)acc;
            }
            if (lane == 0) a.totals[0] = (float)total;
            peer_exchange(a.px, (float)total, lane);
        }
    }
    FPSG_TRACE(a.trace, 1, b * kFinishCluster + rank, 5);
}

bool chamfer_finish_supported(int N, int M) { return N <= kFinishMaxPoints && M <= kFinishMaxPoints; }

// Launches the finish kernel behind a sweep that used plan `pl` and workspace `ws`.
int chamfer_finish_launch(const SegTable& seg, const ChamferPlan& pl, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1,
                          int64_t* idx2, float* loss, float* totals, float w1, float w2, const float* gloss, const float* gtotal, bool multi,
                          void* ws, cudaStream_t stream, const PeerExchange* px) {
    bool want1 = false, want2 = false;
    for (int sg = 0; sg < seg.n; ++sg) {
        if (seg.b0[sg + 1] == seg.b0[sg]) continue;
        want1 |= seg.gp1[sg] != nullptr;
        want2 |= seg.gp2[sg] != nullptr;
    }
    (void)want1;
    char* base = static_cast<char*>(ws);
    FinishArgs fa;
    fa.seg = seg;
    fa.N = N, fa.M = M, fa.R = pl.R, fa.strips = pl.strips, fa.ups = pl.ups, fa.q = pl.q, fa.r = pl.r, fa.maxslots = pl.maxslots, fa.Mpad = pl.Mpad;
    fa.rowmin = reinterpret_cast<const float*>(base + pl.rowmin_off);
    fa.rowsub = reinterpret_cast<const int*>(base + pl.rowsub_off);
    fa.colpart = reinterpret_cast<const u64*>(base + pl.colpart_off);
    fa.dist1 = dist1, fa.dist2 = dist2, fa.idx1 = idx1, fa.idx2 = idx2;
    fa.counters = reinterpret_cast<unsigned*>(base + pl.counters_off);
    fa.loss = loss, fa.w1 = w1, fa.w2 = w2, fa.totals = totals, fa.gloss = gloss, fa.gtotal = gtotal, fa.multi = multi ? 1 : 0, fa.B = B;
    // CTAs per pair = cluster size: 8 CTAs of up to 512 threads. (Measured alternatives at c1, bench_micro/run_r02h.sh: 4 CTAs
    // of 512 threads 59.0 us, 4 CTAs alone on their SMs 59.3 us, 8 CTAs alone on their SMs 70.8 us, against 56.4 us. The
    // hardware places the CTAs of a cluster depth-first, so a finish CTA becomes resident only once its SM has two or three
    // free sweep slots -- 33-40 us into a 43-us sweep -- whatever its own footprint.)
    int cl = kFinishClusterMax;
    if (const char* e = tune_env("FPSG_FINISH_CL")) cl = atoi(e) == 4 ? 4 : 8;
    const int kFinishCluster = cl;
    fa.rpcN = (int)align_up((size_t)cdiv(N, kFinishCluster), 32);
    fa.rpcM = (int)align_up((size_t)cdiv(M, kFinishCluster), 32);
    fa.N32 = (int)align_up((size_t)N, 32);
    fa.M32 = (int)align_up((size_t)M, 32);
    fa.progress = reinterpret_cast<const unsigned*>(base + pl.progress_off);
    fa.trace = trace_buffer();
    if (px) fa.px = *px;
    else {
        fa.px = PeerExchange{};
        fa.px.world = 0;
    }
    const int threads = fa.rpcN > fa.rpcM ? fa.rpcN : fa.rpcM;
    const FinishSmem L = finish_smem_layout(N, M, fa.rpcN, fa.rpcM, want2, threads / 32);
    const FinishSmemOff so{(unsigned)L.p2, (unsigned)L.p1, (unsigned)L.cnt1, (unsigned)L.bkt1, (unsigned)L.ovf1, (unsigned)L.cnt2, (unsigned)L.bkt2,
                           (unsigned)L.ovf2, (unsigned)L.head1, (unsigned)L.head2, (unsigned)L.misc, (unsigned)L.loss};
    auto kern = want2 ? chamfer_finish_kernel<true> : chamfer_finish_kernel<false>;
    size_t smem_req = L.total;
    bool one_per_sm = false;
    if (const char* e = tune_env("FPSG_FINISH_PAD")) one_per_sm = atoi(e) != 0;  // tuning build only
    if (one_per_sm && smem_req < 116 * 1024) smem_req = 116 * 1024;  // more than half an SM's shared memory: one CTA per SM
    if (smem_req > 48 * 1024)  // per-device opt-in, repeated per launch like the other kernels' (about a microsecond of host time)
        FPSG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kFinishCluster, (unsigned)B);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem_req;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tune_env("FPSG_FIN_NOPDL") ? 0 : 1;  // (tuning build: plain stream order)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = kFinishCluster;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    FPSG_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, fa, so));
    return FPSG_OK;
}

// implemented in chamfer_fwd.cu / chamfer_bwd.cu
int chamfer_prepare(const SegTable& seg, int B, int N, int M, void* ws, size_t ws_bytes, ChamferPlan* out);
int chamfer_sweep_launch(const SegTable& seg, const ChamferPlan& pl, int B, int N, int M, void* ws, cudaStream_t stream);
int chamfer_finalize_launch(const SegTable& seg, const ChamferPlan& pl, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1,
                            int64_t* idx2, float* loss, float* totals, float w1, float w2, void* ws, cudaStream_t stream,
                            const PeerExchange* px = nullptr);
int chamfer_backward_launch(const SegTable& seg, const int64_t* idx1, const int64_t* idx2, const float* g1, const float* g2,
                            bool uniform, bool multi, const float* gtotal, float w1, float w2, int B, int N, int M, cudaStream_t stream);

// One loss step with a known upstream gradient: sweep, then the fused finish kernel (clouds up to kFinishMaxPoints
// points) or the finalize + backward kernels (larger clouds). mode (tuning builds): 1 forces the three-kernel chain.
int chamfer_fwd_bwd_launch_px(const SegTable& seg, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1, int64_t* idx2, float* loss,
                              float* totals, float w1, float w2, const float* gloss, const float* gtotal, bool multi, void* ws, size_t ws_bytes,
                              cudaStream_t stream, int mode, const PeerExchange* px) {
    if (B > 0 && (!idx1 || !idx2)) return fail(FPSG_EINVAL, "chamfer_fwd_bwd: null idx1/idx2");
    ChamferPlan pl;
    const int rc = chamfer_prepare(seg, B, N, M, ws, ws_bytes, &pl);
    if (rc != FPSG_OK || B == 0) return rc;
    const int rs = chamfer_sweep_launch(seg, pl, B, N, M, ws, stream);
    if (rs != FPSG_OK) return rs;
    static const bool no_fuse = tune_env("FPSG_NO_FINISH") != nullptr;  // tuning build only
    if (chamfer_finish_supported(N, M) && mode != 1 && !no_fuse)
        return chamfer_finish_launch(seg, pl, B, N, M, dist1, dist2, idx1, idx2, loss, totals, w1, w2, gloss, gtotal, multi, ws, stream, px);
    const int rf = chamfer_finalize_launch(seg, pl, B, N, M, dist1, dist2, idx1, idx2, loss, totals, w1, w2, ws, stream, px);
    if (rf != FPSG_OK) return rf;
    return chamfer_backward_launch(seg, idx1, idx2, gloss, gloss, true, multi, gtotal, w1, w2, B, N, M, stream);
}


int chamfer_fwd_bwd_launch(const SegTable& seg, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1, int64_t* idx2, float* loss,
                           float* totals, float w1, float w2, const float* gloss, const float* gtotal, bool multi, void* ws, size_t ws_bytes,
                           cudaStream_t stream, int mode) {
    return chamfer_fwd_bwd_launch_px(seg, B, N, M, dist1, dist2, idx1, idx2, loss, totals, w1, w2, gloss, gtotal, multi, ws, ws_bytes, stream, mode,
                                     nullptr);
}

}  // namespace fpsg
