// K3b: kNN graph for generic channel counts (DGCNN layers 2-4: C = 64, 64, 128) with a TENSOR-CORE PRE-FILTER, sm_100a.
//
// `knn(x, k)` (reference src/dgcnn/model.py:13-20) ranks, for every query point i, all candidates j by
//     pd_ij = rn( rn(2 * dot_ij - xx_j) - xx_i ),   dot = the (B,N,C) x (B,C,N) matmul of model.py:14.
// The ranking must be reproduced bit for bit (oracle/fpsg_oracle.c oracle_knn_f32: FP32 FMA chain over ascending c), which
// a tensor core cannot do -- but it can tell, at 30x the FP32 rate, which ~25 of the N candidates can possibly be among the
// k nearest. Three kernels, chained with programmatic dependent launch:
//
//   knn_tc_prep_kernel     x (B,C,N) channel-major -> xt (B*N, Kp) point-major ("K-major") rows
//                              [ x_0 .. x_{C-1} | 0.. | 1 1 1 0 0 0 0 0 | h m l xx 0 0 0 0 | 0.. ]
//                          where xx = sum_c rn(x_c^2) in ascending c (model.py:15) and h + m + l = -xx/2 exactly, each piece
//                          representable in TF32. The row serves as BOTH operands of the tensor-core product (the A side
//                          reads the `1 1 1` group where the B side reads `h m l`, so the accumulator receives
//                          s_ij = x_i . x_j - xx_j / 2, which ranks like pd_ij) AND as the contiguous FP32 row the exact
//                          re-scoring gathers (256 B per candidate at C = 64 instead of 64 strided sectors).
//   knn_tc_filter_kernel   one persistent CTA per SM; a tile = 128 queries of one cloud. TMA (cp.async.bulk.tensor, 128-byte
//                          swizzle) streams the cloud's rows through a shared-memory ring, ONE thread issues
//                          tcgen05.mma.kind::tf32 (M = 128, N = 256, K = 8 per instruction) into two 256-column TMEM
//                          accumulators, eight epilogue warps read them back with tcgen05.ld (a thread = half a row).
//                          Pass 1: maxima of <= 64 disjoint column groups per row; their k-th largest value T is reached by
//                          k different candidates. Pass 2 (the product is simply issued again -- it is nearly free): every
//                          candidate with s >= T - 2E is marked in a bit mask (N/32 words per query), E being a rigorous
//                          bound of |s_tensor - s_fp32| (TF32 operand truncation + accumulation): E = 2^-8 |x_i| max_j|x_j|
//                          + 2^-16 max_j|x_j|^2. By construction the marked set contains the exact top k.
//   knn_tc_rescore_kernel  a warp per query: the marked candidates (about k + 5) are re-scored with the reference's exact
//                          FP32 chain, gathered from xt, their 64-bit keys sorted by a warp bitonic network, the first k
//                          written. Any number of marked candidates is handled (massive ties: streaming sort-and-merge);
//                          fewer than k marked (non-finite inputs) -> every candidate is re-scored. The result never depends
//                          on the tensor-core arithmetic being exact, only on the bound E.
//
// Nothing of size N x N reaches memory (the FFMA path streams a 4 N^2-byte score matrix per cloud through a scratch).
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked

#include "common.cuh"
#include "knn_common.cuh"

namespace fpsg {

int sm_count();

constexpr int kTcThreads = 320;                  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int kTcEpiThreads = 256;
constexpr int kTcRows = 128;                     // queries per tile = UMMA M
constexpr int kTcBlk = 256;                      // candidates per accumulator buffer = UMMA N
constexpr int kTcChunk = 32;                     // floats per K chunk = one 128-byte swizzle row
constexpr int kTcChunkBytesA = kTcRows * 128;    // 16 KB
constexpr int kTcStageBytes = kTcBlk * 128;      // 32 KB
constexpr int kTcGroups = 32;                    // column groups per epilogue thread
constexpr int kTcMaxStages = 4;
constexpr int kTcSmemMax = 227 * 1024;
constexpr int kTcSmemFixed = kTcGroups * kTcEpiThreads * 4 + 256 + 1024;  // group maxima, barriers, alignment slack

struct KnnTcPlan {
    bool ok;
    int C8, Kp, NK, stages, W, nblk, gshift, tiles_per_cloud, ntiles;
    size_t smem, xx_off, xt_off, mask_off, ws_bytes;
};

static KnnTcPlan make_knn_tc_plan(int B, int C, int N, int k) {
    KnnTcPlan p = {};
    p.ok = false;
    if (k > 32 || (C & 7) != 0 || C < 16 || C > 240 || N < 256 || N > 8192) return p;
    if ((long long)B * N >= (1ll << 31) / 2) return p;
    if (tune_env("FPSG_KNN_NOTC")) return p;
    p.C8 = C;
    p.Kp = (int)align_up((size_t)p.C8 + 16, kTcChunk);
    p.NK = p.Kp / kTcChunk;
    p.stages = (kTcSmemMax - kTcSmemFixed - p.NK * kTcChunkBytesA) / kTcStageBytes;
    if (p.stages > kTcMaxStages) p.stages = kTcMaxStages;
    if (p.stages < 2) return p;
    p.W = cdiv(N, 32);
    p.nblk = cdiv(N, kTcBlk);
    p.gshift = 3;
    while (((p.nblk * (kTcBlk / 2)) >> p.gshift) > kTcGroups) ++p.gshift;
    if (cdiv(N, 1 << p.gshift) < k) return p;  // fewer than k non-empty groups: the threshold would not be a lower bound
    p.tiles_per_cloud = cdiv(N, kTcRows);
    p.ntiles = B * p.tiles_per_cloud;
    p.smem = (size_t)p.NK * kTcChunkBytesA + (size_t)p.stages * kTcStageBytes + kTcSmemFixed;
    size_t off = 0;
    p.xx_off = off;
    off += align_up((size_t)B * N * sizeof(float), 1024);
    p.xt_off = off;
    off += align_up((size_t)B * N * p.Kp * sizeof(float), 1024);
    p.mask_off = off;
    off += align_up((size_t)B * N * p.W * sizeof(unsigned), 1024);
    p.ws_bytes = off;
    p.ok = true;
    return p;
}

bool knn_tc_eligible(int B, int C, int N, int k) { return make_knn_tc_plan(B, C, N, k).ok; }
size_t knn_tc_workspace(int B, int C, int N, int k) { return make_knn_tc_plan(B, C, N, k).ws_bytes; }

// ---------------------------------------------------------------------------------------------------------
// prep: transpose + norms + operand rows
// ---------------------------------------------------------------------------------------------------------
constexpr int kPrepPts = 64;
__global__ void __launch_bounds__(256) knn_tc_prep_kernel(const float* __restrict__ x, int C, int N, int Kp, float* __restrict__ xx, float* __restrict__ xt) {
    extern __shared__ float sT[];  // [C][kPrepPts + 1]
    __shared__ float sXX[kPrepPts];
    constexpr int LD = kPrepPts + 1;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y, n0 = blockIdx.x * kPrepPts;
    cudaGridDependencySynchronize();  // x may come from the previous kernel of the stream
    cudaTriggerProgrammaticLaunchCompletion();
    const float* xb = x + (size_t)b * C * N;
    for (int c = w; c < C; c += 8) {
        const float* src = xb + (size_t)c * N + n0;
#pragma unroll
        for (int h = 0; h < kPrepPts / 32; ++h) {
            const int n = lane + 32 * h;
            sT[c * LD + n] = n0 + n < N ? __ldg(src + n) : 0.f;
        }
    }
    __syncthreads();
    if (tid < kPrepPts) {  // xx = sum_c rn(x_c^2), ascending c (model.py:15, oracle_knn_f32)
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = sT[c * LD + tid];
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
        sXX[tid] = s;
        if (n0 + tid < N) xx[(size_t)b * N + n0 + tid] = s;
    }
    __syncthreads();
    for (int r = w; r < kPrepPts; r += 8) {
        if (n0 + r >= N) break;
        float* dst = xt + ((size_t)b * N + n0 + r) * Kp;
        const float s = sXX[r];
        // -xx/2 = h + m + l exactly, every piece with at most 11 significant bits (TF32 keeps 1 + 10)
        const float v = -0.5f * s;
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float v1 = __fsub_rn(v, h);
        const float m = __uint_as_float(__float_as_uint(v1) & 0xffffe000u);
        const float l = __fsub_rn(v1, m);
        for (int cc = lane; cc < Kp; cc += 32) {
            float o = 0.f;
            if (cc < C) o = sT[cc * LD + r];
            else if (cc < C + 3) o = 1.0f;                       // the A-side group: 1 1 1 0 0 0 0 0
            else if (cc == C + 8) o = h;                         // the B-side group: h m l xx 0 0 0 0
            else if (cc == C + 9) o = m;
            else if (cc == C + 10) o = l;
            else if (cc == C + 11) o = s;                        // multiplied by the A side's 0: carried for the re-scoring
            dst[cc] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that cannot hang the GPU: a pipeline that never completes (a malformed descriptor, a lost transaction) traps after
// about two seconds instead of spinning for ever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tmap),
                 "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 operands, FP32 accumulation; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor of a K-major operand tile in the 128-byte swizzle: rows of 128 bytes, groups of 8 rows
// 1024 bytes apart (SBO); bits 46-47 = 1 (sm_100 descriptor version), bits 61-63 = 2 (SWIZZLE_128B). A K step of 8 TF32
// values advances the start address by 32 bytes inside the swizzle row.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcArgs {
    const float* xx;   // [B*N]
    unsigned* masks;   // [B*N][W]
    int N, k, C8, NK, stages, W, nblk, gshift, tiles_per_cloud, ntiles;
    float* dump;       // tuning builds: [B*N][N] raw accumulator values of pass 1, else nullptr
};

// ---------------------------------------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                                               // [NK] chunks of 128 rows x 128 B (swizzled)
    unsigned char* sB = sA + (size_t)a.NK * kTcChunkBytesA;                 // [stages] 256 rows x 128 B
    float* sGm = reinterpret_cast<float*>(sB + (size_t)a.stages * kTcStageBytes);  // [32 groups][256 epilogue threads]
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sGm + kTcGroups * kTcEpiThreads);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 16);
    float* sRed = reinterpret_cast<float*>(sTmem + 4);  // [8]

    const uint32_t bar_afull = smem_u32(sBar + 0), bar_aempty = smem_u32(sBar + 1);
    const uint32_t bar_full = smem_u32(sBar + 2), bar_empty = smem_u32(sBar + 2 + kTcMaxStages);    // [stages] each
    const uint32_t bar_tfull = smem_u32(sBar + 2 + 2 * kTcMaxStages), bar_tempty = bar_tfull + 16;  // [2] each

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        for (int s = 0; s < kTcMaxStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, kTcEpiThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the whole TMEM: two accumulator buffers of 256 columns (one CTA per SM: nobody else allocates)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;

    cudaGridDependencySynchronize();            // xt and xx (prep kernel) are complete
    cudaTriggerProgrammaticLaunchCompletion();  // the re-scoring grid may be scheduled as SMs free up; it waits for this grid itself

    const int NK = a.NK, stages = a.stages, nblk = a.nblk, N = a.N;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++ti) {
                const int b = tile / a.tiles_per_cloud, i0 = (tile - b * a.tiles_per_cloud) * kTcRows;
                const int row0 = b * N;
                mbar_wait(bar_aempty, (ti & 1) ^ 1);  // every MMA that read the previous tile's queries has completed
                mbar_expect_tx(bar_afull, (uint32_t)NK * kTcChunkBytesA);
                for (int kc = 0; kc < NK; ++kc) tma_load_2d(smem_u32(sA + (size_t)kc * kTcChunkBytesA), &tmap, bar_afull, kc * kTcChunk, row0 + i0);
                for (int pb = 0; pb < 2 * nblk; ++pb) {
                    const int blk = pb < nblk ? pb : pb - nblk;
                    for (int kc = 0; kc < NK; ++kc, ++it) {
                        const uint32_t s = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                        mbar_expect_tx(bar_full + 8 * s, kTcStageBytes);
                        const uint32_t dst = smem_u32(sB + (size_t)s * kTcStageBytes);
                        tma_load_2d(dst, &tmap, bar_full + 8 * s, kc * kTcChunk, row0 + blk * kTcBlk);
                        tma_load_2d(dst + kTcStageBytes / 2, &tmap, bar_full + 8 * s, kc * kTcChunk, row0 + blk * kTcBlk + kTcBlk / 2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const int ca = a.C8 / kTcChunk, offA = (a.C8 % kTcChunk) * 4;              // the `1 1 1` group of the A rows
            const int cb = (a.C8 + 8) / kTcChunk, offB = ((a.C8 + 8) % kTcChunk) * 4;  // the `h m l` group of the B rows
            const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
            uint32_t it = 0, tb = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++ti) {
                mbar_wait(bar_afull, ti & 1);
                tc_fence_after();
                for (int pb = 0; pb < 2 * nblk; ++pb, ++tb) {
                    const uint32_t buf = tb & 1u;
                    mbar_wait(bar_tempty + 8 * buf, ((tb >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem + buf * kTcBlk;
                    uint32_t acc = 0;
                    for (int kc = 0; kc < NK; ++kc, ++it) {
                        const uint32_t s = it % (uint32_t)stages;
                        mbar_wait(bar_full + 8 * s, (it / (uint32_t)stages) & 1u);
                        tc_fence_after();
                        const uint32_t bbase = sB_addr + s * kTcStageBytes, abase = sA_addr + kc * kTcChunkBytesA;
                        const int nreg = min(4, max(0, (a.C8 - kc * kTcChunk) / 8));
                        for (int s8 = 0; s8 < nreg; ++s8) {
                            umma_tf32(d, umma_desc(abase + s8 * 32), umma_desc(bbase + s8 * 32), kTcIdesc, acc);
                            acc = 1;
                        }
                        if (kc == cb) {
                            umma_tf32(d, umma_desc(sA_addr + ca * kTcChunkBytesA + offA), umma_desc(bbase + offB), kTcIdesc, acc);
                            acc = 1;
                        }
                        umma_commit(bar_empty + 8 * s);  // the stage is free once these MMAs have read it
                    }
                    umma_commit(bar_tfull + 8 * buf);    // accumulator complete
                }
                umma_commit(bar_aempty);
            }
        }
    } else {
        // ===== epilogue: 8 warps, a thread = one query row x one half of every 256-column block =====
        const int te = threadIdx.x - 64, e = te >> 5;
        const int quad = warp & 3, half = e >> 2;  // TMEM lanes 32 * (warp % 4) .. + 31 are the ones this warp may read
        const int rowt = quad * 32 + lane;
        float* gm = sGm + te;
        const float* gmp = sGm + (te ^ 128);  // the thread that owns the other half of this row
        const int GS = 1 << a.gshift;
        uint32_t tb = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int b = tile / a.tiles_per_cloud, i0 = (tile - b * a.tiles_per_cloud) * kTcRows;
            const int i = i0 + rowt;
            const bool row_ok = i < N;
            const float* xxb = a.xx + (size_t)b * N;
            // the error bound needs max_j |x_j|^2 of the cloud
            float mx = 0.f;
            for (int j = te; j < N; j += kTcEpiThreads) mx = fmaxf(mx, __ldg(xxb + j));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            asm volatile("bar.sync 1, 256;" ::: "memory");  // everybody is done with the previous tile's shared values
            if (lane == 0) sRed[e] = mx;
#pragma unroll
            for (int g = 0; g < kTcGroups; ++g) gm[g * kTcEpiThreads] = -CUDART_INF_F;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3])), fmaxf(fmaxf(sRed[4], sRed[5]), fmaxf(sRed[6], sRed[7])));
            const float xxi = __ldg(xxb + min(i, N - 1));

            // ---- pass 1: group maxima ----
            for (int blk = 0; blk < nblk; ++blk, ++tb) {
                const uint32_t buf = tb & 1u;
                mbar_wait(bar_tfull + 8 * buf, (tb >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const int j0 = blk * kTcBlk + half * (kTcBlk / 2) + q * 32;
                    if (j0 >= N) break;  // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + buf * kTcBlk + half * (kTcBlk / 2) + q * 32, v);
                    if (j0 + 32 > N) {
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            if (j0 + r >= N) v[r] = 0xff800000u;  // -INF: a column past the cloud never counts
                    }
#ifdef FPSG_TUNING
                    if (a.dump != nullptr && row_ok) {
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            if (j0 + r < N) a.dump[((size_t)b * N + i) * N + j0 + r] = __uint_as_float(v[r]);
                    }
#endif
                    float m8[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float* f = reinterpret_cast<const float*>(v) + 8 * u;
                        m8[u] = fmaxf(max3(max3(f[0], f[1], f[2]), f[3], f[4]), max3(f[5], f[6], f[7]));
                    }
                    const int lc0 = blk * (kTcBlk / 2) + q * 32;  // index of v[0] among the columns of THIS thread
                    const int g = lc0 >> a.gshift;
                    if (GS == 8) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) gm[(g + u) * kTcEpiThreads] = m8[u];
                    } else if (GS == 16) {
                        gm[g * kTcEpiThreads] = fmaxf(m8[0], m8[1]);
                        gm[(g + 1) * kTcEpiThreads] = fmaxf(m8[2], m8[3]);
                    } else {
                        gm[g * kTcEpiThreads] = max3(max3(m8[0], m8[1], m8[2]), m8[3], gm[g * kTcEpiThreads]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            }

            // ---- threshold: k-th largest of the row's <= 64 group maxima (32 per thread, sorted in registers, merged) ----
            {
                float s_[kTcGroups];
#pragma unroll
                for (int c = 0; c < kTcGroups; ++c) s_[c] = gm[c * kTcEpiThreads];
#pragma unroll
                for (int s = 2; s <= kTcGroups; s <<= 1) {
#pragma unroll
                    for (int d = s >> 1; d > 0; d >>= 1) {
#pragma unroll
                        for (int el = 0; el < kTcGroups; ++el) {
                            const int p = el ^ d;
                            if (p > el) {
                                const bool desc = (s == kTcGroups) || ((el & s) == 0);
                                const float hi = fmaxf(s_[el], s_[p]), lo = fminf(s_[el], s_[p]);
                                s_[el] = desc ? hi : lo;
                                s_[p] = desc ? lo : hi;
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < kTcGroups; ++c) gm[c * kTcEpiThreads] = s_[c];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float T = -CUDART_INF_F;
            {
                int ia = 0, ib = 0;
                for (int t = 0; t < a.k; ++t) {  // ia + ib = t <= 31: both indices stay inside their arrays
                    const float va = gm[ia * kTcEpiThreads], vb = gmp[ib * kTcEpiThreads];
                    if (va >= vb) {
                        T = va;
                        ++ia;
                    } else {
                        T = vb;
                        ++ib;
                    }
                }
            }
            // |s_tensor - s_fp32| <= E for every candidate of this row: TF32 keeps 1 + 10 bits of each operand (relative 2^-10 each,
            // 2^-9 per product, Cauchy-Schwarz over the channels), the FP32 chain and the accumulator add O(C 2^-23) of the same
            // magnitude, the h m l pieces are exact; a factor 2 on top
            const float E = 0.00390625f * sqrtf(xxi * mx) + 1.52587890625e-5f * mx;
            const float thr = T - 2.0f * E;

            // ---- pass 2: mark every candidate that can still be among the k nearest ----
            for (int blk = 0; blk < nblk; ++blk, ++tb) {
                const uint32_t buf = tb & 1u;
                mbar_wait(bar_tfull + 8 * buf, (tb >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const int j0 = blk * kTcBlk + half * (kTcBlk / 2) + q * 32;
                    if (j0 >= N) break;
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + buf * kTcBlk + half * (kTcBlk / 2) + q * 32, v);
                    unsigned m = 0u;
#pragma unroll
                    for (int r = 0; r < 32; ++r) m |= __uint_as_float(v[r]) >= thr ? (1u << r) : 0u;
                    if (j0 + 32 > N) m &= (1u << (N - j0)) - 1u;
                    if (row_ok) a.masks[((size_t)b * N + i) * a.W + (j0 >> 5)] = m;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// exact re-scoring of the marked candidates
// ---------------------------------------------------------------------------------------------------------
constexpr int kRsWarps = 8;
constexpr int kRsCap = 128;
constexpr int kRsWords = 8;  // mask words per lane: N <= 8192

__global__ void __launch_bounds__(kRsWarps * 32)
knn_tc_rescore_kernel(const float* __restrict__ xt, const unsigned* __restrict__ masks, long long rows, int N, int k, int C, int Kp, int W,
                      int64_t* __restrict__ idx) {
    __shared__ unsigned short sList[kRsWarps][kRsCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    cudaGridDependencySynchronize();  // the filter's masks are complete
    const long long row = (long long)blockIdx.x * kRsWarps + w;
    if (row >= rows) return;  // warp-uniform; no CTA-wide barrier below
    const long long cloud0 = row - row % N;  // first row of this query's cloud
    const float* __restrict__ qrow = xt + (size_t)row * Kp;
    const float xxi = __ldg(qrow + C + 11);

    // pd of candidate j, exactly as the reference rounds it: dot = FMA chain over ascending c from 0 (model.py:14),
    // pd = rn(rn(2 dot - xx_j) - xx_i) (model.py:16)
    auto key_of = [&](int j) -> u64 {
        const float4* __restrict__ cr = reinterpret_cast<const float4*>(xt + (size_t)(cloud0 + j) * Kp);
        const float4* __restrict__ qr = reinterpret_cast<const float4*>(qrow);
        float dot = 0.f;
#pragma unroll 4
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 v = __ldg(cr + c4), q = __ldg(qr + c4);
            dot = __fmaf_rn(q.x, v.x, dot);
            dot = __fmaf_rn(q.y, v.y, dot);
            dot = __fmaf_rn(q.z, v.z, dot);
            dot = __fmaf_rn(q.w, v.w, dot);
        }
        const float xxj = __ldg(reinterpret_cast<const float*>(cr) + C + 11);
        return knn_key_of(__fsub_rn(__fmaf_rn(2.0f, dot, -xxj), xxi), j);
    };

    unsigned mw[kRsWords];
    int mine = 0;
#pragma unroll
    for (int s = 0; s < kRsWords; ++s) {
        const int wi = s * 32 + lane;
        mw[s] = wi < W ? __ldcg(masks + (size_t)row * W + wi) : 0u;
        mine += __popc(mw[s]);
    }
    int total = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (total < k) {  // cannot happen for finite inputs; non-finite scores compare false everywhere: take every candidate
        mine = 0;
#pragma unroll
        for (int s = 0; s < kRsWords; ++s) {
            const int j0 = (s * 32 + lane) * 32;
            mw[s] = j0 >= N ? 0u : (j0 + 32 <= N ? 0xffffffffu : (1u << (N - j0)) - 1u);
            mine += __popc(mw[s]);
        }
        total = N;
    }

    u64 key = 0ull;
    if (total <= kRsCap) {
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        int pos = incl - mine;
#pragma unroll
        for (int s = 0; s < kRsWords; ++s) {
            unsigned m = mw[s];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                sList[w][pos++] = (unsigned short)((s * 32 + lane) * 32 + bit);
            }
        }
        __syncwarp();
        for (int h = 0; h * 32 < total; ++h) {
            const int el = h * 32 + lane;
            u64 kb = el < total ? key_of((int)sList[w][el]) : 0ull;
            kb = warp_sort_desc(kb, lane);
            if (h == 0) {
                key = kb;
            } else {
                const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                key = warp_merge_desc(key > rev ? key : rev, lane);
            }
        }
    } else {
        // massive ties (or the take-everything case): streaming sort-and-merge, one mask word = 32 candidates per round
        u64 tau = 0ull;
#pragma unroll
        for (int s = 0; s < kRsWords; ++s) {
            for (int l = 0; l < 32; ++l) {
                const unsigned word = __shfl_sync(0xffffffffu, mw[s], l);
                if (word == 0u) continue;  // warp-uniform
                u64 kb = (word >> lane) & 1u ? key_of((s * 32 + l) * 32 + lane) : 0ull;
                if (!(kb > tau)) kb = 0ull;
                if (__any_sync(0xffffffffu, kb != 0ull)) {
                    kb = warp_sort_desc(kb, lane);
                    const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                    key = warp_merge_desc(key > rev ? key : rev, lane);
                    tau = __shfl_sync(0xffffffffu, key, k - 1);
                }
            }
        }
    }
    if (lane < k) idx[(size_t)row * k + lane] = (int64_t)(0xffffffffu - (unsigned)key);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static int knn_tc_run(const float* x, int B, int C, int N, int k, int64_t* idx, void* ws, float* dump, bool rescoring, cudaStream_t stream) {
    const KnnTcPlan p = make_knn_tc_plan(B, C, N, k);
    if (!p.ok) return fail(FPSG_EINVAL, "knn (tensor-core path): shape B=%d C=%d N=%d k=%d not eligible", B, C, N, k);
    char* wsb = static_cast<char*>(ws);
    float* xx = reinterpret_cast<float*>(wsb + p.xx_off);
    float* xt = reinterpret_cast<float*>(wsb + p.xt_off);
    unsigned* masks = reinterpret_cast<unsigned*>(wsb + p.mask_off);

    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(FPSG_ECUDA, "knn: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)p.Kp, (cuuint64_t)B * (cuuint64_t)N};
    const cuuint64_t gstr[1] = {(cuuint64_t)p.Kp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcChunk, (cuuint32_t)kTcRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, xt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(FPSG_ECUDA, "knn: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const size_t prep_smem = (size_t)C * (kPrepPts + 1) * sizeof(float);
    if (prep_smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_tc_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    FPSG_CUDA_OK(launch_pdl(knn_tc_prep_kernel, dim3((unsigned)cdiv(N, kPrepPts), (unsigned)B), dim3(256), prep_smem, stream, x, C, N, p.Kp, xx, xt));

    TcArgs a;
    a.xx = xx, a.masks = masks, a.N = N, a.k = k, a.C8 = p.C8, a.NK = p.NK, a.stages = p.stages, a.W = p.W, a.nblk = p.nblk, a.gshift = p.gshift;
    a.tiles_per_cloud = p.tiles_per_cloud, a.ntiles = p.ntiles, a.dump = dump;
    FPSG_CUDA_OK(cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const int grid = min(p.ntiles, sm_count());
    FPSG_CUDA_OK(launch_pdl(knn_tc_filter_kernel, dim3((unsigned)grid), dim3(kTcThreads), p.smem, stream, tmap, a));
    if (!rescoring) return FPSG_OK;

    const long long rows = (long long)B * N;
    FPSG_CUDA_OK(launch_pdl(knn_tc_rescore_kernel, dim3((unsigned)((rows + kRsWarps - 1) / kRsWarps)), dim3(kRsWarps * 32), 0, stream,
                            (const float*)xt, (const unsigned*)masks, rows, N, k, C, p.Kp, p.W, idx));
    return FPSG_OK;
}

int knn_tc_launch(const float* x, int B, int C, int N, int k, int64_t* idx, void* ws, cudaStream_t stream) {
    return knn_tc_run(x, B, C, N, k, idx, ws, nullptr, true, stream);
}

}  // namespace fpsg

#ifdef FPSG_TUNING
// Development aid (tuning build only): the raw tensor-core scores s_ij = x_i . x_j - xx_j / 2 of pass 1, (B,N,N) floats,
// and the candidate masks left in the workspace -- to check descriptors and error bounds against numpy.
extern "C" int fpsg_tune_knn_tc_scores(const float* x, int B, int C, int N, int k, float* scores, void* ws, size_t ws_bytes, void* stream) {
    using namespace fpsg;
    if (!knn_tc_eligible(B, C, N, k)) return fail(FPSG_EINVAL, "knn_tc_scores: shape not eligible");
    if (!ws || ws_bytes < knn_tc_workspace(B, C, N, k)) return fail(FPSG_EWORKSPACE, "knn_tc_scores: workspace of %zu bytes required", knn_tc_workspace(B, C, N, k));
    return knn_tc_run(x, B, C, N, k, nullptr, ws, scores, false, static_cast<cudaStream_t>(stream));
}
extern "C" size_t fpsg_tune_knn_tc_mask_offset(int B, int C, int N, int k) { return fpsg::make_knn_tc_plan(B, C, N, k).mask_off; }
#endif
