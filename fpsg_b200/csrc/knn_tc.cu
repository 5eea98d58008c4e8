// K3b: kNN graph for generic channel counts (DGCNN layers 2-4: C = 64, 64, 128) with a TENSOR-CORE PRE-FILTER, sm_100a.
//
// `knn(x, k)` (reference src/dgcnn/model.py:13-20) ranks, for every query point i, all candidates j by
//     pd_ij = rn( rn(2 * dot_ij - xx_j) - xx_i ),   dot = the (B,N,C) x (B,C,N) matmul of model.py:14.
// The ranking must be reproduced bit for bit (oracle/fpsg_oracle.c oracle_knn_f32: FP32 FMA chain over ascending c), which
// a tensor core cannot do -- but it can tell, at 30x the FP32 rate, which ~25 of the N candidates can possibly be among the
// k nearest. Four kernels, chained with programmatic dependent launch:
//
//   knn_tc_prep_kernel     x (B,C,N) channel-major -> xt (B*N, C) point-major FP32 rows (what the exact re-scoring gathers:
//                          256 contiguous bytes per candidate at C = 64 instead of 64 strided sectors), the squared norms
//                          xx = sum_c rn(x_c^2) in ascending c (model.py:15), and the largest norm of every cloud.
//   knn_tc_operand_kernel  the FP16 operand rows (B*N, C64 + 32): [ fp16(s x_c) | 0.. | 1 1 1 0.. | h m l 0.. ]. s is a power
//                          of two per cloud that brings its largest norm into [2^6, 2^7): FP16's 11 significant bits are kept
//                          whatever the magnitude of the features, and scaling by powers of two is exact. h + m + l =
//                          -s^2 xx / 2 exactly. A row serves as BOTH operands of the product (the A side reads the `1 1 1`
//                          group where the B side reads `h m l`), so the accumulator receives s^2 (x_i . x_j - xx_j / 2), which
//                          ranks like pd_ij.
//   knn_tc_filter_kernel   one persistent CTA per SM; a tile = 128 queries of one cloud. TMA (cp.async.bulk.tensor; 128-byte
//                          swizzle for the 64-channel chunks, 32-byte swizzle for the 16-value groups) streams the cloud
//                          through a shared-memory ring, one stage = one block of 256 candidates with all its K chunks; ONE
//                          thread issues tcgen05.mma.kind::f16 (M = 128, N = 256, K = 16 per instruction) into two 256-column
//                          TMEM accumulators; sixteen epilogue warps read them back with tcgen05.ld (a thread = a quarter of
//                          a row). Pass 1: maxima of <= 64 disjoint column groups per row; their k-th largest value T is
//                          reached by k different candidates. Pass 2 (the product is simply issued again -- it is nearly
//                          free): every candidate with score >= T - 2E is marked in a bit mask (N/32 words per query), E being a
//                          rigorous bound of |s_tensor - s_fp32| (operand rounding + accumulation):
//                          E = 1.25 x 2^-10 |x_i| max_j|x_j| + 2^-16 max_j|x_j|^2. By construction the marked set contains
//                          the exact top k.
//   knn_tc_rescore_kernel  a warp per query: the marked candidates (about k + 5) are re-scored with the reference's exact
//                          FP32 chain, gathered from xt, their 64-bit keys sorted by a warp bitonic network, the first k
//                          written. Any number of marked candidates is handled (massive ties: streaming sort-and-merge);
//                          fewer than k marked (non-finite inputs) -> every candidate is re-scored. The result never depends
//                          on the tensor-core arithmetic being exact, only on the bound E.
//
// Nothing of size N x N reaches memory (the FFMA path streams a 4 N^2-byte score matrix per cloud through a scratch).
// Measured dead ends, kept for the record (profiles/r02_knn_tc.md): TF32 operands straight from FP32 rows (three 32-KB
// stages per block: the MMA-issuing thread and the TMA feed serialise, 37 us for the filter at c3 / C = 64 instead of 25);
// BF16 operands (twice the error band: 29 instead of 25 marked candidates, every third query needs a second re-scoring round).
#include <cuda_fp16.h>
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked

#include "common.cuh"
#include "knn_common.cuh"

namespace fpsg {

int sm_count();

constexpr int kTcThreads = 576;                  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-17: epilogue
constexpr int kTcEpiThreads = 512;
constexpr int kTcGmLd = 256;                     // row halves per tile: columns of the group-maxima array
constexpr int kTcRows = 128;                     // queries per tile = UMMA M
constexpr int kTcBlk = 256;                      // candidates per accumulator buffer = UMMA N
constexpr int kTcChunk = 64;                     // FP16 channels per K chunk = one 128-byte swizzle row
constexpr int kTcAux = 16;                       // FP16 values per auxiliary group (one K step): `1 1 1 0..` and `h m l 0..`
constexpr int kTcChunkBytesA = kTcRows * 128;    // 16 KB
constexpr int kTcChunkBytesB = kTcBlk * 128;     // 32 KB
constexpr int kTcAuxBytesA = kTcRows * 32;       // 4 KB
constexpr int kTcAuxBytesB = kTcBlk * 32;        // 8 KB
constexpr int kTcGroups = 32;                    // column groups per epilogue thread
constexpr int kTcMaxStages = 4;
constexpr int kTcSmemMax = 227 * 1024;
constexpr int kTcSmemFixed = kTcGroups * kTcGmLd * 4 + kTcGmLd * 4 + 256 + 1024;  // group maxima, thresholds, barriers, alignment slack

struct KnnTcPlan {
    bool ok;
    int C64, Kp16, NK, stages, W, nblk, gshift, tiles_per_cloud, ntiles;
    size_t a_bytes, stage_bytes;  // shared memory of the query tile / of one candidate block (all K chunks + the auxiliary group)
    size_t smem, cmax_off, xx_off, xt_off, xb_off, mask_off, ws_bytes;
};

static KnnTcPlan make_knn_tc_plan(int B, int C, int N, int k) {
    KnnTcPlan p = {};
    p.ok = false;
    if (k > 32 || (C & 7) != 0 || C < 16 || C > 128 || N < 256 || N > 4096) return p;  // N <= 4096: a group (<= 64 columns) never spans two threads
    if ((long long)B * N >= (1ll << 31) / 2) return p;
    if (tune_env("FPSG_KNN_NOTC")) return p;
    p.C64 = (int)align_up((size_t)C, kTcChunk);
    p.Kp16 = p.C64 + 2 * kTcAux;  // FP16 operand row: channels (zero-padded to whole chunks), the A-side group, the B-side group
    p.NK = p.C64 / kTcChunk;
    p.a_bytes = (size_t)p.NK * kTcChunkBytesA + kTcAuxBytesA;
    p.stage_bytes = (size_t)p.NK * kTcChunkBytesB + kTcAuxBytesB;
    p.stages = (int)((kTcSmemMax - kTcSmemFixed - p.a_bytes) / p.stage_bytes);
    if (p.stages > kTcMaxStages) p.stages = kTcMaxStages;
    if (p.stages < 2) return p;
    p.W = cdiv(N, 32);
    p.nblk = cdiv(N, kTcBlk);
    p.gshift = 3;
    while (((p.nblk * (kTcBlk / 2)) >> p.gshift) > kTcGroups) ++p.gshift;
    if (cdiv(N, 1 << p.gshift) < k) return p;  // fewer than k non-empty groups: the threshold would not be a lower bound
    p.tiles_per_cloud = cdiv(N, kTcRows);
    p.ntiles = B * p.tiles_per_cloud;
    p.smem = p.a_bytes + (size_t)p.stages * p.stage_bytes + kTcSmemFixed;
    size_t off = 0;
    p.cmax_off = off;
    off += align_up((size_t)B * sizeof(unsigned), 1024);
    p.xx_off = off;
    off += align_up((size_t)B * N * sizeof(float), 1024);
    p.xt_off = off;
    off += align_up((size_t)B * N * C * sizeof(float), 1024);
    p.xb_off = off;
    off += align_up((size_t)B * N * p.Kp16 * sizeof(unsigned short), 1024);
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
constexpr int kPrepPts = 64;       // points per CTA of pass 1
constexpr int kPrepThreads = 256;

// The FP16 operands are the cloud scaled by a power of two s chosen from its largest norm, max_j |x_j| s in [2^6, 2^7): every
// value fits FP16 comfortably, whatever the magnitude of the features, and the tensor-core scores are s^2 times the unscaled
// ones (exact scaling: the threshold logic is scale-invariant). cmax = bits of max_j xx_j (non-negative floats order like
// their bit patterns).
__device__ __forceinline__ float knn_tc_scale(unsigned cmax_bits) {
    const float m = __uint_as_float(cmax_bits);
    if (!(m > 0.f) || !(m < CUDART_INF_F)) return 1.0f;  // all-zero cloud, NaN / Inf: nothing sensible to scale
    int e;
    frexpf(sqrtf(m), &e);  // sqrt(m) = f 2^e, f in [0.5, 1)
    return ldexpf(1.0f, 7 - e);
}

// pass 1: norms, the FP32 point-major rows for the exact re-scoring, the cloud's largest squared norm
__global__ void __launch_bounds__(kPrepThreads)
knn_tc_prep_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ xx, float* __restrict__ xt, unsigned* __restrict__ cmax) {
    extern __shared__ float sT[];  // [C][kPrepPts + 1]
    constexpr int LD = kPrepPts + 1;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y, n0 = blockIdx.x * kPrepPts;
    cudaGridDependencySynchronize();  // x may come from the previous kernel of the stream
    cudaTriggerProgrammaticLaunchCompletion();
    const float* xb_ = x + (size_t)b * C * N;
    constexpr int kW = kPrepThreads / 32;
    for (int c = w; c < C; c += kW) {
        const float* src = xb_ + (size_t)c * N + n0;
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
        if (n0 + tid < N) xx[(size_t)b * N + n0 + tid] = s;
        // padded points contribute 0; a NaN norm is ignored by fmaxf (non-finite inputs: the re-scoring's safety net applies)
        float m = fmaxf(s, 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(cmax + b, __float_as_uint(m));
    }
    __syncthreads();
    for (int r = w; r < kPrepPts; r += kW) {
        if (n0 + r >= N) break;
        float* dst = xt + ((size_t)b * N + n0 + r) * C;
        for (int cc = lane; cc < C; cc += 32) dst[cc] = sT[cc * LD + r];
    }
}

// pass 2: the FP16 operand rows [ fp16(s x_c) | 0.. | 1 1 1 0.. | h m l 0.. ], h + m + l = -s^2 xx / 2 (11 significant bits each).
// A lane per four values (16-byte reads of the FP32 rows, 8-byte writes), a warp per row, four rows in flight per warp.
constexpr int kOpWarps = 8, kOpRowsPerWarp = 4;
__global__ void __launch_bounds__(32 * kOpWarps)
knn_tc_operand_kernel(const float* __restrict__ xt, const float* __restrict__ xx, const unsigned* __restrict__ cmax, int C, int N, int C64, int Kp16,
                      uint2* __restrict__ xb) {
    cudaGridDependencySynchronize();  // pass 1 is complete: the cloud's largest norm is final
    cudaTriggerProgrammaticLaunchCompletion();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, n0 = (blockIdx.x * kOpWarps + w) * kOpRowsPerWarp;
    if (n0 >= N) return;
    const float sc = knn_tc_scale(__ldg(cmax + b));
    const int qp = Kp16 / 4;  // 8-byte groups per operand row
    auto pack = [](float a_, float b_) { return (unsigned)__half_as_ushort(__float2half_rn(a_)) | ((unsigned)__half_as_ushort(__float2half_rn(b_)) << 16); };
    for (int p0 = 0; p0 < qp; p0 += 32) {
        const int p = p0 + lane, c0 = 4 * p;
        float4 v[kOpRowsPerWarp];
        float nx[kOpRowsPerWarp];
#pragma unroll
        for (int r = 0; r < kOpRowsPerWarp; ++r) {  // all loads first
            const size_t row = (size_t)b * N + min(n0 + r, N - 1);
            v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            nx[r] = 0.f;
            if (c0 < C) v[r] = __ldg(reinterpret_cast<const float4*>(xt + row * C) + p);  // C is a multiple of 8
            else if (c0 == C64 + kTcAux) nx[r] = __ldg(xx + row);
        }
#pragma unroll
        for (int r = 0; r < kOpRowsPerWarp; ++r) {
            if (p >= qp || n0 + r >= N) continue;
            uint2 o = make_uint2(0u, 0u);
            if (c0 < C) {
                o.x = pack(v[r].x * sc, v[r].y * sc);
                o.y = pack(v[r].z * sc, v[r].w * sc);
            } else if (c0 == C64) {
                o.x = 0x3c003c00u, o.y = 0x00003c00u;  // 1 1 1 0
            } else if (c0 == C64 + kTcAux) {
                const float t = -0.5f * nx[r] * sc * sc;            // exact: powers of two
                const __half hh = __float2half_rz(t);
                const float t1 = __fsub_rn(t, __half2float(hh));    // exact
                const __half hm = __float2half_rz(t1);
                const __half hl = __float2half_rn(__fsub_rn(t1, __half2float(hm)));  // exact unless below the FP16 subnormal step 2^-24
                o.x = (unsigned)__half_as_ushort(hh) | ((unsigned)__half_as_ushort(hm) << 16);
                o.y = (unsigned)__half_as_ushort(hl);
            }
            xb[((size_t)b * N + n0 + r) * qp + p] = o;
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
// D[tmem] (+)= A[smem] * B[smem]^T, FP16 operands, FP32 accumulation, K = 16 per instruction; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor of a K-major operand tile in the 128-byte swizzle: rows of 128 bytes (64 FP16), groups of
// 8 rows 1024 bytes apart (SBO); bits 46-47 = 1 (sm_100 descriptor version), bits 61-63 = 2 (SWIZZLE_128B). A K step of 16
// FP16 values advances the start address by 32 bytes inside the swizzle row.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// the auxiliary group: rows of 32 bytes (one K step) in the 32-byte swizzle, groups of 8 rows 256 bytes apart, bits 61-63 = 6
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (bits 7-9, 10-12 = 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (1u << 4) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);

// 64 consecutive accumulator columns of this thread's TMEM lane (one wait per 64 values)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}

#ifdef FPSG_TUNING
// timeline of CTA 0 (dbg bit 8: `dump` is a u64 buffer [3 roles][512]): role 0 producer, 1 MMA issuer, 2 first epilogue warp
#define TC_STAMP(role, slot)                                                                              \
    do {                                                                                                  \
        if ((a.dbg & 8) && blockIdx.x == 0 && (slot) < 512) {                                             \
            unsigned long long t_;                                                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
            reinterpret_cast<unsigned long long*>(a.dump)[(role) * 512 + (slot)] = t_;                    \
        }                                                                                                 \
    } while (0)
#else
#define TC_STAMP(role, slot) \
    do {                     \
    } while (0)
#endif

struct TcArgs {
    const float* xx;   // [B*N]
    const unsigned* cmax;  // [B] bits of the cloud's largest squared norm
    unsigned* masks;   // [B*N][W]
    int N, k, C64, NK, stages, W, nblk, gshift, tiles_per_cloud, ntiles;
    unsigned a_bytes, stage_bytes;
    float* dump;       // tuning builds: [B*N][N] raw accumulator values of pass 1, else nullptr
    int dbg;           // tuning builds: ablation bits (1: epilogue without TMEM loads and arithmetic, 2: no MMA issue, 4: no TMA for B)
};

// ---------------------------------------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_aux, const TcArgs a) {
    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment (128-byte swizzle) by an OFFSET from the shared array: the pointers keep their address space (LDS / STS)
    unsigned char* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = base;                    // the query tile: [NK] chunks of 128 rows x 128 B (128-byte swizzle), then 128 rows x 32 B
    unsigned char* sB = sA + a.a_bytes;          // [stages] a candidate block: [NK] chunks of 256 rows x 128 B, then 256 rows x 32 B
    float* sGm = reinterpret_cast<float*>(sB + (size_t)a.stages * a.stage_bytes);  // [32 groups][256 row halves]
    float* sThr = sGm + kTcGroups * kTcGmLd;                                        // [256 row halves] pass-2 thresholds
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sThr + kTcGmLd);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 16);

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

    cudaGridDependencySynchronize();            // the operand rows and xx (prep kernel) are complete
    cudaTriggerProgrammaticLaunchCompletion();  // the re-scoring grid may be scheduled as SMs free up; it waits for this grid itself

    const int NK = a.NK, stages = a.stages, nblk = a.nblk, N = a.N;
    const int colA = a.C64, colB = a.C64 + kTcAux;  // where the `1 1 1` group (A side) and the `h m l` group (B side) sit in a row

    if (warp == 0) {
        // ===== TMA producer: one stage = one block of 256 candidates with ALL its K chunks (one barrier per block) =====
        if (lane == 0) {
            uint32_t it = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++ti) {
                const int b = tile / a.tiles_per_cloud, i0 = (tile - b * a.tiles_per_cloud) * kTcRows;
                const int row0 = b * N;
                mbar_wait(bar_aempty, (ti & 1) ^ 1);  // every MMA that read the previous tile's queries has completed
                TC_STAMP(0, 400 + ti);
                mbar_expect_tx(bar_afull, a.a_bytes);
                for (int kc = 0; kc < NK; ++kc) tma_load_2d(smem_u32(sA + (size_t)kc * kTcChunkBytesA), &tmap, bar_afull, kc * kTcChunk, row0 + i0);
                tma_load_2d(smem_u32(sA + (size_t)NK * kTcChunkBytesA), &tmap_aux, bar_afull, colA, row0 + i0);
                for (int pb = 0; pb < 2 * nblk; ++pb, ++it) {
                    const int blk = pb < nblk ? pb : pb - nblk;
                    const uint32_t s = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
#ifdef FPSG_TUNING
                    if (a.dbg & 4) {
                        mbar_arrive(bar_full + 8 * s);
                        continue;
                    }
#endif
                    mbar_expect_tx(bar_full + 8 * s, a.stage_bytes);
                    const uint32_t dst = smem_u32(sB + (size_t)s * a.stage_bytes);
                    const int r0 = row0 + blk * kTcBlk;
                    for (int kc = 0; kc < NK; ++kc) {
                        tma_load_2d(dst + kc * kTcChunkBytesB, &tmap, bar_full + 8 * s, kc * kTcChunk, r0);
                        tma_load_2d(dst + kc * kTcChunkBytesB + kTcChunkBytesB / 2, &tmap, bar_full + 8 * s, kc * kTcChunk, r0 + kTcBlk / 2);
                    }
                    tma_load_2d(dst + NK * kTcChunkBytesB, &tmap_aux, bar_full + 8 * s, colB, r0);
                    tma_load_2d(dst + NK * kTcChunkBytesB + kTcAuxBytesB / 2, &tmap_aux, bar_full + 8 * s, colB, r0 + kTcBlk / 2);
                    TC_STAMP(0, it);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): 4 NK + 1 instructions and two commits per block =====
        if (lane == 0) {
            const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
            uint32_t it = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++ti) {
                mbar_wait(bar_afull, ti & 1);
                tc_fence_after();
                TC_STAMP(1, 400 + ti);
                for (int pb = 0; pb < 2 * nblk; ++pb, ++it) {
                    const uint32_t buf = it & 1u;
                    mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
                    const uint32_t s = it % (uint32_t)stages;
                    mbar_wait(bar_full + 8 * s, (it / (uint32_t)stages) & 1u);
                    tc_fence_after();
                    TC_STAMP(1, 2 * it);
                    const uint32_t d = tmem + buf * kTcBlk;
                    const uint32_t bbase = sB_addr + s * a.stage_bytes;
#ifdef FPSG_TUNING
                    if (!(a.dbg & 2))
#endif
                    {
                        uint32_t acc = 0;
                        for (int kc = 0; kc < NK; ++kc) {
#pragma unroll
                            for (int s16 = 0; s16 < kTcChunk / 16; ++s16) {
                                umma_f16(d, umma_desc(sA_addr + kc * kTcChunkBytesA + s16 * 32), umma_desc(bbase + kc * kTcChunkBytesB + s16 * 32), kTcIdesc, acc);
                                acc = 1;
                            }
                        }
                        umma_f16(d, umma_desc_sw32(sA_addr + NK * kTcChunkBytesA), umma_desc_sw32(bbase + NK * kTcChunkBytesB), kTcIdesc, acc);
                    }
                    umma_commit(bar_empty + 8 * s);    // the stage is free once these MMAs have read it
                    umma_commit(bar_tfull + 8 * buf);  // accumulator complete
                    TC_STAMP(1, 2 * it + 1);
                }
                umma_commit(bar_aempty);
            }
        }
    } else {
        // ===== epilogue: 16 warps, a thread = one query row x one quarter (64 columns) of every 256-column block =====
        // The two threads (sub = 0, 1) of a row half share that half's 32 group slots (disjoint groups); the sub = 0 thread
        // ("main") sorts them and derives the threshold with the main thread of the other half.
        const int te = threadIdx.x - 64, e = te >> 5;
        const int quad = warp & 3, qtr = e >> 2;  // TMEM lanes 32 * (warp % 4) .. + 31 are the ones this warp may read
        const int half = qtr >> 1, sub = qtr & 1;
        const int rowt = quad * 32 + lane;
        const int col = half * kTcRows + rowt;     // this row half's column of sGm / sThr
        float* gm = sGm + col;
        const float* gmp = sGm + (col ^ kTcRows);  // the other half of this row
        const int GS = 1 << a.gshift;
        uint32_t tb = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int b = tile / a.tiles_per_cloud, i0 = (tile - b * a.tiles_per_cloud) * kTcRows;
            const int i = i0 + rowt;
            const bool row_ok = i < N;
            const float* xxb = a.xx + (size_t)b * N;
            // the error bound needs |x_i|^2, max_j |x_j|^2 of the cloud and the operand scale: loaded here, consumed after pass 1
            const unsigned cmax_bits = __ldg(a.cmax + b);
            const float xxi = __ldg(xxb + min(i, N - 1));
            asm volatile("bar.sync 1, 512;" ::: "memory");  // everybody is done with the previous tile's shared values
            if (sub == 0) {
#pragma unroll
                for (int g = 0; g < kTcGroups; ++g) gm[g * kTcGmLd] = -CUDART_INF_F;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");

            // ---- pass 1: group maxima ----
            for (int blk = 0; blk < nblk; ++blk, ++tb) {
                const uint32_t buf = tb & 1u;
                mbar_wait(bar_tfull + 8 * buf, (tb >> 1) & 1u);
                tc_fence_after();
                if (te == 0) TC_STAMP(2, 2 * tb);
                const int jb = blk * kTcBlk + qtr * 64;
                bool live = jb < N;  // warp-uniform
#ifdef FPSG_TUNING
                if (a.dbg & 1) live = false;
#endif
                if (live) {
                    uint32_t v64[64];
                    tmem_ld64(tmem + ((uint32_t)(quad * 32) << 16) + buf * kTcBlk + qtr * 64, v64);
#pragma unroll
                    for (int u2 = 0; u2 < 2; ++u2) {
                        const int j0 = jb + 32 * u2;
                        if (j0 >= N) break;
                        uint32_t* v = v64 + 32 * u2;
                        if (j0 + 32 > N) {
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                if (j0 + r >= N) v[r] = 0xff800000u;  // -INF: a column past the cloud never counts
                        }
#ifdef FPSG_TUNING
                        if (a.dump != nullptr && !(a.dbg & 8) && row_ok) {
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
                        const int lc0 = blk * (kTcBlk / 2) + sub * 64 + u2 * 32;  // index of v[0] among the columns of this row half
                        const int g = lc0 >> a.gshift;
                        if (GS == 8) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) gm[(g + u) * kTcGmLd] = m8[u];
                        } else if (GS == 16) {
                            gm[g * kTcGmLd] = fmaxf(m8[0], m8[1]);
                            gm[(g + 1) * kTcGmLd] = fmaxf(m8[2], m8[3]);
                        } else {  // GS = 32 or 64: the group lies inside this thread's 64 columns
                            gm[g * kTcGmLd] = max3(max3(m8[0], m8[1], m8[2]), m8[3], gm[g * kTcGmLd]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                if (te == 0) TC_STAMP(2, 2 * tb + 1);
            }

            // ---- threshold: k-th largest of the row's <= 64 group maxima (32 per row half, sorted in registers, merged) ----
            asm volatile("bar.sync 1, 512;" ::: "memory");  // all group maxima are written
            if (sub == 0) {
                {
                    float s_[kTcGroups];
#pragma unroll
                    for (int c = 0; c < kTcGroups; ++c) s_[c] = gm[c * kTcGmLd];
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
                    for (int c = 0; c < kTcGroups; ++c) gm[c * kTcGmLd] = s_[c];
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");  // the main threads only: both halves of every row are sorted
                // k-th largest of the union of the two descending arrays A (this half) and B (the other half): taking t values from
                // A and k - t from B, the k-th largest is >= min(A[t-1], B[k-t-1]) for every split t and equal to it for the best
                // one -- k + 1 independent terms (four running maxima) instead of a dependent merge loop
                auto term = [&](int t) { return fminf(gm[(t - 1) * kTcGmLd], gmp[(a.k - t - 1) * kTcGmLd]); };
                float Ta = fmaxf(gm[(a.k - 1) * kTcGmLd], gmp[(a.k - 1) * kTcGmLd]);  // t = k and t = 0
                float Tb = -CUDART_INF_F, Tc = -CUDART_INF_F, Td = -CUDART_INF_F;
                int t = 1;
                for (; t + 3 < a.k; t += 4) {
                    Ta = fmaxf(Ta, term(t));
                    Tb = fmaxf(Tb, term(t + 1));
                    Tc = fmaxf(Tc, term(t + 2));
                    Td = fmaxf(Td, term(t + 3));
                }
                for (; t < a.k; ++t) Ta = fmaxf(Ta, term(t));
                const float T = fmaxf(fmaxf(Ta, Tb), fmaxf(Tc, Td));
                // |s_tensor - s^2 s_fp32| <= E for every candidate of this row (s = the operand scale): FP16 (round to nearest)
                // keeps 11 significant bits of each operand (relative error <= 2^-11 each, <= 2^-10 (1 + 2^-12) per product;
                // Cauchy-Schwarz over the channels: 2^-10 |x_i| |x_j|), the FP32 chain of the reference and the FP32 accumulator
                // add O(C 2^-23) of the same magnitude (C <= 128: < 2^-16), values below the FP16 normal range after scaling
                // (|s x| < 2^-14 against a largest norm of 2^6) lose at most 2^-25 each, the h m l pieces are exact (or off by
                // 2^-25) and the roundings of the two subtractions of pd are below 2^-21 max|x|^2. 1.25 x 2^-10 covers the
                // first three, 2^-16 max|x|^2 the rest.
                const float sc = knn_tc_scale(cmax_bits), mxa = __uint_as_float(cmax_bits);
                const float E = (0.001220703125f * sqrtf(xxi * mxa) + 1.52587890625e-5f * mxa) * (sc * sc);
                sThr[col] = T - 2.0f * E;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float thr = sThr[col];

            // ---- pass 2: mark every candidate that can still be among the k nearest ----
            for (int blk = 0; blk < nblk; ++blk, ++tb) {
                const uint32_t buf = tb & 1u;
                mbar_wait(bar_tfull + 8 * buf, (tb >> 1) & 1u);
                tc_fence_after();
                if (te == 0) TC_STAMP(2, 2 * tb);
                const int jb = blk * kTcBlk + qtr * 64;
                bool live = jb < N;
#ifdef FPSG_TUNING
                if (a.dbg & 1) live = false;
#endif
                if (live) {
                    uint32_t v64[64];
                    tmem_ld64(tmem + ((uint32_t)(quad * 32) << 16) + buf * kTcBlk + qtr * 64, v64);
#pragma unroll
                    for (int u2 = 0; u2 < 2; ++u2) {
                        const int j0 = jb + 32 * u2;
                        if (j0 >= N) break;
                        const uint32_t* v = v64 + 32 * u2;
                        unsigned m4[4] = {0u, 0u, 0u, 0u};  // four short chains instead of one of 32 dependent ORs
#pragma unroll
                        for (int r = 0; r < 32; ++r) m4[r & 3] |= __uint_as_float(v[r]) >= thr ? (1u << r) : 0u;
                        unsigned m = (m4[0] | m4[1]) | (m4[2] | m4[3]);
                        if (j0 + 32 > N) m &= (1u << (N - j0)) - 1u;
                        if (row_ok) a.masks[((size_t)b * N + i) * a.W + (j0 >> 5)] = m;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                if (te == 0) TC_STAMP(2, 2 * tb + 1);
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
// A warp per query, a lane per candidate (the FMA chain over the channels is sequential by definition, so a candidate cannot
// be split over lanes). Letting every lane read ITS candidate's row straight from global memory costs one 32-byte sector
// per lane and instruction (85 us at c3, C = 64: the load/store unit, not the 235 MB, is the limit); instead the 32 rows of
// a round are copied, 64 channels at a time, to a warp-private shared-memory tile with COALESCED 16-byte cp.async (16 lanes x
// 16 B = one 256-byte row segment, two rows per instruction) and each lane then walks its own row with conflict-free LDS.128
// (row stride 68 floats). (32-channel tiles for twice the resident warps were measured: no gain at C = 64, slower at 128.)
constexpr int kRsWarps = 4;
constexpr int kRsCH = 64;          // channels staged per step
constexpr int kRsLD = kRsCH + 4;   // row stride of the tile: 16-byte aligned, lanes 0-7 of an LDS.128 phase hit distinct bank groups
constexpr int kRsListCap = 128;    // candidates compacted into a list; beyond that: word-by-word streaming

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

static size_t knn_tc_rescore_smem_per_warp(int C) { return (size_t)32 * kRsLD * sizeof(float) + align_up((size_t)C * sizeof(float), 16) + kRsListCap * sizeof(unsigned short); }

// NW = mask words per lane (W <= 32 NW)
template <int NW>
__global__ void __launch_bounds__(kRsWarps * 32, 6)
knn_tc_rescore_kernel(const float* __restrict__ xt, const float* __restrict__ xx, const unsigned* __restrict__ masks, int rows, int N, int k,
                      int C, int Kp, int W, int64_t* __restrict__ idx) {
    constexpr int kRsWords = NW;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t qbytes = ((size_t)C * sizeof(float) + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)32 * kRsLD * sizeof(float) + qbytes + kRsListCap * sizeof(unsigned short);
    float* tile = reinterpret_cast<float*>(rs_smem + (size_t)w * per_warp);            // [32 candidates][kRsLD]
    float* sQ = tile + 32 * kRsLD;                                                     // [C] the query's row
    unsigned short* sList = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(sQ) + qbytes);
    const int sub = lane >> 4, l16 = lane & 15;  // staging: 16 lanes x 16 B = one 256-byte row segment, two rows per instruction

    cudaGridDependencySynchronize();  // the filter's masks are complete
    const int tw = (int)gridDim.x * kRsWarps;  // rows = B * N < 2^30 (plan): 32-bit row arithmetic throughout
    int row = (int)blockIdx.x * kRsWarps + w;
    unsigned mwn[kRsWords];
#pragma unroll
    for (int s = 0; s < kRsWords; ++s) {
        const int wi = s * 32 + lane;
        mwn[s] = (row < rows && wi < W) ? __ldcg(masks + (size_t)row * W + wi) : 0u;
    }

    for (; row < rows; row += tw) {  // warp-uniform; no CTA-wide barrier in this kernel
        unsigned mw[kRsWords];
        int mine = 0;
#pragma unroll
        for (int s = 0; s < kRsWords; ++s) {
            mw[s] = mwn[s];
            mine += __popc(mw[s]);
        }
        {  // the next query's mask words travel while this one is scored
            const int nrow = row + tw;
#pragma unroll
            for (int s = 0; s < kRsWords; ++s) {
                const int wi = s * 32 + lane;
                mwn[s] = (nrow < rows && wi < W) ? __ldcg(masks + (size_t)nrow * W + wi) : 0u;
            }
        }
        const int qi = (int)((unsigned)row % (unsigned)N);
        const int cloud0 = row - qi;  // first row of this query's cloud
        const float* __restrict__ xc = xt + (size_t)cloud0 * Kp;  // the cloud's rows; j * Kp < 2^21 fits 32 bits
        const float* __restrict__ qrow = xt + (size_t)row * Kp;
        const float xxi = __ldg(xx + row);
        __syncwarp();  // the previous query's reads of sQ / tile are done
        for (int c4 = lane; c4 < C / 4; c4 += 32) cp_async16(sQ + 4 * c4, qrow + 4 * c4);  // completes with the first tile's copies

        // pd of the lane's candidate j, exactly as the reference rounds it: dot = FMA chain over ascending c from 0 (model.py:14),
        // pd = rn(rn(2 dot - xx_j) - xx_i) (model.py:16). Lanes without a candidate return 0; they copy and score the query's own
        // row (skipping those copies saves 22 % of the gather bytes
        // and 2 % of the time: the kernel is bound by instructions and their dependent chains, so the staging is kept branch-free)
        auto score32 = [&](int j, bool valid) -> u64 {
            if (!valid) j = qi;
            const float xxj = __ldg(xx + cloud0 + j);
            const unsigned joff = (unsigned)(j * Kp) * 4u;  // byte offset of the candidate's row inside the cloud
            float dot = 0.f;
            for (int ch0 = 0; ch0 < C; ch0 += kRsCH) {
                const int cw = min(kRsCH, C - ch0);
                const char* __restrict__ src = reinterpret_cast<const char*>(xc + ch0 + 4 * l16);  // + the candidate's row offset
                const unsigned dst = smem_u32(tile + sub * kRsLD + 4 * l16);                         // + 2 r rows
                const unsigned nbytes = 4 * l16 < cw ? 16u : 0u;  // lanes past a partial chunk copy nothing (zero fill)
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const unsigned jo = __shfl_sync(0xffffffffu, joff, 2 * r + sub);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 2 * r * kRsLD * 4), "l"(src + jo), "r"(nbytes) : "memory");
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                const float4* tr = reinterpret_cast<const float4*>(tile + lane * kRsLD);
                const float4* qr = reinterpret_cast<const float4*>(sQ + ch0);
#pragma unroll 4
                for (int c4 = 0; c4 < cw / 4; ++c4) {
                    const float4 v = tr[c4], q = qr[c4];
                    dot = __fmaf_rn(q.x, v.x, dot);
                    dot = __fmaf_rn(q.y, v.y, dot);
                    dot = __fmaf_rn(q.z, v.z, dot);
                    dot = __fmaf_rn(q.w, v.w, dot);
                }
                __syncwarp();  // the tile may be overwritten
            }
            return valid ? knn_key_of(__fsub_rn(__fmaf_rn(2.0f, dot, -xxj), xxi), j) : 0ull;
        };

        int total = (int)__reduce_add_sync(0xffffffffu, (unsigned)mine);  // REDUX: one instruction, and the compiler knows it is warp-uniform
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

        // Rounds of 32 candidates. Usually the marked candidates (about 28) are compacted into a list and take one or two
        // rounds; with massive ties (or the take-everything case) the mask is streamed word by word, 32 candidates per round.
        const bool listed = total <= kRsListCap;
        if (listed) {
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
                    sList[pos++] = (unsigned short)((s * 32 + lane) * 32 + bit);
                }
            }
            __syncwarp();
        }
        u64 key = 0ull;
        bool first = true;
        int cursor = 0;  // listed: next list element; streaming: next mask word
        for (;;) {       // warp-uniform control flow
            int j = 0;
            bool valid = false;
            if (listed) {
                if (cursor >= total) break;
                const int el = cursor + lane;
                valid = el < total;
                if (valid) j = (int)sList[el];
                cursor += 32;
            } else {
                unsigned word = 0u;
                while (cursor < W) {
                    unsigned sel = mw[0];
#pragma unroll
                    for (int s = 1; s < kRsWords; ++s)
                        if ((cursor >> 5) == s) sel = mw[s];
                    word = __shfl_sync(0xffffffffu, sel, cursor & 31);
                    ++cursor;
                    if (word != 0u) break;
                }
                if (word == 0u) break;
                j = (cursor - 1) * 32 + lane;
                valid = ((word >> lane) & 1u) != 0u;
            }
            __syncwarp();  // (also a hint: the warp is converged here, so the shuffles below need no convergence wrappers)
            u64 kb = score32(j, valid);
            __syncwarp();
            kb = warp_sort_desc(kb, lane);
            if (first) {
                key = kb;
                first = false;
            } else {
                const u64 rev = __shfl_sync(0xffffffffu, kb, 31 - lane);
                key = warp_merge_desc(key > rev ? key : rev, lane);
            }
        }
        if (lane < k) idx[(size_t)row * k + lane] = (int64_t)(0xffffffffu - (unsigned)key);
    }
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
    unsigned* cmax = reinterpret_cast<unsigned*>(wsb + p.cmax_off);
    float* xx = reinterpret_cast<float*>(wsb + p.xx_off);
    float* xt = reinterpret_cast<float*>(wsb + p.xt_off);
    unsigned* xb = reinterpret_cast<unsigned*>(wsb + p.xb_off);
    unsigned* masks = reinterpret_cast<unsigned*>(wsb + p.mask_off);

    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(FPSG_ECUDA, "knn: cuTensorMapEncodeTiled is not available from this driver");
    // the FP16 operand rows as a 2-D tensor (Kp16, B*N); two views: 64-channel chunks (128-byte swizzle) and the 16-value groups
    CUtensorMap tmap, tmap_aux;
    const cuuint64_t gdim[2] = {(cuuint64_t)p.Kp16, (cuuint64_t)B * (cuuint64_t)N};
    const cuuint64_t gstr[1] = {(cuuint64_t)p.Kp16 * sizeof(unsigned short)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcChunk, (cuuint32_t)kTcRows};
    const cuuint32_t box_aux[2] = {(cuuint32_t)kTcAux, (cuuint32_t)kTcRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, xb, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS)
        cr = enc(&tmap_aux, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, xb, gdim, gstr, box_aux, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(FPSG_ECUDA, "knn: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    FPSG_CUDA_OK(cudaMemsetAsync(cmax, 0, (size_t)B * sizeof(unsigned), stream));
    const size_t prep_smem = (size_t)C * (kPrepPts + 1) * sizeof(float);
    if (prep_smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(knn_tc_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    const dim3 pgrid((unsigned)cdiv(N, kPrepPts), (unsigned)B);
    FPSG_CUDA_OK(launch_pdl(knn_tc_prep_kernel, pgrid, dim3(kPrepThreads), prep_smem, stream, x, C, N, xx, xt, cmax));
    FPSG_CUDA_OK(launch_pdl(knn_tc_operand_kernel, dim3((unsigned)cdiv(N, kOpWarps * kOpRowsPerWarp), (unsigned)B), dim3(32 * kOpWarps), 0, stream,
                            (const float*)xt, (const float*)xx, (const unsigned*)cmax, C, N, p.C64, p.Kp16, reinterpret_cast<uint2*>(xb)));

    TcArgs a;
    a.xx = xx, a.cmax = cmax, a.masks = masks, a.N = N, a.k = k, a.C64 = p.C64, a.NK = p.NK, a.stages = p.stages, a.W = p.W, a.nblk = p.nblk, a.gshift = p.gshift;
    a.tiles_per_cloud = p.tiles_per_cloud, a.ntiles = p.ntiles, a.dump = dump;
    a.a_bytes = (unsigned)p.a_bytes, a.stage_bytes = (unsigned)p.stage_bytes;
    a.dbg = tune_env("FPSG_TC_DBG") ? atoi(tune_env("FPSG_TC_DBG")) : 0;
    FPSG_CUDA_OK(cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const int grid = min(p.ntiles, sm_count());
    FPSG_CUDA_OK(launch_pdl(knn_tc_filter_kernel, dim3((unsigned)grid), dim3(kTcThreads), p.smem, stream, tmap, tmap_aux, a));
    if (!rescoring) return FPSG_OK;

    const int rows = B * N;
    const size_t rs_smem = kRsWarps * knn_tc_rescore_smem_per_warp(C);
    void (*rs_kernel)(const float*, const float*, const unsigned*, int, int, int, int, int, int, int64_t*) =
        p.W <= 32 ? knn_tc_rescore_kernel<1> : (p.W <= 64 ? knn_tc_rescore_kernel<2> : (p.W <= 128 ? knn_tc_rescore_kernel<4> : knn_tc_rescore_kernel<8>));
    if (rs_smem > 48 * 1024) FPSG_CUDA_OK(cudaFuncSetAttribute(rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    int per_sm = (int)((size_t)(220 * 1024) / (rs_smem + 1024));  // resident CTAs per SM (shared-memory limited), at most 6 (the kernel's launch bound)
    per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
    const long long want = ((long long)rows + kRsWarps - 1) / kRsWarps;
    const long long rs_grid = want < (long long)per_sm * sm_count() ? want : (long long)per_sm * sm_count();
    FPSG_CUDA_OK(launch_pdl(rs_kernel, dim3((unsigned)rs_grid), dim3(kRsWarps * 32), rs_smem, stream, (const float*)xt, (const float*)xx,
                            (const unsigned*)masks, rows, N, k, C, C, p.W, idx));
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
