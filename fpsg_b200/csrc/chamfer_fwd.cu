// K1: fused bidirectional Chamfer forward for sm_100a.
//
// Replaces the two `sided_distance` forward launches behind the reference's
// `chamfer_distance(syn_pc, ref_pc)` (reference src/models/few_shot.py:110,117,167; Kaolin 0.9.0
// sided_distance_forward_cuda_kernel). One sweep over the N x M distance matrix of a pair
// serves both directions: every distance is evaluated once (6 FP32-pipe instructions) and feeds a row minimum
// and a column minimum.
//
//   sweep kernel    : persistent grid (a multiple of the SM count). The work of the whole batch is cut into
//                     units of (pair, row strip, 16 columns) and every CTA takes an equal contiguous run of
//                     units, so all SMs finish together whatever B, N, M are. CTA = 128 threads; thread t keeps R
//                     consecutive rows (p1 points) in registers and streams <=64-column sub-blocks of p2
//                     through shared memory (AoS, three broadcast LDS.128 per 4 columns; the next sub-block is
//                     prefetched into registers during the current one).
//                     Row direction : best[k] = min3(best[k], d, d') -- value only; once per sub-block the
//                                     thread notes (in shared memory) which sub-block last lowered best[k].
//                     Column direction: per column the thread's minimum over its R rows, then one warp-wide
//                                     CREDUX.MIN.F32 + ballot gives (warp minimum, lanes attaining it); lane 0
//                                     parks that in shared memory and the warp writes its 64 results coalesced.
//                     There is no arg-min bookkeeping per distance and no CTA-wide reduction.
//   finalize kernel : merges the partial minima (rows: over the CTAs that covered the row's strip; columns:
//                     over strips x warps) and recovers the arg-min exactly by re-evaluating the few candidates
//                     the partials point at (the 64 columns of the winning sub-block for a row, the R rows of
//                     the winning thread for a column) with the same FP32 instruction sequence, taking the
//                     lowest index whose distance equals the minimum bit-for-bit. That reproduces the
//                     reference's strict-'<' ascending scan (lowest index wins ties).
//   loss kernel     : optional, w1*mean(dist1) + w2*mean(dist2) per pair from per-CTA partial sums (fixed
//                     order, double) -- the reduction glue of few_shot.py:110 / kaolin's chamfer_distance.
//
// Out-of-range rows and columns of ragged tiles are filled with +INF coordinates: their distances are +INF,
// never win a strict comparison, and never equal a finite minimum.
#include <cstdlib>

#include "common.cuh"

namespace fpsg {

typedef unsigned long long u64;

constexpr int kSweepThreads = 128;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kUnit = 16;           // columns per scheduling unit
constexpr int kSub = 64;            // columns per shared-memory sub-block (4 units)
constexpr int kFinThreads = 256;

struct ChamferPlan {
    int R;            // rows per thread in the sweep
    int TN;           // rows per strip (kSweepThreads * R)
    int strips;       // row strips per pair
    int ups;          // 16-column units per (pair, strip)
    long long units;  // all units of the batch (must fit an int: checked at launch)
    int grid;         // persistent CTAs
    int q, r;         // CTA g owns q + (g < r) units starting at g*q + min(g, r)
    int maxslots;     // most CTAs that can intersect one (pair, strip)
    int Mpad;         // M rounded up to a unit
    int nrb, ncb;     // finalize CTAs per pair: row blocks, column blocks
    size_t rpval_off, rpsub_off, colpart_off, losspart_off, total_bytes;
};

static int g_sm_count = 0;

static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count = n;
        else
            return 148;
    }
    return g_sm_count;
}

ChamferPlan make_chamfer_plan(int B, int N, int M) {
    ChamferPlan p;
    // rows per thread: least padded work; the per-column warp reduction amortises better with more rows
    const int cand[3] = {16, 8, 4};
    double best_cost = 1e300;
    p.R = 16;
    for (int c = 0; c < 3; ++c) {
        int tn = kSweepThreads * cand[c];
        double cost = (double)cdiv(N, tn) * tn * (1.0 + 0.5 / cand[c]);
        if (cost < best_cost) {
            best_cost = cost;
            p.R = cand[c];
        }
    }
    if (const char* e = getenv("FPSG_SWEEP_R")) {  // tuning override (benchmarks only): 4, 8 or 16
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) p.R = v;
    }
    p.TN = kSweepThreads * p.R;
    p.strips = cdiv(N, p.TN);
    p.ups = cdiv(M, kUnit);
    p.Mpad = p.ups * kUnit;
    p.units = (long long)B * p.strips * p.ups;
    const int per_sm = p.R >= 16 ? 4 : 6;
    long long g = (long long)sm_count() * per_sm;
    if (g > p.units) g = p.units;
    if (g < 1) g = 1;
    p.grid = (int)g;
    p.q = (int)(p.units / g);
    p.r = (int)(p.units % g);
    p.maxslots = p.q > 0 ? cdiv(p.ups, p.q) + 1 : 1;
    p.nrb = cdiv(N, kFinThreads);
    p.ncb = cdiv(M, kFinThreads);
    size_t off = 0;
    const size_t rows = (size_t)B * p.strips * p.maxslots * p.TN;
    p.rpval_off = off;
    off += align_up(rows * sizeof(float), 256);
    p.rpsub_off = off;
    off += align_up(rows * sizeof(int), 256);
    p.colpart_off = off;
    off += align_up((size_t)B * p.strips * kSweepWarps * p.Mpad * sizeof(u64), 256);
    p.losspart_off = off;
    off += align_up((size_t)B * (p.nrb + p.ncb) * sizeof(double), 256);
    p.total_bytes = off;
    return p;
}

struct SweepArgs {
    const float* p1;
    const float* p2;
    int N, M, strips, ups, q, r, maxslots, Mpad;
    float* rpval;   // [B*strips][maxslots][TN]   row partial minima
    int* rpsub;     // [B*strips][maxslots][TN]   sub-block (64 columns) that produced them
    u64* colpart;   // [B*strips][4 warps][Mpad]  (min bits << 32) | ballot of lanes attaining it
};

// which CTA owns unit u under the q/r partition
__device__ __forceinline__ int cta_of_unit(int u, int q, int r) {
    const int cut = r * (q + 1);
    return u < cut ? u / (q + 1) : r + (u - cut) / q;
}

__device__ __forceinline__ float warp_min_f32(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

// ---------------------------------------------------------------------------------------------------------
// sweep
// ---------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kSweepThreads, (R >= 16 ? 4 : 6)) chamfer_sweep_kernel(const SweepArgs a) {
    // column tile, AoS xyz, double buffered; always written in full (INF beyond the piece), so the 4-column
    // vector loads of a ragged tail read +INF coordinates
    __shared__ __align__(16) float sCols[2][kSub * 3];
    __shared__ __align__(16) u64 sColRes[kSweepWarps][kSub];
    __shared__ float sPrev[R][kSweepThreads];  // best[k] as of the previous sub-block (thread-private slots)
    __shared__ int sSubIdx[R][kSweepThreads];  // sub-block that last lowered best[k]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = blockIdx.x;
    int u = g * a.q + min(g, a.r);
    const int uend = u + a.q + (g < a.r ? 1 : 0);

    float px[R], py[R], pz[R], best[R];
    int cur_bs = -1;   // (pair, strip) whose rows are in registers
    int slot = 0;

    // prefetch registers for the next sub-block's tile (<= 192 floats over 128 threads)
    float pf0 = CUDART_INF_F, pf1 = CUDART_INF_F;
    int buf = 0;

    // describes the sub-block piece starting at unit uu: its (pair,strip), column range, and the unit after it
    auto piece = [&](int uu, int& bs, int& cbeg, int& cend, int& unext) {
        bs = uu / a.ups;
        const int cu = uu - bs * a.ups;
        int e = uu - (cu & 3) + 4;              // end of the 64-column sub-block
        const int seg_end = (bs + 1) * a.ups;  // end of the (pair, strip)
        e = e < seg_end ? e : seg_end;
        e = e < uend ? e : uend;
        unext = e;
        cbeg = cu * kUnit;
        cend = min(a.M, (cu + (e - uu)) * kUnit);
    };
    auto fetch = [&](int uu) {  // global -> registers; INF beyond the piece so padded columns never win
        int bs, un;
        int cb, ce;
        piece(uu, bs, cb, ce, un);
        const float* src = a.p2 + ((size_t)(bs / a.strips) * a.M + cb) * 3;
        const int n = (ce - cb) * 3;
        pf0 = tid < n ? __ldg(src + tid) : CUDART_INF_F;
        pf1 = tid + kSweepThreads < n ? __ldg(src + tid + kSweepThreads) : CUDART_INF_F;
    };

    if (u < uend) fetch(u);

    while (u < uend) {
        int bs, unext;
        int cbeg, cend;
        piece(u, bs, cbeg, cend, unext);
        const int b = bs / a.strips, strip = bs % a.strips;

        // park the prefetched tile in shared memory
        sCols[buf][tid] = pf0;
        if (tid + kSweepThreads < kSub * 3) sCols[buf][tid + kSweepThreads] = pf1;

        if (bs != cur_bs) {
            // new (pair, strip): bring this thread's R rows into registers
            cur_bs = bs;
            slot = g - cta_of_unit(bs * a.ups, a.q, a.r);
            const float* p1b = a.p1 + (size_t)b * a.N * 3;
            const int row0 = strip * (kSweepThreads * R) + tid * R;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int rr = row0 + k;
                const bool ok = rr < a.N;
                px[k] = ok ? __ldg(p1b + 3 * rr) : CUDART_INF_F;
                py[k] = ok ? __ldg(p1b + 3 * rr + 1) : CUDART_INF_F;
                pz[k] = ok ? __ldg(p1b + 3 * rr + 2) : CUDART_INF_F;
                best[k] = CUDART_INF_F;
                sPrev[k][tid] = CUDART_INF_F;
                sSubIdx[k][tid] = cbeg / kSub;
            }
        }
        __syncthreads();  // tile visible to all warps (and the other buffer is free again)
        if (unext < uend) fetch(unext);  // overlaps the whole sub-block below

        const int ncols = cend - cbeg;
        const int nsteps = (ncols + 3) >> 2;
        const float* tile = sCols[buf];
        u64* res = sColRes[w];
#pragma unroll 1
        for (int step = 0; step < nsteps; ++step) {
            const float4 v0 = *reinterpret_cast<const float4*>(tile + 12 * step);      // x0 y0 z0 x1
            const float4 v1 = *reinterpret_cast<const float4*>(tile + 12 * step + 4);  // y1 z1 x2 y2
            const float4 v2 = *reinterpret_cast<const float4*>(tile + 12 * step + 8);  // z2 x3 y3 z3
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
                c0m = min3(c0m, a0, b0);
                c1m = min3(c1m, a1, b1);
                c2m = min3(c2m, a2, b2);
                c3m = min3(c3m, a3, b3);
            }
            // column direction: warp minimum + the lanes that attain it, one 8-byte record per column
            const float m0 = warp_min_f32(c0m), m1 = warp_min_f32(c1m), m2 = warp_min_f32(c2m), m3 = warp_min_f32(c3m);
            const unsigned h0 = __ballot_sync(0xffffffffu, c0m == m0);
            const unsigned h1 = __ballot_sync(0xffffffffu, c1m == m1);
            const unsigned h2 = __ballot_sync(0xffffffffu, c2m == m2);
            const unsigned h3 = __ballot_sync(0xffffffffu, c3m == m3);
            if (lane == 0) {
                res[4 * step + 0] = ((u64)__float_as_uint(m0) << 32) | h0;
                res[4 * step + 1] = ((u64)__float_as_uint(m1) << 32) | h1;
                res[4 * step + 2] = ((u64)__float_as_uint(m2) << 32) | h2;
                res[4 * step + 3] = ((u64)__float_as_uint(m3) << 32) | h3;
            }
        }

        // row direction: did this sub-block lower best[k]? (thread-private shared-memory slots, no sync needed)
        {
            const int sub = cbeg / kSub;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (best[k] < sPrev[k][tid]) {
                    sPrev[k][tid] = best[k];
                    sSubIdx[k][tid] = sub;
                }
            }
        }
        // column direction: the warp's records for this piece, coalesced
        __syncwarp();
        {
            u64* dst = a.colpart + (((size_t)bs * kSweepWarps + w) * a.Mpad + cbeg);
            for (int c = lane; c < ncols; c += 32) dst[c] = res[c];
        }
        __syncwarp();

        // last piece of this (pair, strip) for this CTA: flush the row partials
        {
            int nbs = -1;
            if (unext < uend) nbs = unext / a.ups;
            if (nbs != bs) {
                const size_t base = ((size_t)bs * a.maxslots + slot) * (kSweepThreads * R) + tid * R;
#pragma unroll
                for (int k = 0; k < R; k += 4) {
                    *reinterpret_cast<float4*>(a.rpval + base + k) = make_float4(best[k], best[k + 1], best[k + 2], best[k + 3]);
                    *reinterpret_cast<int4*>(a.rpsub + base + k) =
                        make_int4(sSubIdx[k][tid], sSubIdx[k + 1][tid], sSubIdx[k + 2][tid], sSubIdx[k + 3][tid]);
                }
            }
        }
        u = unext;
        buf ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    const float* p1;
    const float* p2;
    int N, M, R, strips, ups, q, r, maxslots, Mpad;
    const float* rpval;
    const int* rpsub;
    const u64* colpart;
    float* dist1;
    float* dist2;
    int64_t* idx1;
    int64_t* idx2;
    double* losspart;
    int nrb, ncb;
};

__global__ void __launch_bounds__(kFinThreads) chamfer_finalize_kernel(const FinalizeArgs a) {
    __shared__ double sRed[kFinThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y, blk = blockIdx.x;
    const float* __restrict__ p1b = a.p1 + (size_t)b * a.N * 3;
    const float* __restrict__ p2b = a.p2 + (size_t)b * a.M * 3;
    const int TN = kSweepThreads * a.R;
    float mine = 0.f;  // this thread's distance, for the loss partial

    if (blk < a.nrb) {
        // ---- rows: each thread merges the partials of the CTAs that covered its row's strip ... ----
        const int i = blk * kFinThreads + tid;
        const bool valid = i < a.N;
        float m = CUDART_INF_F, ax = 0.f, ay = 0.f, az = 0.f;
        int sub = 0;
        if (valid) {
            const int strip = i / TN, within = i - strip * TN;
            const int bs = b * a.strips + strip;
            const int u0 = bs * a.ups;
            const int g0 = cta_of_unit(u0, a.q, a.r), g1 = cta_of_unit(u0 + a.ups - 1, a.q, a.r);
            for (int s = 0; s <= g1 - g0; ++s) {  // ascending slot = ascending columns: strict '<' keeps the lowest
                const size_t o = ((size_t)bs * a.maxslots + s) * TN + within;
                const float v = a.rpval[o];
                if (v < m) {
                    m = v;
                    sub = a.rpsub[o];
                }
            }
            ax = __ldg(p1b + 3 * i);
            ay = __ldg(p1b + 3 * i + 1);
            az = __ldg(p1b + 3 * i + 2);
            mine = m;
        }
        // ... then the warp scans the 64 columns of every row's winning sub-block together: two coalesced columns
        // per lane, and the lowest lane whose distance equals the minimum gives the arg-min
        int found = 0;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const float mm = __shfl_sync(0xffffffffu, m, rr);
            const int j0 = __shfl_sync(0xffffffffu, sub, rr) * kSub;
            const float bx = __shfl_sync(0xffffffffu, ax, rr), by = __shfl_sync(0xffffffffu, ay, rr), bz = __shfl_sync(0xffffffffu, az, rr);
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            float da = CUDART_INF_F, db = CUDART_INF_F;
            if (ja < a.M) da = sqdist(bx, by, bz, __ldg(p2b + 3 * ja), __ldg(p2b + 3 * ja + 1), __ldg(p2b + 3 * ja + 2));
            if (jb < a.M) db = sqdist(bx, by, bz, __ldg(p2b + 3 * jb), __ldg(p2b + 3 * jb + 1), __ldg(p2b + 3 * jb + 2));
            const unsigned ha = __ballot_sync(0xffffffffu, da == mm), hb = __ballot_sync(0xffffffffu, db == mm);
            const int f = ha ? j0 + __ffs(ha) - 1 : (hb ? j0 + 32 + __ffs(hb) - 1 : j0);
            if (lane == rr) found = f;
        }
        if (valid) {
            if (a.dist1) a.dist1[(size_t)b * a.N + i] = m;
            a.idx1[(size_t)b * a.N + i] = found;
        }
    } else {
        // ---- columns: each thread merges strips x warps ... ----
        const int j = (blk - a.nrb) * kFinThreads + tid;
        const bool valid = j < a.M;
        float m = CUDART_INF_F, qx = 0.f, qy = 0.f, qz = 0.f;
        int r0 = 0;
        if (valid) {
            for (int s = 0; s < a.strips; ++s) {
                for (int ww = 0; ww < kSweepWarps; ++ww) {  // ascending (strip, warp) = ascending rows
                    const u64 rec = a.colpart[(((size_t)b * a.strips + s) * kSweepWarps + ww) * a.Mpad + j];
                    const float v = __uint_as_float((unsigned)(rec >> 32));
                    if (v < m) {
                        m = v;
                        r0 = (s * kSweepThreads + ww * 32 + ((unsigned)rec ? __ffs((unsigned)rec) - 1 : 0)) * a.R;
                    }
                }
            }
            qx = __ldg(p2b + 3 * j);
            qy = __ldg(p2b + 3 * j + 1);
            qz = __ldg(p2b + 3 * j + 2);
            mine = m;
        }
        // ... then groups of R lanes scan the R consecutive rows of each column's winning thread (coalesced)
        const int R = a.R, per = 32 / R, grp = lane / R, off = lane - grp * R;
        const unsigned gmask = (R == 32 ? 0xffffffffu : ((1u << R) - 1u)) << (grp * R);
        int found = 0;
        for (int it = 0; it < R; ++it) {
            const int src = it * per + grp;  // the lane whose column this group scans now
            const float mm = __shfl_sync(0xffffffffu, m, src);
            const int rr0 = __shfl_sync(0xffffffffu, r0, src);
            const float cx = __shfl_sync(0xffffffffu, qx, src), cy = __shfl_sync(0xffffffffu, qy, src), cz = __shfl_sync(0xffffffffu, qz, src);
            const int r = rr0 + off;
            float d = CUDART_INF_F;
            if (r < a.N) d = sqdist(cx, cy, cz, __ldg(p1b + 3 * r), __ldg(p1b + 3 * r + 1), __ldg(p1b + 3 * r + 2));
            const unsigned hit = __ballot_sync(0xffffffffu, d == mm) & gmask;
            int f = hit ? rr0 + (__ffs(hit) - 1 - grp * R) : rr0;
            if (f >= a.N) f = 0;
            // hand the group's answer to the lane that owns the column
            const int want = lane - it * per;  // group index that worked for this lane's column, if in [0, per)
            const int fo = __shfl_sync(0xffffffffu, f, (want * R) & 31);
            if (want >= 0 && want < per) found = fo;
        }
        if (valid) {
            if (a.dist2) a.dist2[(size_t)b * a.M + j] = m;
            a.idx2[(size_t)b * a.M + j] = found;
        }
    }

    if (a.losspart) {
        // fixed-order block sum in double: lanes by shuffle tree, warps in ascending order
        double v = (double)mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[w] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int q = 0; q < kFinThreads / 32; ++q) s += sRed[q];
            a.losspart[(size_t)b * (a.nrb + a.ncb) + blk] = s;
        }
    }
}

// loss[b] = w1 * mean(dist1[b]) + w2 * mean(dist2[b]) from the finalize CTAs' partial sums
__global__ void __launch_bounds__(32)
chamfer_loss_kernel(const double* __restrict__ losspart, int nrb, int ncb, int N, int M, float w1, float w2,
                    float* __restrict__ loss) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const double* lp = losspart + (size_t)b * (nrb + ncb);
    double s1 = 0.0, s2 = 0.0;
    for (int q = lane; q < nrb; q += 32) s1 += lp[q];
    for (int q = lane; q < ncb; q += 32) s2 += lp[nrb + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) loss[b] = (float)((double)w1 * (s1 / N) + (double)w2 * (s2 / M));
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// sweep_only: launch just the sweep kernel (partials land in the workspace); bench.py times the dominant kernel
// alone this way for the roofline line.
int chamfer_forward_launch(const float* p1, const float* p2, int B, int N, int M, float* dist1, float* dist2,
                           int64_t* idx1, int64_t* idx2, float* loss, float w1, float w2, void* ws, size_t ws_bytes,
                           cudaStream_t stream, bool sweep_only) {
    if (B < 0 || N < 1 || M < 1) return fail(FPSG_EINVAL, "chamfer: need B >= 0, N >= 1, M >= 1 (got B=%d N=%d M=%d)", B, N, M);
    if (B == 0) return FPSG_OK;
    if (!p1 || !p2 || (!sweep_only && (!idx1 || !idx2))) return fail(FPSG_EINVAL, "chamfer: null p1/p2/idx1/idx2");
    if (B > 65535) return fail(FPSG_EINVAL, "chamfer: B > 65535 not supported in one call");
    if ((long long)N * 3 > 0x7fffffffLL || (long long)M * 3 > 0x7fffffffLL) return fail(FPSG_EINVAL, "chamfer: cloud too large");
    const ChamferPlan pl = make_chamfer_plan(B, N, M);
    if (pl.units > 0x7fffffffLL) return fail(FPSG_EINVAL, "chamfer: B*N*M too large for one call (%lld units)", pl.units);
    if (!ws || ws_bytes < pl.total_bytes)
        return fail(FPSG_EWORKSPACE, "chamfer: workspace of %zu bytes required, %zu given", pl.total_bytes, ws ? ws_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(FPSG_EWORKSPACE, "chamfer: workspace must be 256-byte aligned");
    char* base = static_cast<char*>(ws);
    float* rpval = reinterpret_cast<float*>(base + pl.rpval_off);
    int* rpsub = reinterpret_cast<int*>(base + pl.rpsub_off);
    u64* colpart = reinterpret_cast<u64*>(base + pl.colpart_off);
    double* losspart = loss ? reinterpret_cast<double*>(base + pl.losspart_off) : nullptr;

    const SweepArgs sa{p1, p2, N, M, pl.strips, pl.ups, pl.q, pl.r, pl.maxslots, pl.Mpad, rpval, rpsub, colpart};
    switch (pl.R) {
        case 16:
            chamfer_sweep_kernel<16><<<pl.grid, kSweepThreads, 0, stream>>>(sa);
            break;
        case 8:
            chamfer_sweep_kernel<8><<<pl.grid, kSweepThreads, 0, stream>>>(sa);
            break;
        default:
            chamfer_sweep_kernel<4><<<pl.grid, kSweepThreads, 0, stream>>>(sa);
            break;
    }
    FPSG_CUDA_OK(cudaGetLastError());
    if (sweep_only) return FPSG_OK;
    const FinalizeArgs fa{p1, p2, N, M, pl.R, pl.strips, pl.ups, pl.q, pl.r, pl.maxslots, pl.Mpad, rpval, rpsub, colpart,
                          dist1, dist2, idx1, idx2, losspart, pl.nrb, pl.ncb};
    const dim3 fgrid((unsigned)(pl.nrb + pl.ncb), (unsigned)B);
    chamfer_finalize_kernel<<<fgrid, kFinThreads, 0, stream>>>(fa);
    FPSG_CUDA_OK(cudaGetLastError());
    if (loss) {
        chamfer_loss_kernel<<<B, 32, 0, stream>>>(losspart, pl.nrb, pl.ncb, N, M, w1, w2, loss);
        FPSG_CUDA_OK(cudaGetLastError());
    }
    return FPSG_OK;
}

size_t chamfer_forward_workspace(int B, int N, int M) {
    if (B <= 0 || N < 1 || M < 1) return 0;
    return make_chamfer_plan(B, N, M).total_bytes;
}

}  // namespace fpsg
