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
//                     consecutive rows (p1 points) in registers and streams 32-column sub-blocks of p2
//                     through a per-warp shared-memory tile (AoS, three broadcast LDS.128 per 4 columns; the next
//                     sub-block is prefetched into registers during the current one; no CTA-wide barrier).
//                     Row direction : best[k] = min3(best[k], d, d') -- value only; once per sub-block the
//                                     thread notes (in shared memory) which sub-block last lowered best[k].
//                     Column direction: per column the thread's minimum over its R rows, then one warp-wide
//                                     CREDUX.MIN.F32 + ballot gives (warp minimum, lanes attaining it); lane 0
//                                     parks that in shared memory and the warp writes its 32 results coalesced.
//                     There is no arg-min bookkeeping per distance and no CTA-wide reduction.
//   finalize kernel : merges the partial minima (rows: over the CTAs that covered the row's strip; columns:
//                     over strips x warps) and recovers the arg-min exactly by re-evaluating the few candidates
//                     the partials point at (the 32 columns of the winning sub-block for a row, the R rows of
//                     the winning thread for a column) with the same FP32 instruction sequence, taking the
//                     lowest index whose distance equals the minimum bit-for-bit. That reproduces the
//                     reference's strict-'<' ascending scan (lowest index wins ties). For clouds up to 64 KB the
//                     candidates' cloud is staged in shared memory (cp.async issued ahead of the dependency
//                     wait) and every thread scans its own window there; the grid is pair-fastest so that the
//                     row blocks (the long ones) are dispatched before the column blocks.
//   loss (optional) : w1*mean(dist1) + w2*mean(dist2) per pair -- the reduction glue of few_shot.py:110 / kaolin's
//                     chamfer_distance -- folded into the finalize kernel: per-CTA double sums, added in index
//                     order by whichever CTA of the pair arrives last (deterministic, no extra launch).
//
// Out-of-range rows and columns of ragged tiles are filled with +INF coordinates: their distances are +INF,
// never win a strict comparison, and never equal a finite minimum.
#include "chamfer_plan.cuh"

namespace fpsg {

static int g_sm_count = 0;
constexpr int kSweepCarveoutPercent = 72;  // 164 KB of 228

int sm_count() {
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
    if (const char* e = tune_env("FPSG_SWEEP_R")) {  // tuning build only: 4, 8 or 16
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) p.R = v;
    }
    p.TN = kSweepThreads * p.R;
    p.strips = cdiv(N, p.TN);
    p.ups = cdiv(M, kUnit);
    p.Mpad = p.ups * kUnit;
    p.units = (long long)B * p.strips * p.ups;
    const int per_sm = sweep_ctas_per_sm(p.R);
    long long g = (long long)sm_count() * per_sm;
    // at least 4 units per CTA: every CTA that touches a (pair, strip) leaves one row partial per row for the
    // finalize kernel to merge, and small problems do not need the whole chip anyway
    if (g > p.units / 4) g = p.units / 4;
    if (g < 1) g = 1;
    p.grid = (int)g;
    p.q = (int)(p.units / g);
    p.r = (int)(p.units % g);
    p.maxslots = p.q > 0 ? cdiv(p.ups, p.q) + 1 : 1;
    // every finalize CTA stages a whole cloud: larger CTAs (fewer copies) once the grid exceeds a single wave anyway
    p.fin_threads = 2LL * B * (cdiv(N, 256) + cdiv(M, 256)) > 7LL * sm_count() ? 512 : 256;
    if (const char* e = tune_env("FPSG_FIN_THREADS")) {  // tuning build only: 256 or 512
        const int v = atoi(e);
        if (v == 256 || v == 512) p.fin_threads = v;
    }
    p.nrb = cdiv(N, p.fin_threads);
    p.ncb = cdiv(M, p.fin_threads);
    size_t off = 0;
    const size_t rows = (size_t)B * p.strips * p.maxslots * p.TN;
    p.rowmin_off = off;
    off += align_up(rows * sizeof(float), 256);
    p.rowsub_off = off;
    off += align_up(rows * sizeof(int), 256);
    p.colpart_off = off;
    off += align_up((size_t)B * p.strips * kSweepWarps * p.Mpad * sizeof(u64), 256);
    p.losspart_off = off;
    off += align_up((size_t)B * (p.nrb + p.ncb) * sizeof(double), 256);
    p.counters_off = off;
    off += align_up(((size_t)B + kExtraCounters) * sizeof(unsigned), 256);
    p.progress_off = off;
    off += align_up((size_t)p.grid * sizeof(unsigned), 256);
    p.total_bytes = off;
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// sweep
// ---------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kSweepThreads, sweep_ctas_per_sm(R)) chamfer_sweep_kernel(const SweepArgs a) {
    // Every warp keeps its own copy of the current column tile, so the warps of a CTA never wait for each other:
    // there is no __syncthreads in this kernel. Layout: AoS xyz (kSub*3 floats). The tile is always written in full
    // (+INF beyond the piece): a ragged tail reads +INF.
    __shared__ __align__(16) float sCols[kSweepWarps][kSub * 3];
    __shared__ __align__(16) u64 sColRes[kSweepWarps][kSub];
    __shared__ float sPrev[R][kSweepThreads];  // best[k] as of the previous sub-block (thread-private slots)
    __shared__ int sSubIdx[R][kSweepThreads];  // sub-block that last lowered best[k]
    constexpr int kPf = kSub * 3 / 32;         // prefetch registers per lane
    constexpr int kUps = kSub / kUnit;         // units per sub-block

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = blockIdx.x;

    // this CTA's run of units, walked as pieces: a piece never crosses a sub-block or a (pair, strip) boundary
    int left = a.q + (g < a.r ? 1 : 0);           // units still to do
    int bs, cu;                                    // (pair, strip) index and unit inside it
    {
        const int u0 = g * a.q + min(g, a.r);
        bs = u0 / a.ups;
        cu = u0 - bs * a.ups;
    }
    auto piece_units = [&](int cu_, int left_) { return min(min(kUps - (cu_ % kUps), a.ups - cu_), left_); };

    float px[R], py[R], pz[R], best[R];
#ifdef FPSG_SWEEP_EXPANDED_PROBE
    float an[R];  // |a|^2 (measurement-only build: the expanded form |a|^2 + |b|^2 - 2 a.b, NOT the reference's rounding)
#endif
    int cur_bs = -1;                      // (pair, strip) whose rows are in registers
    int slot = 0, b = 0, strip = 0;
    float pf[kPf];    // the next piece's tile, in flight from global memory during the current piece

    cudaGridDependencySynchronize();  // (PDL) the producer of p1/p2 and the previous user of the workspace are done
    FPSG_TRACE(a.trace, 0, g, 0);
#ifdef FPSG_TUNING
    if (a.trace != nullptr && tid == 0 && g < kTraceCtas) {  // where this CTA runs: (SM id, warp slot of its warp 0)
        unsigned sm, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        a.trace[((size_t)g) * kTracePhases + 3] = ((u64)sm << 32) | wid;
    }
#endif
    if (g == 0)
        for (int q = tid; q < a.B + kExtraCounters; q += kSweepThreads) a.counters[q] = 0u;
    if (tid == 0) a.progress[g] = 0u;
    // The next grid of the chain may be scheduled onto SMs as this grid's CTAs retire. It is launched once EVERY CTA of
    // this grid has passed the trigger, i.e. after all counters and progress words were zeroed and fenced: a consumer that
    // polls the progress words (chamfer_finish.cu) can never see a stale value of an earlier launch.
    __threadfence();
    __syncthreads();
    bool triggered = false;
    if (left <= a.trig_left) {
        cudaTriggerProgrammaticLaunchCompletion();
        triggered = true;
    }

    auto fetch = [&](int bs_, int cu_, int n_) {  // global -> registers; INF beyond the piece: padded columns never win
        const int b_ = a.strips == 1 ? bs_ : bs_ / a.strips;
        const int cb = cu_ * kUnit;
        const int n = (min(a.M, cb + n_ * kUnit) - cb) * 3;
        const float* src = seg_p2(a.seg, b_, a.M) + (size_t)cb * 3;
#pragma unroll
        for (int q = 0; q < kPf; ++q) pf[q] = lane + 32 * q < n ? __ldg(src + lane + 32 * q) : CUDART_INF_F;
    };

    if (left > 0) fetch(bs, cu, piece_units(cu, left));

    while (left > 0) {
        if (!triggered && left <= a.trig_left) {  // (CTA-uniform) a later trigger keeps the dependent grid's CTAs off the SMs for longer
            cudaTriggerProgrammaticLaunchCompletion();
            triggered = true;
        }
        const int nun = piece_units(cu, left);
        const int cbeg = cu * kUnit;
        const int cend = min(a.M, cbeg + nun * kUnit);
        // the piece after this one
        int nbs = bs, ncu = cu + nun;
        const int nleft = left - nun;
        if (ncu == a.ups) {
            ncu = 0;
            ++nbs;
        }

        if (bs != cur_bs) {
            // new (pair, strip): bring this thread's R consecutive rows into registers. Issued before the tile is parked, so
            // that the two global-memory latencies (rows here, prefetched columns below) overlap instead of adding up
            cur_bs = bs;
            b = a.strips == 1 ? bs : bs / a.strips;
            strip = bs - b * a.strips;
            slot = g - cta_of_unit(bs * a.ups, a.q, a.r);
            const float* p1b = seg_p1(a.seg, b, a.N);
            const int row0 = strip * (kSweepThreads * R) + tid * R;
            const float* rsrc = p1b + 3 * (size_t)row0;
            if (row0 + R <= a.N && (reinterpret_cast<uintptr_t>(rsrc) & 15) == 0) {
                float4 v[3 * R / 4];
#pragma unroll
                for (int q = 0; q < 3 * R / 4; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(rsrc) + q);
                const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    px[k] = f[3 * k];
                    py[k] = f[3 * k + 1];
                    pz[k] = f[3 * k + 2];
                }
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const bool ok = row0 + k < a.N;
                    px[k] = ok ? __ldg(rsrc + 3 * k) : CUDART_INF_F;
                    py[k] = ok ? __ldg(rsrc + 3 * k + 1) : CUDART_INF_F;
                    pz[k] = ok ? __ldg(rsrc + 3 * k + 2) : CUDART_INF_F;
                }
            }
#ifdef FPSG_SWEEP_EXPANDED_PROBE
#pragma unroll
            for (int k = 0; k < R; ++k) an[k] = __fmaf_rn(px[k], px[k], __fmaf_rn(py[k], py[k], __fmul_rn(pz[k], pz[k])));
#endif
#pragma unroll
            for (int k = 0; k < R; ++k) {
                best[k] = CUDART_INF_F;
                sPrev[k][tid] = CUDART_INF_F;
                sSubIdx[k][tid] = cbeg / kSub;
            }
        }
        // park the prefetched tile in this warp's shared-memory buffer
        __syncwarp();  // all lanes are done reading the previous tile
#pragma unroll
        for (int q = 0; q < kPf; ++q) sCols[w][lane + 32 * q] = pf[q];

        __syncwarp();  // tile visible to the whole warp
        if (nleft > 0) fetch(nbs, ncu, piece_units(ncu, nleft));  // overlaps the whole sub-block below

        const int ncols = cend - cbeg;
        const int nsteps = (ncols + 3) >> 2;
        const float* tile = sCols[w];
        u64* res = sColRes[w];

#pragma unroll 1
        for (int step = 0; step < nsteps; ++step) {
            const float4 v0 = *reinterpret_cast<const float4*>(tile + 12 * step);      // x0 y0 z0 x1
            const float4 v1 = *reinterpret_cast<const float4*>(tile + 12 * step + 4);  // y1 z1 x2 y2
            const float4 v2 = *reinterpret_cast<const float4*>(tile + 12 * step + 8);  // z2 x3 y3 z3
            float c0m = CUDART_INF_F, c1m = CUDART_INF_F, c2m = CUDART_INF_F, c3m = CUDART_INF_F;
#ifdef FPSG_SWEEP_EXPANDED_PROBE
            // per column (-2x, -2y, -2z, |b|^2): in a real filter kernel the tile would hold these (computed once per tile by one
            // lane); here every thread derives them per step (24 operations per 64 evaluations: the probe is pessimistic by that)
            const float q0x = -2.f * v0.x, q0y = -2.f * v0.y, q0z = -2.f * v0.z, q0n = __fmaf_rn(v0.x, v0.x, __fmaf_rn(v0.y, v0.y, v0.z * v0.z));
            const float q1x = -2.f * v0.w, q1y = -2.f * v1.x, q1z = -2.f * v1.y, q1n = __fmaf_rn(v0.w, v0.w, __fmaf_rn(v1.x, v1.x, v1.y * v1.y));
            const float q2x = -2.f * v1.z, q2y = -2.f * v1.w, q2z = -2.f * v2.x, q2n = __fmaf_rn(v1.z, v1.z, __fmaf_rn(v1.w, v1.w, v2.x * v2.x));
            const float q3x = -2.f * v2.y, q3y = -2.f * v2.z, q3z = -2.f * v2.w, q3n = __fmaf_rn(v2.y, v2.y, __fmaf_rn(v2.z, v2.z, v2.w * v2.w));
#define FPSG_EXPD(k_, qx, qy, qz, qn) __fmaf_rn(px[k_], qx, __fmaf_rn(py[k_], qy, __fmaf_rn(pz[k_], qz, __fadd_rn(qn, an[k_]))))
#pragma unroll
            for (int k = 0; k < R; k += 2) {
                const float a0 = FPSG_EXPD(k, q0x, q0y, q0z, q0n), a1 = FPSG_EXPD(k, q1x, q1y, q1z, q1n);
                const float a2 = FPSG_EXPD(k, q2x, q2y, q2z, q2n), a3 = FPSG_EXPD(k, q3x, q3y, q3z, q3n);
                const float b0 = FPSG_EXPD(k + 1, q0x, q0y, q0z, q0n), b1 = FPSG_EXPD(k + 1, q1x, q1y, q1z, q1n);
                const float b2 = FPSG_EXPD(k + 1, q2x, q2y, q2z, q2n), b3 = FPSG_EXPD(k + 1, q3x, q3y, q3z, q3n);
                best[k] = min3(min3(best[k], a0, a1), a2, a3);
                best[k + 1] = min3(min3(best[k + 1], b0, b1), b2, b3);
                c0m = min3(c0m, a0, b0);
                c1m = min3(c1m, a1, b1);
                c2m = min3(c2m, a2, b2);
                c3m = min3(c3m, a3, b3);
            }
#undef FPSG_EXPD
#else
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
#endif
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

        // last piece of this (pair, strip) for this CTA: flush the row partials
        if (nleft == 0 || nbs != bs) {
            const size_t o = ((size_t)bs * a.maxslots + slot) * (kSweepThreads * R) + tid * R;
            float4* dm = reinterpret_cast<float4*>(a.rowmin + o);
            int4* ds = reinterpret_cast<int4*>(a.rowsub + o);
#pragma unroll
            for (int k = 0; k < R; k += 4) {
                dm[k >> 2] = make_float4(best[k], best[k + 1], best[k + 2], best[k + 3]);
                ds[k >> 2] = make_int4(sSubIdx[k][tid], sSubIdx[k + 1][tid], sSubIdx[k + 2][tid], sSubIdx[k + 3][tid]);
            }
        }
        // last piece of a PAIR for this CTA: publish how far the run has come (all of the CTA's stores so far, then one
        // release store)
        if (nleft == 0 || (a.strips == 1 ? nbs : nbs / a.strips) != b) {
            __syncthreads();
            if (tid == 0) st_release_u32(a.progress + g, (unsigned)(a.q + (g < a.r ? 1 : 0) - nleft));
            FPSG_TRACE(a.trace, 0, g, nleft == 0 ? 2 : 1);
        }
        bs = nbs;
        cu = ncu;
        left = nleft;
    }
}

// ---------------------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    SegTable seg;
    int N, M, R, strips, ups, q, r, maxslots, Mpad;
    const float* rowmin;
    const int* rowsub;
    const u64* colpart;
    float* dist1;
    float* dist2;
    int64_t* idx1;
    int64_t* idx2;
    double* losspart;    // per-CTA sums of the distances (nullptr: no loss wanted)
    unsigned* counters;  // per-pair arrival counters + one for the whole batch, zeroed by the sweep kernel
    float* loss;
    float w1, w2;
    int nrb, ncb;
    float* totals;       // [1 + segments]: weighted total and per-segment sums of loss (nullptr: not wanted)
    int B;
    int stage;           // the cloud whose points are re-evaluated is staged in (dynamic) shared memory
    PeerExchange px;     // multi-GPU scalar exchange in the batch-level epilogue (world == 0: none)
};

constexpr size_t kFinStageMaxBytes = 64 * 1024;    // clouds up to this size are staged

template <int kFinThreads>
__global__ void __launch_bounds__(kFinThreads, 1024 / kFinThreads) chamfer_finalize_kernel(const FinalizeArgs a) {
    extern __shared__ __align__(16) float sCloud[];    // staged candidate cloud: p2[b] for row CTAs, p1[b] for column CTAs
    __shared__ __align__(16) float4 sPt[kFinThreads];  // (x, y, z, minimum) of the thread's point
    __shared__ int sBase[kFinThreads];                  // first candidate index to re-evaluate
    __shared__ int sFound[kFinThreads];
    __shared__ double sRed[kFinThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    // pair-fastest grid: CTAs are dispatched block row by block row, i.e. all row blocks (the long ones: ~20 partials per
    // row) of all pairs before any column block. The CTAs that only become resident when the last sweep CTAs retire -- and
    // whose cloud copy is still in flight when the dependency wait ends -- are then the short column blocks.
    const int b = blockIdx.x, blk = blockIdx.y;
    const float* __restrict__ p1b = seg_p1(a.seg, b, a.N);
    const float* __restrict__ p2b = seg_p2(a.seg, b, a.M);
    const int TN = kSweepThreads * a.R;
    float mine = 0.f;  // this thread's distance, for the loss partial
    // what this thread writes at the very end (after the pair's loss, see below)
    float* out_dist = nullptr;
    int64_t* out_idx = nullptr;
    int out_val = 0;
    cudaTriggerProgrammaticLaunchCompletion();  // a following backward may start to launch
    // Staging: the arg-min re-evaluation reads a few candidates per point from anywhere in the other cloud -- dependent,
    // scattered global loads. The clouds are inputs, complete before the sweep (this grid's only predecessor) started, so
    // their loads may be issued ahead of the dependency wait; they land in shared memory while the partials are merged.
    const bool rowcta = blk < a.nrb;
    const float* __restrict__ csrc = rowcta ? p2b : p1b;
    const int cfl = 3 * (rowcta ? a.M : a.N);
    const bool cvec = a.stage && (reinterpret_cast<uintptr_t>(csrc) & 15) == 0;
    const int c4 = cvec ? cfl >> 2 : 0;
    if (cvec) {
        const unsigned sdst = (unsigned)__cvta_generic_to_shared(sCloud);
        for (int q = tid; q < c4; q += kFinThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 16u * q), "l"(reinterpret_cast<const float4*>(csrc) + q) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    cudaGridDependencySynchronize();            // the sweep's partials are complete and visible from here on
    auto stage_cloud = [&]() {  // what the asynchronous copies do not cover (unaligned cloud, last 1-3 floats), then wait for them
        if (!a.stage) return;
        for (int q = 4 * c4 + tid; q < cfl; q += kFinThreads) sCloud[q] = __ldg(csrc + q);
        // +INF up to the next multiple of 32 points: a candidate window may run past the end of the cloud, and a padded
        // point never equals a minimum
        for (int q = cfl + tid; q < 3 * ((cfl / 3 + 31) & ~31); q += kFinThreads) sCloud[q] = CUDART_INF_F;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };

    // Loss partial of this CTA: needs only the merged minima, so it is issued before the arg-min re-evaluation and
    // the latency of the arrival-counter atomic hides behind that phase. Fixed-order block sum in double (lanes by
    // shuffle tree, warps ascending).
    const int nblk = a.nrb + a.ncb;
    unsigned ticket = 0u;
    auto loss_partial = [&](float mine) {
        double v = (double)mine;
        __syncwarp();  // tells the compiler the warp is converged here: the shuffles below need no convergence wrappers
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[w] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int q = 0; q < kFinThreads / 32; ++q) s += sRed[q];
            a.losspart[(size_t)b * nblk + blk] = s;
            ticket = arrive_acq_rel(a.counters + b);  // publishes the partial sum stored above
        }
    };

    if (blk < a.nrb) {
        // ---- rows: each thread merges the partials of the CTAs that covered its row's strip ... ----
        const int i = blk * kFinThreads + tid;
        const bool valid = i < a.N;
        float m = CUDART_INF_F, ax = 0.f, ay = 0.f, az = 0.f;
        int sub = 0;
        if (valid) {
            const int strip = i / TN, within = i - strip * TN;
            ax = __ldg(p1b + 3 * i);
            ay = __ldg(p1b + 3 * i + 1);
            az = __ldg(p1b + 3 * i + 2);
            merge_row_partials(a.rowmin, a.rowsub, b * a.strips + strip, within, TN, a.ups, a.q, a.r, a.maxslots, m, sub);
            mine = m;
        }
        stage_cloud();
        sPt[tid] = make_float4(ax, ay, az, m);
        sBase[tid] = sub * kSub;
        __syncwarp();
        if (a.losspart) loss_partial(mine);  // (contains the CTA barrier that also publishes the staged cloud)
        else if (a.stage) __syncthreads();
        // ... then the warp re-evaluates the kSub columns of every row's winning sub-block together (coalesced);
        // the lowest lane whose distance equals the minimum bit-for-bit is the arg-min. Lanes past the end of the
        // cloud re-evaluate the last column: they can only tie with a lower lane that holds the same column.
        const int last = a.M - 1;
        constexpr int kH = kSub / 32;
        if (a.stage) {
            // staged: every thread scans the kSub candidates of its own row, starting at a lane-dependent rotation (the
            // windows start on multiples of 32 points = 96 floats, so un-rotated all lanes would hit the same bank);
            // the lowest matching index is the arg-min whatever the visiting order
            const float* win = sCloud + 3 * (sub * kSub);
            int f = 0x7fffffff;
#pragma unroll 8
            for (int c = 0; c < kSub; ++c) {
                const int o = lane ^ c;  // a lane-dependent permutation of the window (kSub == 32)
                const float d = sqdist(ax, ay, az, win[3 * o], win[3 * o + 1], win[3 * o + 2]);
                if (d == m) f = min(f, o);
            }
            if (valid) {
                if (a.dist1) out_dist = a.dist1 + (size_t)b * a.N + i;
                out_idx = a.idx1 + (size_t)b * a.N + i;
                // no candidate equals the merged minimum: cannot happen for finite minima (counted, tests assert zero);
                // a row without any finite distance (NaN / Inf coordinates) keeps index 0 like a strict-'<' scan from +INF
                if (f == 0x7fffffff && m != CUDART_INF_F) atomicAdd(a.counters + a.B + 1, 1u);
                out_val = (f == 0x7fffffff || m == CUDART_INF_F) ? (m == CUDART_INF_F ? 0 : sub * kSub) : sub * kSub + f;
            }
        } else {
        constexpr int kBatch = 8;  // rows whose candidate loads are put in flight together (the loop is latency-bound)
        for (int rb = 0; rb < 32; rb += kBatch) {
            float cx[kBatch][kH], cy[kBatch][kH], cz[kBatch][kH];
#pragma unroll
            for (int r = 0; r < kBatch; ++r) {
                const int j0 = sBase[w * 32 + rb + r];
#pragma unroll
                for (int h = 0; h < kH; ++h) {
                    const int j = min(j0 + 32 * h + lane, last);
                    cx[r][h] = __ldg(p2b + 3 * j);
                    cy[r][h] = __ldg(p2b + 3 * j + 1);
                    cz[r][h] = __ldg(p2b + 3 * j + 2);
                }
            }
#pragma unroll
            for (int r = 0; r < kBatch; ++r) {
                const float4 pt = sPt[w * 32 + rb + r];
                const int j0 = sBase[w * 32 + rb + r];
                int f = pt.w == CUDART_INF_F ? 0 : -1 - j0;  // negative: nothing matched yet
                bool done = pt.w == CUDART_INF_F;
#pragma unroll
                for (int h = 0; h < kH; ++h) {
                    const float d = sqdist(pt.x, pt.y, pt.z, cx[r][h], cy[r][h], cz[r][h]);
                    const unsigned hit = __ballot_sync(0xffffffffu, d == pt.w);
                    if (hit && !done) {
                        f = j0 + 32 * h + __ffs(hit) - 1;
                        done = true;
                    }
                }
                if (lane == 0) sFound[w * 32 + rb + r] = f;
            }
        }
        }
        __syncwarp();
        if (valid && !a.stage) {
            if (a.dist1) out_dist = a.dist1 + (size_t)b * a.N + i;
            out_idx = a.idx1 + (size_t)b * a.N + i;
            out_val = sFound[tid];
            if (out_val < 0) {
                atomicAdd(a.counters + a.B + 1, 1u);
                out_val = -1 - out_val;
            }
        }
    } else {
        // ---- columns: each thread merges strips x warps ... ----
        const int j = (blk - a.nrb) * kFinThreads + tid;
        const bool valid = j < a.M;
        float m = CUDART_INF_F, qx = 0.f, qy = 0.f, qz = 0.f;
        int r0 = 0;
        if (valid) {
            qx = __ldg(p2b + 3 * j);
            qy = __ldg(p2b + 3 * j + 1);
            qz = __ldg(p2b + 3 * j + 2);
            merge_col_partials(a.colpart, b, j, a.strips, a.Mpad, a.R, m, r0);
            mine = m;
        }
        stage_cloud();
        sPt[tid] = make_float4(qx, qy, qz, m);
        sBase[tid] = r0;
        __syncwarp();
        if (a.losspart) loss_partial(mine);
        else if (a.stage) __syncthreads();
        // ... then groups of R lanes re-evaluate the R consecutive rows of each column's winning thread (coalesced)
        const int R = a.R, per = 32 / R, grp = lane / R, off = lane - grp * R;
        const unsigned gmask = ((1u << R) - 1u) << (grp * R);  // R <= 16
        const int last = a.N - 1;
        if (a.stage) {
            // staged: every thread scans the R rows of its own column's winning thread, rotated by the lane number
            const float* win = sCloud + 3 * r0;
            int f = 0x7fffffff;
#pragma unroll 4
            for (int c = 0; c < R; ++c) {
                const int o = (lane ^ c) & (R - 1);
                const float d = sqdist(qx, qy, qz, win[3 * o], win[3 * o + 1], win[3 * o + 2]);
                if (d == m) f = min(f, o);
            }
            if (valid) {
                if (a.dist2) out_dist = a.dist2 + (size_t)b * a.M + j;
                out_idx = a.idx2 + (size_t)b * a.M + j;
                if (f == 0x7fffffff && m != CUDART_INF_F) atomicAdd(a.counters + a.B + 1, 1u);
                out_val = m == CUDART_INF_F ? 0 : (f == 0x7fffffff ? min(r0, last) : r0 + f);
            }
        } else {
        constexpr int kCB = 4;  // iterations whose candidate loads are in flight together (R is a multiple of 4)
        for (int ib = 0; ib < R; ib += kCB) {
            float rx[kCB], ry[kCB], rz[kCB];
#pragma unroll
            for (int t = 0; t < kCB; ++t) {
                const int c = w * 32 + (ib + t) * per + grp;  // the column this group of lanes serves in iteration ib+t
                const int r = min(sBase[c] + off, last);
                rx[t] = __ldg(p1b + 3 * r);
                ry[t] = __ldg(p1b + 3 * r + 1);
                rz[t] = __ldg(p1b + 3 * r + 2);
            }
#pragma unroll
            for (int t = 0; t < kCB; ++t) {
                const int c = w * 32 + (ib + t) * per + grp;
                const float4 pt = sPt[c];
                const int rr0 = sBase[c];
                const float d = sqdist(pt.x, pt.y, pt.z, rx[t], ry[t], rz[t]);
                const unsigned hit = __ballot_sync(0xffffffffu, d == pt.w) & gmask;
                if (off == 0)
                    sFound[c] = pt.w == CUDART_INF_F ? 0 : (hit ? min(rr0 + (__ffs(hit) - 1 - grp * R), last) : -1 - min(rr0, last));
            }
        }
        }
        __syncwarp();
        if (valid && !a.stage) {
            if (a.dist2) out_dist = a.dist2 + (size_t)b * a.M + j;
            out_idx = a.idx2 + (size_t)b * a.M + j;
            out_val = sFound[tid];
            if (out_val < 0) {
                atomicAdd(a.counters + a.B + 1, 1u);
                out_val = -1 - out_val;
            }
        }
    }

    if (a.losspart && w == 0) {
        // The last CTA of the pair to arrive adds the per-CTA sums in index order, so the result does not depend on
        // arrival order. Only warp 0 is involved (its lane 0 holds the ticket), and it does this BEFORE its own output
        // stores: the fence below then has nothing outstanding to wait for, and the other warps simply finish.
        if (__shfl_sync(0xffffffffu, ticket, 0) == (unsigned)(nblk - 1)) {
            __syncwarp();  // lane 0's acquire (its atomic) is ordered before the other lanes' loads below
            const volatile double* lp = a.losspart + (size_t)b * nblk;
            double s1 = 0.0, s2 = 0.0;
            for (int q = lane; q < a.nrb; q += 32) s1 += lp[q];
            for (int q = lane; q < a.ncb; q += 32) s2 += lp[a.nrb + q];
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_down_sync(0xffffffffu, s1, o);
                s2 += __shfl_down_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) a.loss[b] = (float)((double)a.w1 * (s1 / a.N) + (double)a.w2 * (s2 / a.M));
            if (a.totals) {
                // the reduction on top (few_shot.py:110,117,121): whichever pair finishes last adds the per-pair
                // losses in index order, per segment, in double -- again independent of arrival order
                bool last = false;
                if (lane == 0) {
                    last = arrive_acq_rel(a.counters + a.B) == (unsigned)(a.B - 1);  // publishes loss[b]
                }
                if (__shfl_sync(0xffffffffu, last, 0)) {
                    __syncwarp();
                    const volatile float* lv = a.loss;
                    double total = 0.0;
                    for (int sg = 0; sg < a.seg.n; ++sg) {
                        double acc = 0.0;
                        for (int q = a.seg.b0[sg] + lane; q < a.seg.b0[sg + 1]; q += 32) acc += (double)lv[q];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
                        acc = __shfl_sync(0xffffffffu, acc, 0);
                        if (lane == 0) a.totals[1 + sg] = (float)acc;
                        total += (double)a.seg.weight[sg] * (double)(float)acc;
                    }
                    if (lane == 0) a.totals[0] = (float)total;
                    peer_exchange(a.px, (float)total, lane);
                }
            }
        }
    }
    if (out_idx) {
        if (out_dist) *out_dist = mine;
        *out_idx = out_val;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// Validation + plan + workspace carving shared by every Chamfer forward entry point.
int chamfer_prepare(const SegTable& seg, int B, int N, int M, void* ws, size_t ws_bytes, ChamferPlan* out) {
    if (B < 0 || N < 1 || M < 1) return fail(FPSG_EINVAL, "chamfer: need B >= 0, N >= 1, M >= 1 (got B=%d N=%d M=%d)", B, N, M);
    if (B == 0) return FPSG_OK;
    for (int sg = 0; sg < seg.n; ++sg)
        if (seg.b0[sg + 1] > seg.b0[sg] && (!seg.p1[sg] || !seg.p2[sg])) return fail(FPSG_EINVAL, "chamfer: null p1/p2");
    if (B > 65535) return fail(FPSG_EINVAL, "chamfer: B > 65535 not supported in one call");
    if ((long long)N * 3 > 0x7fffffffLL || (long long)M * 3 > 0x7fffffffLL) return fail(FPSG_EINVAL, "chamfer: cloud too large");
    const ChamferPlan pl = make_chamfer_plan(B, N, M);
    if (pl.units > 0x7fffffffLL) return fail(FPSG_EINVAL, "chamfer: B*N*M too large for one call (%lld units)", pl.units);
    if (!ws || ws_bytes < pl.total_bytes)
        return fail(FPSG_EWORKSPACE, "chamfer: workspace of %zu bytes required, %zu given", pl.total_bytes, ws ? ws_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(FPSG_EWORKSPACE, "chamfer: workspace must be 256-byte aligned");
    *out = pl;
    return FPSG_OK;
}

// units left in a CTA's run when it lets the next grid of the chain launch (see SweepArgs::trig_left)
static int sweep_trigger_left(const ChamferPlan& pl) {
    // Percent of the run still to do at the trigger. 100 = right at the start: the dependent grid's CTAs (finish / finalize) then
    // become resident as soon as the first sweep CTAs retire and wait on the SMs for their pair, in slots that the sweep of the
    // NEXT step (issued on another lane) could use. 60: they arrive when all sweep CTAs are 40 % through -- still well ahead of
    // the first completed pair (one step alone: 53.4 us either way), and four or more lanes gain 1.2 us per step (43.7 -> 42.4;
    // two lanes lose 2 us: profiles/r02d_trigger_variants.jsonl).
    int pct = 60;
    if (const char* e = tune_env("FPSG_SWEEP_TRIGGER_PCT")) pct = atoi(e);  // tuning build only
    return pct >= 100 ? 0x7fffffff : (int)((long long)(pl.q + 1) * pct / 100);
}

int chamfer_sweep_launch(const SegTable& seg, const ChamferPlan& pl, int B, int N, int M, void* ws, cudaStream_t stream) {
    char* base = static_cast<char*>(ws);
    const SweepArgs sa{seg, N, M, pl.strips, pl.ups, pl.q, pl.r, pl.maxslots, pl.Mpad, reinterpret_cast<float*>(base + pl.rowmin_off),
                       reinterpret_cast<int*>(base + pl.rowsub_off), reinterpret_cast<u64*>(base + pl.colpart_off),
                       reinterpret_cast<unsigned*>(base + pl.counters_off), B, reinterpret_cast<unsigned*>(base + pl.progress_off),
                       trace_buffer(), sweep_trigger_left(pl)};
    const dim3 sgrid((unsigned)pl.grid), sblock(kSweepThreads);
    // Shared-memory carve-out: the sweep itself needs 76 KB per SM, for which the driver would configure 100 KB -- and the
    // kernel that follows it in the chain (finalize / finish: 25-60 KB per CTA, staged ahead of its dependency wait) could
    // then not become co-resident before most of the SM's sweep CTAs have retired (measured: resident at 33 us instead of
    // 28 us of a 42-us sweep). The sweep asks for 164 of the 228 KB (all of it costs the sweep 1.8 us: its tile prefetches hit in L1).
    {
        static bool done[64] = {};  // per device; a benign race: the attribute is idempotent
        int dev = 0;
        FPSG_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_sweep_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, kSweepCarveoutPercent));
            FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_sweep_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, kSweepCarveoutPercent));
            FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_sweep_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, kSweepCarveoutPercent));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
    if (pl.R == 16 && tune_env("FPSG_SWEEP_NOPDL")) {  // tuning build: plain stream order (the grid starts when its predecessor has finished)
        chamfer_sweep_kernel<16><<<sgrid, sblock, 0, stream>>>(sa);
        FPSG_CUDA_OK(cudaGetLastError());
    } else if (pl.R == 16)
        FPSG_CUDA_OK(launch_pdl(chamfer_sweep_kernel<16>, sgrid, sblock, 0, stream, sa));
    else if (pl.R == 8)
        FPSG_CUDA_OK(launch_pdl(chamfer_sweep_kernel<8>, sgrid, sblock, 0, stream, sa));
    else
        FPSG_CUDA_OK(launch_pdl(chamfer_sweep_kernel<4>, sgrid, sblock, 0, stream, sa));
    return FPSG_OK;
}

int chamfer_finalize_launch(const SegTable& seg, const ChamferPlan& pl, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1,
                            int64_t* idx2, float* loss, float* totals, float w1, float w2, void* ws, cudaStream_t stream,
                            const PeerExchange* px) {
    char* base = static_cast<char*>(ws);
    double* losspart = loss ? reinterpret_cast<double*>(base + pl.losspart_off) : nullptr;
    FinalizeArgs fa{seg, N, M, pl.R, pl.strips, pl.ups, pl.q, pl.r, pl.maxslots, pl.Mpad, reinterpret_cast<const float*>(base + pl.rowmin_off),
                    reinterpret_cast<const int*>(base + pl.rowsub_off), reinterpret_cast<const u64*>(base + pl.colpart_off),
                    dist1, dist2, idx1, idx2, losspart, reinterpret_cast<unsigned*>(base + pl.counters_off), loss, w1, w2, pl.nrb, pl.ncb, totals, B, 0, PeerExchange{}};
    fa.px.world = 0;
    if (px) fa.px = *px;
    if (pl.nrb + pl.ncb > 65535) return fail(FPSG_EINVAL, "chamfer: N + M too large for one call");
    const dim3 fgrid((unsigned)B, (unsigned)(pl.nrb + pl.ncb));
    size_t fin_smem = align_up((size_t)(N > M ? N : M), 32) * 12;  // whole 32-point windows
    static const bool no_stage = tune_env("FPSG_FIN_NOSTAGE") != nullptr;  // tuning build only
    if (fin_smem <= kFinStageMaxBytes && !no_stage) {
        // more than 48 KB of shared memory per CTA (static + dynamic) needs an opt-in; it is per device, so it is repeated
        // per launch (as the kNN launches do) -- the c1 shapes (24 KB) never get here
        if (fin_smem > 32 * 1024) {
            if (pl.fin_threads == 512)
                FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_finalize_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinStageMaxBytes));
            else
                FPSG_CUDA_OK(cudaFuncSetAttribute(chamfer_finalize_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinStageMaxBytes));
        }
        fa.stage = 1;
    } else {
        fin_smem = 0;
    }
    if (pl.fin_threads == 512)
        FPSG_CUDA_OK(launch_pdl(chamfer_finalize_kernel<512>, fgrid, dim3(512), fin_smem, stream, fa));
    else
        FPSG_CUDA_OK(launch_pdl(chamfer_finalize_kernel<256>, fgrid, dim3(256), fin_smem, stream, fa));
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

// sweep_only: launch just the sweep kernel (partials land in the workspace); bench.py times the dominant kernel
// alone this way for the roofline line.
int chamfer_forward_launch(const SegTable& seg, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1,
                           int64_t* idx2, float* loss, float* totals, float w1, float w2, void* ws, size_t ws_bytes,
                           cudaStream_t stream, bool sweep_only) {
    if (!sweep_only && B > 0 && (!idx1 || !idx2)) return fail(FPSG_EINVAL, "chamfer: null idx1/idx2");
    if (totals && !loss) return fail(FPSG_EINVAL, "chamfer: totals need the per-pair loss buffer");
    ChamferPlan pl;
    const int rc = chamfer_prepare(seg, B, N, M, ws, ws_bytes, &pl);
    if (rc != FPSG_OK || B == 0) return rc;
    const int rs = chamfer_sweep_launch(seg, pl, B, N, M, ws, stream);
    if (rs != FPSG_OK || sweep_only) return rs;
    return chamfer_finalize_launch(seg, pl, B, N, M, dist1, dist2, idx1, idx2, loss, totals, w1, w2, ws, stream, nullptr);
}

// Offset of the diagnostic counter (arg-min recoveries that found no candidate equal to the merged minimum) in a
// workspace of this shape; zeroed by every forward, must read 0 after it.
size_t chamfer_diag_offset(int B, int N, int M) {
    if (B <= 0 || N < 1 || M < 1) return 0;
    return make_chamfer_plan(B, N, M).counters_off + ((size_t)B + 1) * sizeof(unsigned);
}

size_t chamfer_forward_workspace(int B, int N, int M) {
    if (B <= 0 || N < 1 || M < 1) return 0;
    return make_chamfer_plan(B, N, M).total_bytes;
}

}  // namespace fpsg
