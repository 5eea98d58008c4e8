// Work decomposition and workspace layout shared by the Chamfer kernels (chamfer_fwd.cu: sweep + finalize,
// chamfer_finish.cu: fused finalize + backward).
#pragma once
#include "common.cuh"

namespace fpsg {

typedef unsigned long long u64;

constexpr int kSweepThreads = 128;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kUnit = 16;  // columns per scheduling unit
constexpr int kSub = 32;   // columns per shared-memory sub-block (2 units)
static_assert(kSub == 32, "the arg-min recovery permutes a row's window with lane ^ c");

// resident sweep CTAs per SM (register-limited): the persistent grid is this times the SM count
#ifdef FPSG_SWEEP_EXPANDED_PROBE
// measurement-only build (bench_micro/expanded_probe.py): 16 more registers per thread for |a|^2 -> three CTAs per SM
constexpr int sweep_ctas_per_sm(int R) { return R >= 16 ? FPSG_SWEEP_EXPANDED_PROBE : 6; }
#else
constexpr int sweep_ctas_per_sm(int R) { return R >= 16 ? 4 : 6; }
#endif

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
    int fin_threads;  // threads per finalize CTA
    int nrb, ncb;     // finalize CTAs per pair: row blocks, column blocks
    size_t rowmin_off, rowsub_off, colpart_off, losspart_off, counters_off, progress_off, total_bytes;
};

// counters region of the workspace: [0, B) per-pair arrival counters, [B] batch-level arrival counter,
// [B + 1] diagnostic: arg-min recoveries that found no candidate equal to the merged minimum (must stay 0)
constexpr int kExtraCounters = 2;
// progress region: one word per sweep CTA = the units of its run it has completed AND published. Every sweep CTA zeroes
// its own word before it lets the next grid of the chain launch (the launch-completion trigger), so a consumer that was
// launched programmatically behind the sweep can poll the words of the CTAs that cover its pair instead of waiting for
// the whole sweep grid (chamfer_finish.cu): pairs are laid out pair-major over the CTAs and the hardware retires the
// resident CTAs of an SM one after the other, so most pairs are complete long before the grid is.

ChamferPlan make_chamfer_plan(int B, int N, int M);
int sm_count();

struct SweepArgs {
    SegTable seg;   // the clouds: p1 (B,N,3), p2 (B,M,3), possibly in several separately allocated segments
    int N, M, strips, ups, q, r, maxslots, Mpad;
    float* rowmin;  // [B*strips][maxslots][TN]   row partial minima, one slot per CTA that covered the (pair, strip) ...
    int* rowsub;    // [B*strips][maxslots][TN]   ... and the sub-block that produced each (two arrays: the merge reads the
                    //                            4-byte minima and then one sub-block index per row)
    u64* colpart;   // [B*strips][4 warps][Mpad]  (min bits << 32) | ballot of lanes attaining it
    unsigned* counters;  // [B + kExtraCounters], zeroed by the sweep kernel
    int B;
    unsigned* progress;         // [grid] see above
    unsigned long long* trace;  // timeline stamps (tuning builds), else nullptr
    int trig_left;              // the launch-completion trigger is executed once at most this many units of the run are left
                                // (>= the run length: right away, after the progress word was zeroed)
};

// Scalar all-reduce through peer memory (multi-GPU, one process per GPU; include/fpsg_b200.h fpsg_peer_exchange_t): every
// rank owns a small table of 64-bit words [kPeerRing][FPSG_MAX_PEERS] that all ranks can address (NVLink peer mapping).
// The batch-level epilogue of a loss step publishes (seq << 32 | float bits of the local sum) into slot [seq % ring][rank]
// of EVERY rank's table with one 8-byte store each -- value and flag travel in one word, so no fence is needed -- and then
// sums, in rank order and in double, the `world` words of step seq - lag in its own table, polling until all have arrived.
struct PeerExchange {
    u64* table[FPSG_MAX_PEERS];  // table[p] = rank p's table as mapped into this process; nullptr entries beyond world
    float* global_out;           // [kPeerRing] device: global_out[(seq - lag) % ring] = sum over ranks of that step
    int rank, world;
    unsigned seq;                // 1, 2, 3, ... one per step, identical on all ranks
    int lag;                     // 0: wait for this step's values (a full all-reduce inside the step); L >= 1: for those of step seq - L
};
constexpr int kPeerRing = FPSG_PEER_RING;

#ifdef __CUDACC__
// one warp: sums, in rank order and in double, the `world` words of step `want` in this rank's own table, polling until
// all have arrived, and writes global_out[want % ring]
__device__ __forceinline__ void peer_collect(const PeerExchange& px, unsigned want, int lane) {
    float v = 0.f;
    if (lane < px.world) {
        const u64* src = px.table[px.rank] + (want % kPeerRing) * FPSG_MAX_PEERS + lane;
        u64 got;
        long long spins = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
        } while ((unsigned)(got >> 32) != want && ++spins < (1ll << 28));  // a dead peer must not hang the GPU for ever
        v = (unsigned)(got >> 32) == want ? __uint_as_float((unsigned)got) : CUDART_NAN_F;
    }
    double acc = 0.0;
    for (int r = 0; r < px.world; ++r) acc += (double)__shfl_sync(0xffffffffu, v, r);  // rank order: the same bits on every rank
    if (lane == 0) px.global_out[want % kPeerRing] = (float)acc;
}

// executed by one warp (the batch-level last arriver); `total` valid in lane 0
__device__ __forceinline__ void peer_exchange(const PeerExchange& px, float total, int lane) {
    if (px.world <= 0) return;
    total = __shfl_sync(0xffffffffu, total, 0);
    const u64 word = ((u64)px.seq << 32) | (u64)__float_as_uint(total);
    if (lane < px.world) {
        u64* dst = px.table[lane] + (px.seq % kPeerRing) * FPSG_MAX_PEERS + px.rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    }
    if (px.seq <= (unsigned)px.lag) return;  // the first `lag` steps: nothing to collect yet
    const unsigned want = px.seq - (unsigned)px.lag;
    peer_collect(px, want, lane);
}

// which CTA owns unit u under the q/r partition
__device__ __forceinline__ int cta_of_unit(int u, int q, int r) {
    const int cut = r * (q + 1);
    return u < cut ? u / (q + 1) : r + (u - cut) / q;
}

// Arrival counters of the loss reductions. The increment is an acq_rel atomic: it releases the arriving thread's
// earlier stores (no MEMBAR.SC in front of it) and, for the last arriver, acquires everybody else's -- at the place
// where the atomic is issued, i.e. early and off the critical path, instead of a fence in front of the final loads.
__device__ __forceinline__ unsigned arrive_acq_rel(unsigned* counter) {
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// first unit and number of units of sweep CTA g under the q/r partition
__device__ __forceinline__ int cta_first_unit(int g, int q, int r) { return g * q + min(g, r); }

// Blocks (warp 0 polls, then a CTA barrier) until every sweep CTA that covers pair b has published the units of that
// pair: acquire loads of the progress words, one CTA per lane.
__device__ __forceinline__ void wait_pair_swept(const unsigned* progress, int b, int upp, int q, int r) {
    if (threadIdx.x < 32) {
        const int u0 = b * upp, u1 = u0 + upp;  // the pair's units
        const int g0 = cta_of_unit(u0, q, r), g1 = cta_of_unit(u1 - 1, q, r);
        for (int g = g0 + (int)threadIdx.x; g <= g1; g += 32) {
            const int s = cta_first_unit(g, q, r);
            const unsigned need = (unsigned)(min(s + q + (g < r ? 1 : 0), u1) - s);
            while (ld_acquire_u32(progress + g) < need) __nanosleep(20);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float warp_min_f32(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

// Row merge shared by the finalize and the finish kernels: minimum over the slots of the CTAs that covered the
// row's (pair, strip), in ascending slot (= column) order under strict '<', and the sub-block that produced it.
// Slots are in ascending column order, so the first slot that attains the minimum holds the lowest sub-block among
// equal minima: the reference's strict-'<' ascending scan at sub-block granularity. All 4-byte minima of a batch
// are in flight together (one L2 round trip; slots past the last one re-read it), then one dependent load fetches
// the winner's sub-block index.
__device__ __forceinline__ void merge_row_partials(const float* __restrict__ rowmin, const int* __restrict__ rowsub, int bs, int within,
                                                   int TN, int ups, int q, int r, int maxslots, float& m, int& sub) {
    const int u0 = bs * ups;
    const int g0 = cta_of_unit(u0, q, r), g1 = cta_of_unit(u0 + ups - 1, q, r);
    const size_t ro = (size_t)bs * maxslots * TN + within;
    const float* rm = rowmin + ro;
    const int nsl = g1 - g0 + 1;
    int ms = 0;
    m = CUDART_INF_F;
    constexpr int kSlotBatch = 24;
    for (int s0 = 0; s0 < nsl; s0 += kSlotBatch) {
        float v[kSlotBatch];
#pragma unroll
        for (int s = 0; s < kSlotBatch; ++s) v[s] = __ldcg(rm + (unsigned)(min(s0 + s, nsl - 1) * TN));  // < maxslots * TN: 32 bits
#pragma unroll
        for (int s = 0; s < kSlotBatch; ++s) {
            if (v[s] < m) {
                m = v[s];
                ms = min(s0 + s, nsl - 1);
            }
        }
    }
    sub = __ldcg(rowsub + ro + (unsigned)(ms * TN));
}

// Column merge: minimum over strips x warps in ascending (strip, warp) = ascending row order; r0 = first row of the
// winning thread's R consecutive rows.
__device__ __forceinline__ void merge_col_partials(const u64* __restrict__ colpart, int b, int j, int strips, int Mpad, int R, float& m, int& r0) {
    m = CUDART_INF_F;
    r0 = 0;
    for (int s = 0; s < strips; ++s) {
#pragma unroll
        for (int ww = 0; ww < kSweepWarps; ++ww) {
            const u64 rec = __ldcg(colpart + (((size_t)b * strips + s) * kSweepWarps + ww) * Mpad + j);
            const float v = __uint_as_float((unsigned)(rec >> 32));
            if (v < m) {
                m = v;
                r0 = (s * kSweepThreads + ww * 32 + ((unsigned)rec ? __ffs((unsigned)rec) - 1 : 0)) * R;
            }
        }
    }
}
#endif

}  // namespace fpsg
