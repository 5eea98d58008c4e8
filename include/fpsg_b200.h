/*
 * fpsg_b200.h -- C ABI of libfpsg_b200.so: the B200 (sm_100a) implementation of FPSG's point-set-distance
 * hot path. Plain pointers and sizes only; no torch types. All pointers are DEVICE pointers unless a name
 * ends in _host. Every function is re-entrant, enqueues work on `stream` and returns without synchronising
 * (CUDA-graph capturable); the caller owns every buffer, including the workspace, and makes the buffers'
 * device current before calling (torch's device guard does). Outputs are fully overwritten; nothing needs
 * pre-zeroing. Return value: 0 on success, negative FPSG_E* on failure with a message in fpsg_last_error().
 *
 * What each entry point replaces in the reference (voidstrike/FPSG, paths relative to /root/reference):
 *
 *   fpsg_chamfer_fwd / fpsg_chamfer_bwd
 *       the forward / backward of `kaolin.metrics.pointcloud.sided_distance` x2 inside
 *       `chamfer_distance`, imported at src/models/few_shot.py:13, bound at :57, called at :110, :117, :167
 *       (Kaolin 0.9.0's `_C.metrics.sided_distance_forward_cuda` / `_backward_cuda`).
 *   fpsg_chamfer_loss_fwd / fpsg_chamfer_loss_bwd
 *       the same plus the reduction glue `w1*mean + w2*mean` and the uniform upstream gradient produced by
 *       `.sum()` at src/models/few_shot.py:110,117.
 *   fpsg_chamfer_multi_loss_fwd / fpsg_chamfer_multi_loss_bwd
 *       both Chamfer calls of one training step (query set :110, support set :117), their `.sum()`s and the
 *       weighted total (:121-122) in one launch chain.
 *   fpsg_knn_fwd
 *       `knn(x, k)` at src/dgcnn/model.py:13-20 (called from get_graph_feature, :28).
 *   fpsg_edge_feature_fwd / fpsg_edge_feature_bwd
 *       the gather / repeat / cat / permute tail of `get_graph_feature`, src/dgcnn/model.py:31-41, and its
 *       autograd backward.
 *
 *   fpsg_emd_sinkhorn_fwd
 *       the evaluation metric `emd_wrapper` (src/models/utils.py:9-13, used at src/models/few_shot.py:168): entropic
 *       optimal transport by log-domain Sinkhorn iterations (parity unpinned, see below).
 *
 *   fpsg_normalize_clouds / fpsg_ply_read_ascii_host
 *       the host input pipeline in front of the path: `ply_reader` (src/datasets/modelnet.py:15-29) and the
 *       centre / scale block of `FewShotSubModelNet.__getitem__` (:66-69).
 *
 * INTEGRATION.md shows the reference-side bindings (ctypes) a maintainer would add.
 */
#ifndef FPSG_B200_H_
#define FPSG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPSG_ABI_VERSION 1

#define FPSG_OK 0
#define FPSG_EINVAL (-1)     /* bad shape / null pointer / unsupported size */
#define FPSG_EWORKSPACE (-2) /* workspace missing or too small              */
#define FPSG_ECUDA (-3)      /* a CUDA runtime call failed                  */

typedef void* fpsg_stream_t; /* a cudaStream_t */

int fpsg_abi_version(void);
/* Thread-local, valid until the next failing call on the same thread. Never NULL. */
const char* fpsg_last_error(void);

/* ---- Chamfer / sided distance, squared L2, FP32 ------------------------------------------------------- */
/* p1 (B,N,3), p2 (B,M,3) row-major FP32. dist1 (B,N), idx1 (B,N) int64: nearest p2 point of each p1 point;
 * dist2 (B,M), idx2 (B,M): nearest p1 point of each p2 point. Lowest index wins exact ties. N, M >= 1. */
size_t fpsg_chamfer_fwd_workspace_bytes(int B, int N, int M);
int fpsg_chamfer_fwd(const float* p1, const float* p2, int B, int N, int M, float* dist1, float* dist2,
                     int64_t* idx1, int64_t* idx2, void* ws, size_t ws_bytes, fpsg_stream_t stream);

/* g1 (B,N), g2 (B,M): upstream gradients of dist1 / dist2. gp1 (B,N,3) and gp2 (B,M,3) are each optional
 * (NULL = not requested, as for the ground-truth cloud in the reference's training loss). */
size_t fpsg_chamfer_bwd_workspace_bytes(int B, int N, int M);
int fpsg_chamfer_bwd(const float* p1, const float* p2, const int64_t* idx1, const int64_t* idx2, const float* g1,
                     const float* g2, int B, int N, int M, float* gp1, float* gp2, void* ws, size_t ws_bytes,
                     fpsg_stream_t stream);

/* Fused loss: loss[b] = w1 * mean_i dist1[b,i] + w2 * mean_j dist2[b,j]  (what kaolin's chamfer_distance
 * returns). dist1/dist2 are optional outputs here (NULL = not wanted); idx1/idx2 are required (the backward
 * consumes them). Uses the same workspace size as fpsg_chamfer_fwd. */
int fpsg_chamfer_loss_fwd(const float* p1, const float* p2, int B, int N, int M, float w1, float w2, float* loss,
                          float* dist1, float* dist2, int64_t* idx1, int64_t* idx2, void* ws, size_t ws_bytes,
                          fpsg_stream_t stream);
/* gloss (B,): upstream gradient of loss (required); equivalent to fpsg_chamfer_bwd with g1 = gloss*w1/N, g2 = gloss*w2/M. */
int fpsg_chamfer_loss_bwd(const float* p1, const float* p2, const int64_t* idx1, const int64_t* idx2,
                          const float* gloss, float w1, float w2, int B, int N, int M, float* gp1, float* gp2,
                          void* ws, size_t ws_bytes, fpsg_stream_t stream);

/* Several separately allocated runs of pairs ("segments") in ONE launch chain, plus the reduction the reference
 * applies on top: the query-set and support-set calls of the training loss and their weighted sum,
 *     loss_rec_q = pc_metric(syn_q, ref_q).sum(); loss_rec_s = pc_metric(syn_s, ref_s).sum()
 *     ttl_loss   = query_factor * loss_rec_q + support_factor * loss_rec_s      (src/models/few_shot.py:109-122)
 * (the dual-call batching half). All segments share N and M. `segs_host` is a HOST array
 * read during the call. loss (Btot,) per-pair values in segment order, Btot = sum of batch; totals (1 + nseg,):
 * totals[0] = sum_s weight_s * sum_{b in s} loss[b], totals[1 + s] = sum_{b in s} loss[b]; idx1 (Btot,N),
 * idx2 (Btot,M). Workspace: fpsg_chamfer_fwd_workspace_bytes(Btot, N, M). */
#define FPSG_MAX_SEGMENTS 8
typedef struct fpsg_chamfer_segment {
    const float* p1; /* (batch, N, 3) */
    const float* p2; /* (batch, M, 3) */
    float* gp1;      /* backward only: (batch, N, 3), or NULL = not requested */
    float* gp2;      /* backward only: (batch, M, 3), or NULL = not requested */
    int batch;       /* >= 0 */
    float weight;    /* factor of this segment's summed loss in totals[0] */
} fpsg_chamfer_segment_t;
int fpsg_chamfer_multi_loss_fwd(const fpsg_chamfer_segment_t* segs_host, int nseg, int N, int M, float w1, float w2,
                                float* loss, float* totals, int64_t* idx1, int64_t* idx2, void* ws, size_t ws_bytes,
                                fpsg_stream_t stream);
/* gtotal: DEVICE scalar, the upstream gradient of totals[0]; every pair of segment s receives
 * gloss = gtotal * weight_s, then as fpsg_chamfer_loss_bwd. Gradients go to each segment's gp1 / gp2. */
int fpsg_chamfer_multi_loss_bwd(const fpsg_chamfer_segment_t* segs_host, int nseg, const int64_t* idx1,
                                const int64_t* idx2, const float* gtotal, float w1, float w2, int N, int M,
                                fpsg_stream_t stream);

/* Loss forward AND backward in one call, for a loss whose upstream gradient is known when the step is launched --
 * the reference's training step: `loss = pc_metric(syn, ref).sum()` (src/models/few_shot.py:110,117), weighted total
 * (:121), `ttl_loss.backward()` (src/trainNetwork.py:144): the gradient that reaches `loss` is a constant.
 * gloss (B,) DEVICE, the upstream gradient of loss, or NULL = ones (the `.sum()` case). Outputs as
 * fpsg_chamfer_loss_fwd (dist1/dist2 optional, idx1/idx2 required) plus gp1 (B,N,3) / gp2 (B,M,3), each optional.
 * Results are identical to fpsg_chamfer_loss_fwd followed by fpsg_chamfer_loss_bwd; for clouds of up to 4096 points
 * the merge, the arg-min recovery, the loss and the backward of a pair run in ONE kernel (a thread-block cluster per
 * pair, inverted index through distributed shared memory), larger clouds take the three-kernel chain.
 * Workspace: fpsg_chamfer_fwd_workspace_bytes(B, N, M). */
int fpsg_chamfer_loss_fwd_bwd(const float* p1, const float* p2, int B, int N, int M, float w1, float w2,
                              const float* gloss, float* loss, float* dist1, float* dist2, int64_t* idx1,
                              int64_t* idx2, float* gp1, float* gp2, void* ws, size_t ws_bytes, fpsg_stream_t stream);
/* The multi-segment form: fpsg_chamfer_multi_loss_fwd + fpsg_chamfer_multi_loss_bwd in one call. gtotal: DEVICE
 * scalar, the upstream gradient of totals[0], or NULL = 1 (`ttl_loss.backward()`). Gradients go to each segment's
 * gp1 / gp2 (NULL = not requested). */
int fpsg_chamfer_multi_loss_fwd_bwd(const fpsg_chamfer_segment_t* segs_host, int nseg, int N, int M, float w1,
                                    float w2, const float* gtotal, float* loss, float* totals, int64_t* idx1,
                                    int64_t* idx2, void* ws, size_t ws_bytes, fpsg_stream_t stream);

/* ---- Multi-GPU: the scalar loss all-reduce through NVLink peer memory ------------------------------------ */
/* One process per GPU, the batch sharded over the ranks (pairs are independent); the only exchange of the path is the
 * scalar the reference logs, sum_b loss[b] (src/models/few_shot.py:110,117,121). Instead of a collective launched from
 * the host after the step, the loss kernel's own epilogue writes the rank's partial sum into every rank's table and
 * sums what the other ranks wrote: `table[p]` is the device address, IN THIS PROCESS, of rank p's table of
 * FPSG_PEER_RING * FPSG_MAX_PEERS 64-bit words (zero-initialised once; mapped through CUDA IPC / symmetric memory by
 * the host code, see fpsg_b200/sharded.py). `seq` numbers the steps 1, 2, 3, ... identically on all ranks. With
 * lag = 0 the epilogue waits for this step's values of all ranks -- a complete all-reduce inside the step; with
 * lag = L >= 1 it collects the values of step seq - L, which have long arrived, so that no rank ever waits for another
 * (the reduced values of the last L steps are then fetched with fpsg_peer_exchange_collect). Steps that are issued
 * round-robin over L streams (independent batches overlapping on the device) use lag = L: the step they collect ran
 * earlier on the SAME stream, so its own contribution is certainly published; L <= FPSG_PEER_RING / 4.
 * global_out (FPSG_PEER_RING floats, device): global_out[(seq - lag) % FPSG_PEER_RING] = sum over the ranks, added
 * in rank order (bit-identical on every rank). Pass px = NULL for a single-GPU step. */
#define FPSG_MAX_PEERS 8
#define FPSG_PEER_RING 32
typedef struct fpsg_peer_exchange {
    unsigned long long* table[FPSG_MAX_PEERS];
    float* global_out;
    int rank, world;
    unsigned seq;
    int lag;
} fpsg_peer_exchange_t;
/* fpsg_chamfer_multi_loss_fwd_bwd plus the exchange above in the same launch chain (no additional kernel). */
int fpsg_chamfer_multi_loss_fwd_bwd_dist(const fpsg_chamfer_segment_t* segs_host, int nseg, int N, int M, float w1,
                                         float w2, const float* gtotal, float* loss, float* totals, int64_t* idx1,
                                         int64_t* idx2, void* ws, size_t ws_bytes, const fpsg_peer_exchange_t* px,
                                         fpsg_stream_t stream);

/* Collects step px->seq without publishing anything (the flush after the last step of a lag = 1 sequence): a one-warp
 * kernel that waits for the `world` values of that step and writes global_out[seq % FPSG_PEER_RING]. px->lag is ignored. */
int fpsg_peer_exchange_collect(const fpsg_peer_exchange_t* px, fpsg_stream_t stream);

/* ---- Host-buffer streaming step ------------------------------------------------------------------------- */
/* The reference's training loop around the loss (src/trainNetwork.py:141-148): `to_cuda(sample)`, `model.loss(sample)`,
 * `ttl_loss.backward()`, `.item()` -- per batch: H2D copy of p1 (B,N,3) and p2 (B,M,3) from HOST memory (pinned memory
 * makes the copies asynchronous), the fused loss step (fpsg_chamfer_multi_loss_fwd_bwd with one segment: per-pair loss,
 * batch sum, gradient w.r.t. p1), D2H copy of the losses. `submit` only enqueues (copies on a copy stream, kernels on a
 * compute stream, chained by events) and returns; `result` waits for the OLDEST submitted batch. Every one of the `depth`
 * slots is a LANE of its own -- its own compute stream, workspace and outputs -- so with depth >= 2 the copies of batch
 * i+1 overlap the kernels of batch i, and the latency-bound tail of batch i (merge, arg-min recovery, backward) overlaps
 * the distance sweep of batch i+1 (measured at B=32, N=M=2048: 56 us per batch on one lane, 44.6 us on three).
 * (Run the process with CUDA_DEVICE_MAX_CONNECTIONS >= the number of streams in use, e.g. 32: with the default of 8 hardware
 * queues the lanes may alias onto shared queues and serialise.)
 * Up to `depth` (1..8; 1..4 with an exchange) batches may be in flight; the host buffers of
 * a batch must stay valid until its result was taken. The caller owns all memory: a 256-byte aligned DEVICE arena and
 * a PINNED HOST arena of the sizes fpsg_chamfer_pipeline_arena_bytes reports, both alive as long as the handle; the
 * handle itself holds 1 + depth CUDA streams and 2*depth events, and is used by one host thread at a time.
 * result: *loss_sum_host = sum_b loss[b]; loss_host (B,) optional (NULL); *gp1_device / *idx1_device / *idx2_device
 * (each optional) point INTO the device arena: gradient (B,N,3) and arg-min indices of that batch, complete when the
 * call returns and valid until `depth` further batches have been submitted. */
typedef struct fpsg_chamfer_pipeline fpsg_chamfer_pipeline_t;
int fpsg_chamfer_pipeline_arena_bytes(int B, int N, int M, int depth, size_t* device_bytes, size_t* pinned_host_bytes);
int fpsg_chamfer_pipeline_create(int B, int N, int M, float w1, float w2, int depth, void* device_arena,
                                 size_t device_bytes, void* pinned_host_arena, size_t pinned_host_bytes,
                                 fpsg_chamfer_pipeline_t** out);
int fpsg_chamfer_pipeline_submit(fpsg_chamfer_pipeline_t* p, const float* p1_host, const float* p2_host);
/* The same step for inputs that are already on the DEVICE (the decoder's output, src/models/few_shot.py:106-110): no copies;
 * the lane's kernels are ordered behind everything `producer_stream` holds at the time of the call. p1 (B,N,3) / p2 (B,M,3)
 * must stay unchanged until the batch's result was taken. Host and device submits may be mixed. */
int fpsg_chamfer_pipeline_submit_device(fpsg_chamfer_pipeline_t* p, const float* p1_device, const float* p2_device,
                                        fpsg_stream_t producer_stream);
int fpsg_chamfer_pipeline_result(fpsg_chamfer_pipeline_t* p, float* loss_sum_host, float* loss_host,
                                 const float** gp1_device, const int64_t** idx1_device, const int64_t** idx2_device);
/* Multi-GPU: every batch of this pipeline takes part in the scalar exchange above with lag = depth (seq = the batch's
 * number, so all ranks must submit the same number of batches; depth <= 4). tables[r] as in fpsg_peer_exchange_t; call
 * before the first submit. fpsg_chamfer_pipeline_previous_global_sum: after the result of batch n was taken, the
 * all-reduced loss sum of batch n - depth (the most recent one that is complete on every rank without waiting). */
int fpsg_chamfer_pipeline_set_exchange(fpsg_chamfer_pipeline_t* p, unsigned long long* const* tables, int rank, int world);
int fpsg_chamfer_pipeline_previous_global_sum(fpsg_chamfer_pipeline_t* p, float* global_sum_host);
/* The stream the kernels of the batch whose result was taken LAST ran on (to order a consumer of its gp1 behind them
 * without a host round trip; before the first result: lane 0's stream). */
fpsg_stream_t fpsg_chamfer_pipeline_compute_stream(fpsg_chamfer_pipeline_t* p);
int fpsg_chamfer_pipeline_destroy(fpsg_chamfer_pipeline_t* p);

/* Diagnostics: byte offset, inside a workspace of fpsg_chamfer_fwd_workspace_bytes(B, N, M), of a 32-bit counter
 * that every Chamfer forward zeroes and then increments once per point whose arg-min recovery found no candidate
 * bit-equal to its merged minimum (the kernels then fall back to the first candidate). It must read 0 after any
 * call; the parity tests assert that. */
size_t fpsg_chamfer_ws_diag_offset(int B, int N, int M);

/* Non-finite inputs (all Chamfer entry points): a distance that evaluates to NaN is skipped, as by the strict-'<'
 * scan of the reference kernel; a point without any finite distance (NaN / Inf coordinates, or an all-NaN other
 * cloud) gets dist = +INF and idx = 0, what that scan yields when started from +INF. The backward clamps indices
 * to the valid range before gathering, so a caller-supplied idx tensor cannot cause out-of-bounds reads. */

/* ---- Sinkhorn EMD ------------------------------------------------------------------- */
/* The evaluation metric `emd_wrapper(pc1, pc2)` (src/models/utils.py:12-13: neuralnet_pytorch.metrics.emd_loss(pc1, pc2,
 * reduce='sum', sinkhorn=True), called at src/models/few_shot.py:168). neuralnet_pytorch and the geomloss.SamplesLoss() it
 * delegates to are absent from the reference tree: PARITY UNPINNED; the algorithm restated is geomloss' default, the
 * debiased Sinkhorn divergence with cost |x-y|^2/2, uniform weights and symmetrised log-domain updates down a schedule of
 * regularisation strengths (see fpsg_b200/csrc/emd_sinkhorn.cu for the exact recurrences; fpsg_b200/emd.py builds the
 * schedule from blur / scaling / diameter as geomloss does). eps_schedule_host: HOST array of n_eps >= 1 positive values,
 * one averaged update each, after an initialisation at eps[0] and before one final un-averaged step at eps[n_eps-1].
 * out (B,): debias != 0: S_eps(p1,p2) = OT(p1,p2) - OT(p1,p1)/2 - OT(p2,p2)/2; debias == 0: the dual value of OT_eps(p1,p2).
 * f_out (B,N) / g_out (B,M): optional (NULL), the cross potentials of the last step. */
size_t fpsg_emd_sinkhorn_workspace_bytes(int B, int N, int M);
int fpsg_emd_sinkhorn_fwd(const float* p1, const float* p2, int B, int N, int M, const float* eps_schedule_host, int n_eps,
                          int debias, float* out, float* f_out, float* g_out, void* ws, size_t ws_bytes,
                          fpsg_stream_t stream);

/* ---- kNN graph ------------------------------------------------------------------------------------------ */
/* x (B,C,N) channel-major FP32; idx (B,N,k) int64: the k nearest points of every point (self included),
 * nearest first, ranked by the reference's FP32 expression -|xi|^2 + 2 xi.xj - |xj|^2; equal values are
 * ordered by ascending index. 1 <= k <= min(N, FPSG_KNN_MAX_K).
 * Replaces `knn(x, k)`, src/dgcnn/model.py:13-20 (called at :28 for C = 3, 64, 64, 128). Three code paths, one result:
 * C = 3 -- an FP32 threshold-selection kernel; C % 8 == 0, 16 <= C <= 128, k <= 32, 256 <= N <= 4096 (the DGCNN feature
 * layers) -- a tensor-core pre-filter (tcgen05.mma on FP16 copies of the cloud, TMA-fed, TMEM accumulators) that marks a
 * rigorous superset of the k nearest, followed by the exact FP32 expression on the marked candidates only; everything else
 * -- FP32 list kernels. The workspace (fpsg_knn_workspace_bytes; 16-byte aligned) is scratch: no state survives the call.
 * Non-finite features: the call terminates and every index is in [0, N); the ranking of NaN scores is unspecified. */
#define FPSG_KNN_MAX_K 128
size_t fpsg_knn_workspace_bytes(int B, int C, int N, int k);
int fpsg_knn_fwd(const float* x, int B, int C, int N, int k, int64_t* idx, void* ws, size_t ws_bytes,
                 fpsg_stream_t stream);

/* ---- EdgeConv graph feature ------------------------------------------------------- */
/* The tail of `get_graph_feature` (src/dgcnn/model.py:31-41): x (B,C,N), idx (B,N,k) int64 as knn returns it,
 * out (B,2C,N,k): out[b,c,n,j] = x[b,c,idx[b,n,j]] - x[b,c,n] for c < C, out[b,C+c,n,j] = x[b,c,n].
 * Every idx value must lie in [0, N). */
int fpsg_edge_feature_fwd(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out,
                          fpsg_stream_t stream);
/* gout (B,2C,N,k) -> gx (B,C,N); gathers through an inverted index of idx built in the workspace (no
 * floating-point atomics, double accumulation). */
size_t fpsg_edge_feature_bwd_workspace_bytes(int B, int C, int N, int k);
int fpsg_edge_feature_bwd(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, void* ws,
                          size_t ws_bytes, fpsg_stream_t stream);

/* ---- Point-cloud input pipeline ---------------------------------------------------- */
/* The per-item normalisation of `FewShotSubModelNet.__getitem__` (src/datasets/modelnet.py:66-69) for a whole
 * batch on the device: out[b] = (in[b] - mean_i in[b,i]) / max_i |in[b,i] - mean|, in/out (B,N,3) FP32, bit-identical
 * to the reference's numpy float32 arithmetic (rows summed in order). in and out must not overlap. */
int fpsg_normalize_clouds(const float* in, int B, int N, float* out, fpsg_stream_t stream);
/* `ply_reader` (src/datasets/modelnet.py:15-29): reads at most max_points (the reference uses 2048) xyz vertices
 * of an ASCII PLY file into xyz_host (max_points*3 floats, HOST memory); *n_points_host receives the count. */
int fpsg_ply_read_ascii_host(const char* path, int max_points, float* xyz_host, int* n_points_host);

/* ---- Profiling hooks (used by bench.py for the roofline line; not part of the reference-facing surface) ---- */
/* Launches only K1's sweep kernel (the dominant kernel) with the workspace of fpsg_chamfer_fwd; the partial
 * minima stay in the workspace and no output tensor is written. */
int fpsg_chamfer_sweep_only(const float* p1, const float* p2, int B, int N, int M, void* ws, size_t ws_bytes,
                            fpsg_stream_t stream);

/* ---- Roofline probe -------------------------------------------------------------------------------------- */
/* Runs a dependent-free FFMA chain on every SM and returns the achieved FP32 rate in TFLOP/s (2 FLOP per
 * FFMA per lane) through *tflops_host; synchronises `stream`. Used by bench.py as the measured FP32 peak. */
int fpsg_probe_fp32_peak(double* tflops_host, double* sm_mhz_host, fpsg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FPSG_B200_H_ */
