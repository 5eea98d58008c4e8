// Host-buffer streaming step: the shape of the reference's training loop around the loss
// (reference src/trainNetwork.py:141-148: `to_cuda(sample)` -> `model.loss(sample)` -> `ttl_loss.backward()` -> `.item()`),
// as a small native runtime on top of the fused loss step (chamfer_finish.cu).
//
//   submit(p1_host, p2_host) : enqueue the H2D copies of one batch on a COPY stream, then -- on the slot's COMPUTE stream,
//                              behind an event -- sweep + finish (loss, batch sum, gradient w.r.t. p1) and the D2H copy of
//                              the batch's losses into pinned memory. Returns at once.
//   result(...)              : waits for the oldest submitted batch (event), hands out its loss sum, its per-pair losses and
//                              the device address of its gradient.
//
// With `depth` >= 2 slots the copies of batch i+1 run while batch i is being computed: the PCIe time of a batch (1.6 MB,
// about 30 us) hides behind its predecessor's kernels. Every slot is a LANE: it has its own compute stream and its own
// Chamfer workspace, so the kernels of consecutive batches are independent on the device and overlap -- the finish kernel
// of batch i (merge, arg-min recovery, loss, backward: latency-bound, most issue slots idle) runs beside the sweep of batch
// i+1 (issue-bound). Measured at B=32, N=M=2048 with device-resident inputs: 56.1 us per batch on one lane, 45.7 on two,
// 44.6 on three (bench_micro/overlap_probe.py). Inside a lane the usual stream order holds, so a workspace is never
// written before its previous reader has finished. The caller owns all memory -- one device arena and one pinned host
// arena, sized by fpsg_chamfer_pipeline_arena_bytes -- the handle owns 1 + depth streams and 2 x depth events.
// A handle is used by one host thread at a time.
#include <new>

#include "chamfer_plan.cuh"

namespace fpsg {

size_t chamfer_forward_workspace(int B, int N, int M);
int chamfer_fwd_bwd_launch_px(const SegTable& seg, int B, int N, int M, float* dist1, float* dist2, int64_t* idx1, int64_t* idx2, float* loss,
                              float* totals, float w1, float w2, const float* gloss, const float* gtotal, bool multi, void* ws, size_t ws_bytes,
                              cudaStream_t stream, int mode, const PeerExchange* px);

struct PipelineLayout {
    size_t in_off, gp1_off, idx_off, out_off, ws_off, slot_bytes;  // per slot (device); every slot has its own workspace
    size_t ws_bytes, device_bytes;
    size_t out_bytes, pinned_bytes;                        // per-slot result record: totals[2] + loss[B]
};

static PipelineLayout pipeline_layout(int B, int N, int M, int depth) {
    PipelineLayout L;
    size_t off = 0;
    L.in_off = off;
    off += align_up((size_t)B * (N + M) * 3 * sizeof(float), 256);
    L.gp1_off = off;
    off += align_up((size_t)B * N * 3 * sizeof(float), 256);
    L.idx_off = off;
    off += align_up((size_t)B * (N + M) * sizeof(int64_t), 256);
    L.out_off = off;
    L.out_bytes = align_up((2 + (size_t)B + FPSG_PEER_RING) * sizeof(float), 256);  // + the all-reduced sums (multi-GPU)
    off += L.out_bytes;
    L.ws_off = off;
    L.ws_bytes = chamfer_forward_workspace(B, N, M);
    off += align_up(L.ws_bytes, 256);
    L.slot_bytes = off;
    L.device_bytes = L.slot_bytes * depth;
    L.pinned_bytes = L.out_bytes * depth;
    return L;
}

}  // namespace fpsg

using namespace fpsg;

struct fpsg_chamfer_pipeline {
    int B, N, M, depth;
    float w1, w2;
    PipelineLayout L;
    char* dev;
    char* pinned;
    cudaStream_t copy, compute[8];  // one compute stream per slot (lane)
    cudaEvent_t copied[8], done[8];
    long long head, tail;  // batches submitted / results taken
    PeerExchange px;       // multi-GPU: the scalar exchange every step takes part in (world == 0: none), lag = depth
};

extern "C" {

int fpsg_chamfer_pipeline_arena_bytes(int B, int N, int M, int depth, size_t* device_bytes, size_t* pinned_host_bytes) {
    if (B < 1 || N < 1 || M < 1 || depth < 1 || depth > 8) return fail(FPSG_EINVAL, "pipeline: need B, N, M >= 1 and 1 <= depth <= 8");
    const PipelineLayout L = pipeline_layout(B, N, M, depth);
    if (device_bytes) *device_bytes = L.device_bytes;
    if (pinned_host_bytes) *pinned_host_bytes = L.pinned_bytes;
    return FPSG_OK;
}

int fpsg_chamfer_pipeline_create(int B, int N, int M, float w1, float w2, int depth, void* device_arena, size_t device_bytes,
                                 void* pinned_host_arena, size_t pinned_host_bytes, fpsg_chamfer_pipeline_t** out) {
    if (!out) return fail(FPSG_EINVAL, "pipeline: null handle pointer");
    *out = nullptr;
    if (B < 1 || N < 1 || M < 1 || depth < 1 || depth > 8) return fail(FPSG_EINVAL, "pipeline: need B, N, M >= 1 and 1 <= depth <= 8");
    if (B > 65535) return fail(FPSG_EINVAL, "pipeline: B > 65535 not supported");
    const PipelineLayout L = pipeline_layout(B, N, M, depth);
    if (!device_arena || device_bytes < L.device_bytes || (reinterpret_cast<uintptr_t>(device_arena) & 255) != 0)
        return fail(FPSG_EWORKSPACE, "pipeline: a 256-byte aligned device arena of %zu bytes is required", L.device_bytes);
    if (!pinned_host_arena || pinned_host_bytes < L.pinned_bytes)
        return fail(FPSG_EWORKSPACE, "pipeline: a pinned host arena of %zu bytes is required", L.pinned_bytes);
    fpsg_chamfer_pipeline* p = new (std::nothrow) fpsg_chamfer_pipeline();
    if (!p) return fail(FPSG_EINVAL, "pipeline: out of host memory");
    p->B = B, p->N = N, p->M = M, p->depth = depth, p->w1 = w1, p->w2 = w2, p->L = L;
    p->dev = static_cast<char*>(device_arena);
    p->pinned = static_cast<char*>(pinned_host_arena);
    p->head = p->tail = 0;
    p->px = PeerExchange{};
    p->copy = nullptr;
    for (int i = 0; i < 8; ++i) {
        p->copied[i] = p->done[i] = nullptr;
        p->compute[i] = nullptr;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking);
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&p->compute[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->copied[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        fpsg_chamfer_pipeline_destroy(p);
        return fail(FPSG_ECUDA, "pipeline: stream / event creation failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return FPSG_OK;
}

// the compute part of a submit: sweep + finish on the slot's lane, D2H copy of the result record, `done` event
static int pipeline_launch_step(fpsg_chamfer_pipeline* p, int s, const float* d1, const float* d2) {
    char* slot = p->dev + (size_t)s * p->L.slot_bytes;
    cudaStream_t compute = p->compute[s];
    float* gp1 = reinterpret_cast<float*>(slot + p->L.gp1_off);
    int64_t* idx1 = reinterpret_cast<int64_t*>(slot + p->L.idx_off);
    float* outrec = reinterpret_cast<float*>(slot + p->L.out_off);  // totals[2], then loss[B]
    SegTable t = single_segment(d1, d2, gp1, nullptr, p->B);
    PeerExchange px = p->px;
    px.seq = (unsigned)(p->head + 1);
    px.lag = p->depth;  // the batch collected ran in this lane: its own contribution is certainly published
    px.global_out = outrec + 2 + p->B;  // this slot's copy of the ring: [(seq - depth) % ring] = the all-reduced sum of that batch
    const int rc = chamfer_fwd_bwd_launch_px(t, p->B, p->N, p->M, nullptr, nullptr, idx1, idx1 + (size_t)p->B * p->N, outrec + 2, outrec, p->w1, p->w2,
                                             nullptr, nullptr, true, slot + p->L.ws_off, p->L.ws_bytes, compute, 0, px.world > 0 ? &px : nullptr);
    if (rc != FPSG_OK) return rc;
    FPSG_CUDA_OK(cudaMemcpyAsync(p->pinned + (size_t)s * p->L.out_bytes, outrec, (2 + (size_t)p->B + FPSG_PEER_RING) * sizeof(float),
                                 cudaMemcpyDeviceToHost, compute));
    FPSG_CUDA_OK(cudaEventRecord(p->done[s], compute));
    ++p->head;
    return FPSG_OK;
}

int fpsg_chamfer_pipeline_submit(fpsg_chamfer_pipeline_t* p, const float* p1_host, const float* p2_host) {
    if (!p || !p1_host || !p2_host) return fail(FPSG_EINVAL, "pipeline_submit: null argument");
    if (p->head - p->tail >= p->depth) return fail(FPSG_EINVAL, "pipeline_submit: %d batches in flight, take a result first", p->depth);
    const int s = (int)(p->head % p->depth);
    char* slot = p->dev + (size_t)s * p->L.slot_bytes;
    float* d1 = reinterpret_cast<float*>(slot + p->L.in_off);
    float* d2 = d1 + (size_t)p->B * p->N * 3;
    // the slot's inputs are free again once the step that last used them has run (its `done` event: the caller took that
    // result before this slot could come round again, so the event is long complete -- the wait costs nothing)
    if (p->head >= p->depth) FPSG_CUDA_OK(cudaStreamWaitEvent(p->copy, p->done[s], 0));
    FPSG_CUDA_OK(cudaMemcpyAsync(d1, p1_host, (size_t)p->B * p->N * 3 * sizeof(float), cudaMemcpyHostToDevice, p->copy));
    FPSG_CUDA_OK(cudaMemcpyAsync(d2, p2_host, (size_t)p->B * p->M * 3 * sizeof(float), cudaMemcpyHostToDevice, p->copy));
    FPSG_CUDA_OK(cudaEventRecord(p->copied[s], p->copy));
    FPSG_CUDA_OK(cudaStreamWaitEvent(p->compute[s], p->copied[s], 0));
    return pipeline_launch_step(p, s, d1, d2);
}

int fpsg_chamfer_pipeline_submit_device(fpsg_chamfer_pipeline_t* p, const float* p1_device, const float* p2_device, fpsg_stream_t producer_stream) {
    if (!p || !p1_device || !p2_device) return fail(FPSG_EINVAL, "pipeline_submit_device: null argument");
    if (p->head - p->tail >= p->depth) return fail(FPSG_EINVAL, "pipeline_submit_device: %d batches in flight, take a result first", p->depth);
    const int s = (int)(p->head % p->depth);
    // the lane starts behind whatever the producer's stream holds now (the kernels that wrote p1 / p2)
    FPSG_CUDA_OK(cudaEventRecord(p->copied[s], static_cast<cudaStream_t>(producer_stream)));
    FPSG_CUDA_OK(cudaStreamWaitEvent(p->compute[s], p->copied[s], 0));
    return pipeline_launch_step(p, s, p1_device, p2_device);
}

int fpsg_chamfer_pipeline_result(fpsg_chamfer_pipeline_t* p, float* loss_sum_host, float* loss_host, const float** gp1_device,
                                 const int64_t** idx1_device, const int64_t** idx2_device) {
    if (!p) return fail(FPSG_EINVAL, "pipeline_result: null handle");
    if (p->tail >= p->head) return fail(FPSG_EINVAL, "pipeline_result: nothing in flight");
    const int s = (int)(p->tail % p->depth);
    FPSG_CUDA_OK(cudaEventSynchronize(p->done[s]));
    const float* rec = reinterpret_cast<const float*>(p->pinned + (size_t)s * p->L.out_bytes);
    if (loss_sum_host) *loss_sum_host = rec[0];
    if (loss_host)
        for (int b = 0; b < p->B; ++b) loss_host[b] = rec[2 + b];
    char* slot = p->dev + (size_t)s * p->L.slot_bytes;
    if (gp1_device) *gp1_device = reinterpret_cast<const float*>(slot + p->L.gp1_off);
    if (idx1_device) *idx1_device = reinterpret_cast<const int64_t*>(slot + p->L.idx_off);
    if (idx2_device) *idx2_device = reinterpret_cast<const int64_t*>(slot + p->L.idx_off) + (size_t)p->B * p->N;
    ++p->tail;
    return FPSG_OK;
}

int fpsg_chamfer_pipeline_set_exchange(fpsg_chamfer_pipeline_t* p, unsigned long long* const* tables, int rank, int world) {
    if (!p) return fail(FPSG_EINVAL, "pipeline_set_exchange: null handle");
    if (p->head != 0) return fail(FPSG_EINVAL, "pipeline_set_exchange: must be called before the first submit");
    if (!tables || world < 1 || world > FPSG_MAX_PEERS || rank < 0 || rank >= world) return fail(FPSG_EINVAL, "pipeline_set_exchange: bad rank/world");
    if (p->depth > FPSG_PEER_RING / 4) return fail(FPSG_EINVAL, "pipeline_set_exchange: depth must be <= %d with an exchange", FPSG_PEER_RING / 4);
    p->px = PeerExchange{};
    for (int r = 0; r < world; ++r) {
        if (!tables[r]) return fail(FPSG_EINVAL, "pipeline_set_exchange: null table for rank %d", r);
        p->px.table[r] = tables[r];
    }
    p->px.rank = rank, p->px.world = world;
    return FPSG_OK;
}

int fpsg_chamfer_pipeline_previous_global_sum(fpsg_chamfer_pipeline_t* p, float* global_sum_host) {
    if (!p || !global_sum_host) return fail(FPSG_EINVAL, "pipeline_previous_global_sum: null argument");
    if (p->px.world == 0 || p->tail <= p->depth)
        return fail(FPSG_EINVAL, "pipeline_previous_global_sum: needs an exchange and more than `depth` results taken");
    // the batch whose result was taken last (number tail, 1-based) collected the sum of batch tail - depth
    const int s = (int)((p->tail - 1) % p->depth);
    const float* rec = reinterpret_cast<const float*>(p->pinned + (size_t)s * p->L.out_bytes);
    *global_sum_host = rec[2 + p->B + (p->tail - p->depth) % FPSG_PEER_RING];
    return FPSG_OK;
}

fpsg_stream_t fpsg_chamfer_pipeline_compute_stream(fpsg_chamfer_pipeline_t* p) {
    if (!p) return nullptr;
    return static_cast<fpsg_stream_t>(p->compute[p->tail > 0 ? (p->tail - 1) % p->depth : 0]);
}

int fpsg_chamfer_pipeline_destroy(fpsg_chamfer_pipeline_t* p) {
    if (!p) return FPSG_OK;
    for (int i = 0; i < 8; ++i)
        if (p->compute[i]) cudaStreamSynchronize(p->compute[i]);
    if (p->copy) cudaStreamSynchronize(p->copy);
    for (int i = 0; i < 8; ++i) {
        if (p->copied[i]) cudaEventDestroy(p->copied[i]);
        if (p->done[i]) cudaEventDestroy(p->done[i]);
    }
    if (p->copy) cudaStreamDestroy(p->copy);
    for (int i = 0; i < 8; ++i)
        if (p->compute[i]) cudaStreamDestroy(p->compute[i]);
    delete p;
    return FPSG_OK;
}

}  // extern "C"
