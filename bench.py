#!/usr/bin/env python
"""bench.py -- headline benchmark of fpsg_b200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): Chamfer fwd+bwd pairs/s at B=32, N=M=2048 per GPU (weak scaling over 1/2/4/8 B200),
plus the dominant kernel's fraction of the FP32 FMA roofline.

A *step* = one pass of the hot path over one batch of 32 synthetic cloud pairs: Chamfer loss forward and backward
to p1 (the ground truth has no gradient, as in the reference's training loss, src/models/few_shot.py:109-117) as
two kernels -- the sweep and the fused finish kernel (merge, arg-min, loss, batch sum, backward) -- and, for N > 1,
the scalar loss all-reduce done inside the finish kernel's epilogue through NVLink peer memory, every step.

  value : whole-job pairs/s, inputs resident in HBM, the C ABI called directly, CUDA events, max over ranks.
  e2e   : the same metric through the host-buffer call of the C ABI (fpsg_chamfer_pipeline_*): every step copies
          its inputs from pinned HOST memory and reads its losses back; copies inside the timed region.
          e2e_autograd: the reference-facing Python call (kaolin shim, torch.autograd), serial, for comparison.
  roofline     : the sweep kernel timed alone (fpsg_chamfer_sweep_only) against the FP32 FMA peak measured in
                 the same run by an FFMA micro-kernel (MEASURED_PEAKS.json has no FP32 row).
  cpu_baseline : the reference's PyTorch CPU formulation (oracle/torch_formulation.py) on a bounded sample.

`--impl reference` times that CPU formulation on the host cores as its own arm.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

# The step is issued over several streams (lanes) and the host-buffer pipeline owns 1 + depth more: with the default of 8
# hardware work queues, streams get aliased onto shared queues and falsely serialise each other (measured: e2e 410-560 k
# pairs/s instead of 715 k). Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_PTS, M_PTS = 32, 2048, 2048
WORKLOAD = "c1: Chamfer loss fwd+bwd(p1), B=32 per GPU, N=M=2048, synthetic xyz float32"
METRIC = "chamfer_fwd_bwd_pairs_per_s"
RING_SETS = 96  # distinct input batches cycled through: 96 x 1.57 MB = 151 MB > 126 MB L2
N_LANES = 4     # streams (each with its own workspace and outputs) the independent steps are issued over, round-robin
# dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch, from the ncu --set full capture named below (the
# kernel streams its 1.573 MB of input once; a number taken under the profiler is cited, not re-measured in the timed run)
SWEEP_DRAM_TRAFFIC_BYTES = 1.599e6
SWEEP_DRAM_TRAFFIC_SOURCE = "profiles/r02e_step_kernels_raw.csv / r02e_step_kernels_digest.txt (dram__bytes_read.sum + dram__bytes_write.sum = 1.5987 MB + 256 B, chamfer_sweep_kernel, final build)"


def make_inputs(batch, n, m, seed, device=None, pin=False):
    """p1 = tanh(0.5*randn) (decoder-like), p2 = unit-ball normalised randn; seeded on CPU."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    p1 = torch.tanh(0.5 * torch.randn(batch, n, 3, generator=g))
    p2 = torch.randn(batch, m, 3, generator=g)
    p2 = p2 - p2.mean(1, keepdim=True)
    p2 = p2 / p2.norm(dim=2).amax(1)[:, None, None]
    if pin:
        return p1.pin_memory(), p2.pin_memory()
    if device is not None:
        return p1.to(device), p2.to(device)
    return p1, p2


class ClockSampler:
    """Polls nvidia-smi (the recipe's clocks line) while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []      # (host time of arrival, fields)
        self.windows = []   # [t_begin, t_end] of the timed regions: only samples taken inside them are reported
        self.proc = None
        self.thread = None

    def mark(self, t_begin, t_end):
        self.windows.append((t_begin, t_end))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "10", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append((time.perf_counter(), parts))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [p for t, p in self.rows if any(a <= t <= b + 0.012 for a, b in self.windows)]
        total = len(self.rows)
        for p in (inside if inside else [p for _, p in self.rows]):
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_regions": len(inside), "samples_total": total}


def cpu_formulation_pairs_per_s(sample_pairs: int, iters: int, threads: int | None = None):
    """Times the reference's PyTorch CPU formulation (difference-form squared L2 -> min both ways -> mean+mean ->
    sum -> backward to p1) on `sample_pairs` pairs of the c1 workload. Returns (pairs/s, threads, seconds/iter)."""
    import torch
    from oracle import torch_formulation as TF
    if threads:
        torch.set_num_threads(threads)
    p1, p2 = make_inputs(sample_pairs, N_PTS, M_PTS, seed=1234 + 1)
    best = None
    for it in range(iters + 1):  # first iteration is the warm-up
        a = p1.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = TF.chamfer_distance_torch(a, p2, chunk=4)
        loss.sum().backward()
        dt = time.perf_counter() - t0
        if it > 0:
            best = dt if best is None else min(best, dt)
    return sample_pairs / best, torch.get_num_threads(), best


def run_reference_arm(args):
    """`--impl reference`: the CPU formulation on the box's host cores. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    t_all0 = time.perf_counter()
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import torch_formulation as TF

    def run(p1, p2):
        a = p1.clone().requires_grad_(True)
        TF.chamfer_distance_torch(a, p2, chunk=4).sum().backward()

    # size the per-step sample (1..8 of the 32 pairs) so that warmup + steps finish in about a minute
    probe1, probe2 = make_inputs(4, N_PTS, M_PTS, seed=1234 + 1)
    run(probe1, probe2)
    t0 = time.perf_counter()
    run(probe1, probe2)
    per_pair = (time.perf_counter() - t0) / 4
    sample = int(max(1, min(8, 60.0 / ((args.steps + args.warmup) * per_pair))))
    p1, p2 = make_inputs(sample, N_PTS, M_PTS, seed=1234 + 1)

    def step():
        run(p1, p2)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}; reference arm: each step is a bounded sample of {sample} of the 32 pairs (value is a per-pair rate)",
                   "sample_pairs_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} pairs x {args.steps} steps of the c1 workload, pure-PyTorch difference-form Chamfer fwd+bwd "
                                   f"(oracle/torch_formulation.py; Kaolin 0.9.0 itself is absent), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def _timed_us(torch, stream, dev, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(iters):
        fn()
    t1.record(stream)
    torch.cuda.synchronize(dev)
    return t0.elapsed_time(t1) / iters * 1e3


def secondary_configs(torch, L, _lib, dev, stream, fp32_peak_tflops, ptrs, N, M):
    """The other configurations of BASELINE.json (c2, c3, c4), device-resident inputs, the C ABI called directly. Not the
    metric: reported for context, rank 0 only. Every Chamfer entry: us per call, pairs/s, fraction of the FP32 peak in the
    16*N*M-FLOP-per-pair convention."""
    sp = stream.cuda_stream
    sec = {}

    def chamfer_entry(us, pairs, n, m):
        tf = 16.0 * pairs * n * m / us / 1e6
        return {"us_per_call": us, "pairs_per_s": pairs / us * 1e6, "tflops": tf, "frac_of_fp32_peak": tf / fp32_peak_tflops}

    # ---- c3: DGCNN kNN graph, k=20, B=32, N=1024: the xyz layer and the 64- / 128-channel layers (src/dgcnn/model.py:62-74)
    gk = torch.Generator(device="cpu").manual_seed(1234 + 3)
    # (N = 1024 is BASELINE.json's c3; the reference's own clouds have 2048 points, src/datasets/modelnet.py: reported beside it)
    for NK_, C_ in ((1024, 3), (1024, 64), (1024, 128), (2048, 3), (2048, 64), (2048, 128)):
        xk = torch.randn(32, C_, NK_, generator=gk).to(dev)
        ik = torch.empty(32, NK_, 20, dtype=torch.int64, device=dev)
        nbk = L.fpsg_knn_workspace_bytes(32, C_, NK_, 20)
        wsk = torch.empty(max(nbk, 1), dtype=torch.uint8, device=dev)
        us = _timed_us(torch, stream, dev, lambda: _lib.check(L.fpsg_knn_fwd(xk.data_ptr(), 32, C_, NK_, 20, ik.data_ptr(), wsk.data_ptr(), nbk, sp)), 50, warm=5)
        tf = (2 * C_ + 2) * 32 * NK_ * NK_ / us / 1e6
        sec[("c3_" if NK_ == 1024 else "") + f"knn_k20_B32_N{NK_}_C{C_}"] = {"us_per_call": us, "clouds_per_s": 32 / us * 1e6, "tflops": tf, "frac_of_fp32_peak": tf / fp32_peak_tflops,
                                             "path": "FP32 threshold-selection kernel (knn.cu)" if C_ == 3 else
                                                     "tcgen05 FP16 pre-filter + exact FP32 re-scoring of the marked candidates (knn_tc.cu); tflops = the "
                                                     "algorithmic (2C+2) N^2 FLOP of the reference's matmul formulation over the time, not an FP32 rate"}
        if C_ in (3, 64) and NK_ == 1024:
            # the EdgeConv graph feature that consumes this kNN graph (src/dgcnn/model.py:31-41): forward writes
            # (B, 2C, N, k) once, backward reads it once -- both HBM-bound, reported against the measured copy bandwidth of MEASURED_PEAKS.json
            ko = torch.empty(32, 2 * C_, NK_, 20, device=dev)
            gxk = torch.empty(32, C_, NK_, device=dev)
            nbe = L.fpsg_edge_feature_bwd_workspace_bytes(32, C_, NK_, 20)
            wse = torch.empty(max(nbe, 1), dtype=torch.uint8, device=dev)
            us_f = _timed_us(torch, stream, dev, lambda: _lib.check(L.fpsg_edge_feature_fwd(xk.data_ptr(), ik.data_ptr(), 32, C_, NK_, 20, ko.data_ptr(), sp)), 10)
            us_b = _timed_us(torch, stream, dev, lambda: _lib.check(L.fpsg_edge_feature_bwd(ko.data_ptr(), ik.data_ptr(), 32, C_, NK_, 20, gxk.data_ptr(), wse.data_ptr(), nbe, sp)), 10)
            nbytes = ko.numel() * 4
            hbm = _hbm_peak()
            sec[f"edge_feature_k20_B32_N{NK_}_C{C_}"] = {"fwd_us": us_f, "fwd_gb_per_s_written": nbytes / us_f / 1e3, "bwd_us": us_b, "bwd_gb_per_s_read": nbytes / us_b / 1e3,
                                                         "bytes": nbytes, "hbm_copy_peak_gb_per_s": hbm, "fwd_frac_of_hbm_peak": nbytes / us_f / 1e3 / hbm,
                                                         "bwd_frac_of_hbm_peak": nbytes / us_b / 1e3 / hbm}
            del ko, gxk, wse
        del xk, ik, wsk

    # ---- c2: one training step of 32 episodes = 32 x (5 query + 32 support) = 1184 pairs, N=M=2048
    B2 = 1184
    p1c2, p2c2 = make_inputs(B2, N, M, seed=1234 + 2, device=dev)
    gp = torch.empty(B2, N, 3, device=dev)
    ix = torch.empty(B2 * (N + M), dtype=torch.int64, device=dev)
    lo = torch.empty(B2 + 4, device=dev)
    nb2 = L.fpsg_chamfer_fwd_workspace_bytes(B2, N, M)
    ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
    a2, b2, ip = p1c2.data_ptr(), p2c2.data_ptr(), ix.data_ptr()

    def c2_one_call():
        _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a2, b2, B2, N, M, 1.0, 1.0, None, lo.data_ptr(), None, None, ip, ip + 8 * B2 * N, gp.data_ptr(), None,
                                               ws2.data_ptr(), nb2, sp))
    sec["c2_B1184_one_call"] = chamfer_entry(_timed_us(torch, stream, dev, c2_one_call, 10), B2, N, M)

    def c2_as_issued():  # the reference's own call pattern: per episode a 5-pair call and a 32-pair call (few_shot.py:110,117)
        for e in range(32):
            o5, o32 = e * 37, e * 37 + 5
            for off, nb_ in ((o5, 5), (o32, 32)):
                _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a2 + off * N * 12, b2 + off * M * 12, nb_, N, M, 1.0, 1.0, None, lo.data_ptr() + 4 * off, None, None,
                                                       ip + 8 * off * N, ip + 8 * (B2 * N + off * M), gp.data_ptr() + off * N * 12, None,
                                                       ws2.data_ptr(), nb2, sp))
    sec["c2_B1184_as_64_calls"] = chamfer_entry(_timed_us(torch, stream, dev, c2_as_issued, 5, 2), B2, N, M)

    # the same 64 calls, issued round-robin over N_LANES lanes (a stream + workspace each; outputs are disjoint slices anyway):
    # the calls of a training step are independent, so the 5-pair calls and the tails of the 32-pair calls overlap their neighbours
    nb37 = L.fpsg_chamfer_fwd_workspace_bytes(32, N, M)
    lane_s = [torch.cuda.Stream(device=dev) for _ in range(N_LANES)]
    lane_ws = [torch.empty(nb37, dtype=torch.uint8, device=dev) for _ in range(N_LANES)]

    def c2_as_issued_lanes():
        ev = torch.cuda.Event()
        ev.record(stream)
        for st_ in lane_s:
            st_.wait_event(ev)
        c = 0
        for e in range(32):
            o5, o32 = e * 37, e * 37 + 5
            for off, nb_ in ((o5, 5), (o32, 32)):
                ln = c % N_LANES
                c += 1
                _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a2 + off * N * 12, b2 + off * M * 12, nb_, N, M, 1.0, 1.0, None, lo.data_ptr() + 4 * off, None, None,
                                                       ip + 8 * off * N, ip + 8 * (B2 * N + off * M), gp.data_ptr() + off * N * 12, None,
                                                       lane_ws[ln].data_ptr(), nb37, lane_s[ln].cuda_stream))
        for st_ in lane_s:
            e2_ = torch.cuda.Event()
            e2_.record(st_)
            stream.wait_event(e2_)
    sec[f"c2_B1184_as_64_calls_{N_LANES}_lanes"] = chamfer_entry(_timed_us(torch, stream, dev, c2_as_issued_lanes, 5, 2), B2, N, M)
    del lane_ws

    sg2 = (_lib.ChamferSegment * 2)()
    sg2[0].p1, sg2[0].p2, sg2[0].gp1, sg2[0].gp2, sg2[0].batch, sg2[0].weight = a2, b2, gp.data_ptr(), None, 5, 1.0
    sg2[1].p1, sg2[1].p2, sg2[1].gp1, sg2[1].gp2 = a2 + 5 * N * 12, b2 + 5 * M * 12, gp.data_ptr() + 5 * N * 12, None
    sg2[1].batch, sg2[1].weight = 32, 1.0

    def episode_two_calls():
        for off, nb_ in ((0, 5), (5, 32)):
            _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a2 + off * N * 12, b2 + off * M * 12, nb_, N, M, 1.0, 1.0, None, lo.data_ptr() + 4 * off, None, None,
                                                   ip + 8 * off * N, ip + 8 * (B2 * N + off * M), gp.data_ptr() + off * N * 12, None, ws2.data_ptr(), nb2, sp))

    def episode_one_call():
        _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd(sg2, 2, N, M, 1.0, 1.0, None, lo.data_ptr() + 16, lo.data_ptr(), ip, ip + 8 * 37 * N,
                                                     ws2.data_ptr(), nb2, sp))
    sec["c2_episode_5q_32s"] = {"two_calls_us": _timed_us(torch, stream, dev, episode_two_calls, 50),
                                "multi_segment_us": _timed_us(torch, stream, dev, episode_one_call, 50)}
    del p1c2, p2c2, gp, ix, ws2

    # ---- the evaluation metric beside the Chamfer distance (few_shot.py:168): Sinkhorn EMD of 32 pairs of 2048 points,
    # geomloss' default schedule for diameter 2 (10 soft-min steps of four N x M sweeps each)
    import fpsg_b200
    from fpsg_b200.emd import epsilon_schedule
    e1, e2 = make_inputs(32, N, M, seed=1234 + 6, device=dev)
    sched = epsilon_schedule(2.0)
    us = _timed_us(torch, stream, dev, lambda: fpsg_b200.emd_sinkhorn(e1, e2, eps_schedule=sched), 5, 2)
    nexp = (len(sched) + 2) * 2.0 * 32 * (N * N + M * M + 2.0 * N * M) / 2
    sec["emd_sinkhorn_B32_N2048"] = {"us_per_call": us, "pairs_per_s": 32 / us * 1e6, "softmin_steps": len(sched) + 2,
                                     "exp_per_s": nexp / us * 1e6, "note": "evaluation-only metric, parity unpinned (geomloss absent)"}
    del e1, e2

    # ---- c4: large clouds, B=64, N=M in {4096, 8192, 16384}
    for n4 in (4096, 8192, 16384):
        B4 = 64
        p1, p2 = make_inputs(B4, n4, n4, seed=1234 + 4, device=dev)
        gp = torch.empty(B4, n4, 3, device=dev)
        ix = torch.empty(B4 * 2 * n4, dtype=torch.int64, device=dev)
        lo4 = torch.empty(B4, device=dev)
        nb4 = L.fpsg_chamfer_fwd_workspace_bytes(B4, n4, n4)
        ws4 = torch.empty(nb4, dtype=torch.uint8, device=dev)

        def c4_call():
            _lib.check(L.fpsg_chamfer_loss_fwd_bwd(p1.data_ptr(), p2.data_ptr(), B4, n4, n4, 1.0, 1.0, None, lo4.data_ptr(), None, None, ix.data_ptr(),
                                                   ix.data_ptr() + 8 * B4 * n4, gp.data_ptr(), None, ws4.data_ptr(), nb4, sp))
        sec[f"c4_B64_N{n4}"] = chamfer_entry(_timed_us(torch, stream, dev, c4_call, 10 if n4 < 16384 else 4, 2), B4, n4, n4)
        del p1, p2, gp, ix, ws4
    return sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the bounded CPU-baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c2 / c3 / c4 / c5 context measurements")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fpsg_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    exchange = None
    collective = "none (single GPU)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the communicator is created (first collective); stdout must
        # carry exactly one JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
            # the scalar all-reduce of the path: done inside the loss kernel's epilogue through NVLink peer memory
            # (fpsg_b200/sharded.py); torch.distributed only sets the tables up and reduces the timings at the end
            from fpsg_b200.sharded import PeerScalarExchange
            try:
                exchange = PeerScalarExchange(dev)
                collective = (f"one all-reduce per step, done inside the loss kernel's epilogue through NVLink peer memory, software-pipelined "
                              f"(lag {N_LANES} = the number of lanes: step s publishes its sum and collects the sum of step s-{N_LANES}, which ran in the "
                              f"same lane; a final flush collects the last {N_LANES} -- no rank ever waits for another); value_collective_in_step = the "
                              f"same with lag 0 (every step waits for all ranks' values of that step); value_without_collective = no exchange at all")
            except Exception as exc:  # symmetric memory unavailable: NCCL, amortised, as in round 1
                print(f"[bench] peer exchange unavailable ({exc!r}); falling back to NCCL all-reduce, 8 steps per collective", file=sys.stderr)
                ok = torch.zeros(1, device=dev)
                dist.all_reduce(ok)  # keep the ranks aligned
                collective = "NCCL all-reduce, 8 steps per collective (peer exchange unavailable)"
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi takes a few hundred ms to deliver its first sample: started ahead of the input set-up

    from fpsg_b200 import _lib
    import fpsg_b200
    import fpsg_b200.shims
    L = _lib.lib()
    B, N, M = B_PER_GPU, N_PTS, M_PTS

    # ---- FP32 peak probe (roofline denominator), same process, same GPU ----
    tf, mhz = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.fpsg_probe_fp32_peak(ctypes.byref(tf), ctypes.byref(mhz), None))
    fp32_peak_tflops, probe_mhz = tf.value, mhz.value

    # ---- inputs: a ring of distinct batches larger than L2, resident in HBM ----
    ring = [make_inputs(B, N, M, seed=1234 + 1 + 1000 * rank + s, device=dev) for s in range(RING_SETS)]
    gloss = torch.ones(B, device=dev)
    nbytes = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in ring]

    # A lane = a stream with its own workspace and outputs. Consecutive steps work on independent batches, so they are
    # issued round-robin over LANES lanes: the latency-bound finish kernel of step i (merge, arg-min recovery, loss,
    # backward) overlaps the issue-bound sweep of step i+1. Inside a lane the usual stream order holds.
    class Lane:
        def __init__(self, own_stream):
            self.idx = torch.empty(B * (N + M), dtype=torch.int64, device=dev)
            self.out = torch.zeros(2 + B, device=dev)       # totals[2] (weighted total, segment sum), then the per-pair losses
            self.gp1 = torch.empty(B, N, 3, device=dev)
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.stream = torch.cuda.Stream(device=dev) if own_stream else stream
            self.sp = self.stream.cuda_stream
            self.ip1, self.ip2 = self.idx.data_ptr(), self.idx.data_ptr() + 8 * B * N
            self.segs = []
            for a_, b_ in ptrs:
                sg = (_lib.ChamferSegment * 1)()
                sg[0].p1, sg[0].p2, sg[0].gp1, sg[0].gp2, sg[0].batch, sg[0].weight = a_, b_, self.gp1.data_ptr(), None, B, 1.0
                self.segs.append(sg)

    nccl_fallback = world > 1 and exchange is None
    LANES = 1 if nccl_fallback else N_LANES
    lanes = [Lane(False)] + [Lane(True) for _ in range(LANES - 1)]
    idx, gp1, ws = lanes[0].idx, lanes[0].gp1, lanes[0].ws
    ip1, ip2 = lanes[0].ip1, lanes[0].ip2
    KLOG = 8
    nccl_ring = [torch.zeros(KLOG, 2, device=dev) for _ in range(2)]
    pending = [None, None]

    # One step = sweep + finish: per-pair losses, their sum over the local batch, the gradient w.r.t. p1 (the upstream
    # gradient of `.sum()` / `ttl_loss.backward()` is the constant 1), and for N > 1 the scalar exchange -- two kernels.
    def step(i, mode="lag", nl=1):
        ln = lanes[i % nl]
        out = ln.out
        if world == 1 or mode == "none":
            _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd(ln.segs[i % RING_SETS], 1, N, M, 1.0, 1.0, None, out.data_ptr() + 8, out.data_ptr(), ln.ip1, ln.ip2,
                                                         ln.ws.data_ptr(), nbytes, ln.sp))
        elif exchange is not None:
            _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd_dist(ln.segs[i % RING_SETS], 1, N, M, 1.0, 1.0, None, out.data_ptr() + 8, out.data_ptr(), ln.ip1,
                                                              ln.ip2, ln.ws.data_ptr(), nbytes, exchange.next(nl if mode == "lag" else 0), ln.sp))
        else:
            k, j = (i // KLOG) & 1, i % KLOG
            if j == 0 and pending[k] is not None:
                if not pending[k].is_completed():
                    pending[k].wait()
                pending[k] = None
            _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd(ln.segs[i % RING_SETS], 1, N, M, 1.0, 1.0, None, out.data_ptr() + 8,
                                                         nccl_ring[k].data_ptr() + 8 * j, ln.ip1, ln.ip2, ln.ws.data_ptr(), nbytes, ln.sp))
            if j == KLOG - 1:
                pending[k] = dist.all_reduce(nccl_ring[k], op=dist.ReduceOp.SUM, async_op=True)

    kernels_per_step = 2  # sweep, finish (merge + arg-min + loss + batch sum + backward + scalar exchange)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_steps(nsteps, mode="lag", nl=1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        tw0 = time.perf_counter()
        e0.record(stream)
        for ln in lanes[1:nl]:
            ln.stream.wait_event(e0)
        for i in range(nsteps):
            step(args.warmup + i, mode, nl)
        for ln in lanes[1:nl]:   # the timed region ends when every lane has finished
            ev = torch.cuda.Event()
            ev.record(ln.stream)
            stream.wait_event(ev)
        if exchange is not None and mode == "lag":
            for sq in range(max(1, exchange.seq - nl + 1), exchange.seq + 1):
                exchange.collect(sq, sp)  # the flush: the sums of the last `nl` steps
        e1.record(stream)
        for k in (0, 1):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
        barrier()
        sampler.mark(tw0, time.perf_counter())
        return e0.elapsed_time(e1)

    for i in range(args.warmup):
        step(i, "lag", LANES)
    barrier()
    ms = timed_steps(args.steps, "lag", LANES)
    # the same steps strictly one after the other on ONE stream: the latency of a step
    for i in range(10):
        step(i, "lag", 1)
    nq = max(50, args.steps // 4)
    ms_serial = timed_steps(nq, "lag", 1) / nq * args.steps
    ms_nocoll = ms_lag0 = None
    if world > 1:  # the same loop without the collective, and with the collective completed inside every step
        for i in range(10):
            step(i, "none", LANES)
        ms_nocoll = timed_steps(nq, "none", LANES) / nq * args.steps
        for i in range(10):
            step(i, "lag0", LANES)
        ms_lag0 = timed_steps(nq, "lag0", LANES) / nq * args.steps

    # ---- dominant kernel alone (roofline), same ring, same stream ----
    def sweep(i):
        a, b = ptrs[i % RING_SETS]
        _lib.check(L.fpsg_chamfer_sweep_only(a, b, B, N, M, ws.data_ptr(), nbytes, sp))
    for i in range(5):
        sweep(i)
    torch.cuda.synchronize(dev)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_sweeps = max(50, min(args.steps, 400))
    tw0 = time.perf_counter()
    s0.record(stream)
    for i in range(n_sweeps):
        sweep(5 + i)
    s1.record(stream)
    torch.cuda.synchronize(dev)
    sampler.mark(tw0, time.perf_counter())
    sweep_us = s0.elapsed_time(s1) / n_sweeps * 1e3

    # the same kernel alone, launches alternating over the lanes (each with its own workspace): ramp and tail of consecutive
    # launches overlap, which is how the kernel runs inside the timed step above (reported beside the one-stream figure)
    def sweep_lane(i):
        a, b = ptrs[i % RING_SETS]
        ln = lanes[i % LANES]
        _lib.check(L.fpsg_chamfer_sweep_only(a, b, B, N, M, ln.ws.data_ptr(), nbytes, ln.sp))
    for i in range(8):
        sweep_lane(i)
    torch.cuda.synchronize(dev)
    so0, so1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    so0.record(stream)
    for ln in lanes[1:LANES]:
        ln.stream.wait_event(so0)
    for i in range(n_sweeps):
        sweep_lane(8 + i)
    for ln in lanes[1:LANES]:
        ev = torch.cuda.Event()
        ev.record(ln.stream)
        stream.wait_event(ev)
    so1.record(stream)
    torch.cuda.synchronize(dev)
    sweep_lanes_us = so0.elapsed_time(so1) / n_sweeps * 1e3

    # ---- the separate backward kernel (autograd path), HBM view ----
    def bwd(i):
        a, b = ptrs[i % RING_SETS]
        _lib.check(L.fpsg_chamfer_loss_bwd(a, b, ip1, ip2, gloss.data_ptr(), 1.0, 1.0, B, N, M, gp1.data_ptr(), None, None, 0, sp))
    bwd_us = _timed_us(torch, stream, dev, lambda: bwd(0), n_sweeps, 5)
    clocks = sampler.stop() if rank == 0 else None

    # ---- c5 (BASELINE.json configs[4]): STRONG scaling -- 256 pairs of 2048 points in total, sharded over the ranks,
    # Chamfer loss step with the per-step scalar exchange, and the kNN graph (k=20, xyz) of the same 256 clouds ----
    c5 = None
    if not args.no_secondary:
        B5 = 256 // world
        p1s, p2s = make_inputs(B5, N, M, seed=1234 + 5 + 1000 * rank, device=dev)
        gp5 = torch.empty(B5, N, 3, device=dev)
        ix5 = torch.empty(B5 * (N + M), dtype=torch.int64, device=dev)
        out5 = torch.zeros(2 + B5, device=dev)
        nb5 = L.fpsg_chamfer_fwd_workspace_bytes(B5, N, M)
        ws5 = torch.empty(nb5, dtype=torch.uint8, device=dev)
        sg5 = (_lib.ChamferSegment * 1)()
        sg5[0].p1, sg5[0].p2, sg5[0].gp1, sg5[0].gp2, sg5[0].batch, sg5[0].weight = p1s.data_ptr(), p2s.data_ptr(), gp5.data_ptr(), None, B5, 1.0

        def c5_step():
            if exchange is not None:
                _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd_dist(sg5, 1, N, M, 1.0, 1.0, None, out5.data_ptr() + 8, out5.data_ptr(), ix5.data_ptr(),
                                                                  ix5.data_ptr() + 8 * B5 * N, ws5.data_ptr(), nb5, exchange.next(0), sp))
            else:
                _lib.check(L.fpsg_chamfer_multi_loss_fwd_bwd(sg5, 1, N, M, 1.0, 1.0, None, out5.data_ptr() + 8, out5.data_ptr(), ix5.data_ptr(),
                                                             ix5.data_ptr() + 8 * B5 * N, ws5.data_ptr(), nb5, sp))
        xk5 = p2s.transpose(1, 2).contiguous()  # (B5, 3, 2048): the ground-truth clouds as the DGCNN encoder sees them
        ik5 = torch.empty(B5, N, 20, dtype=torch.int64, device=dev)
        nbk5 = L.fpsg_knn_workspace_bytes(B5, 3, N, 20)
        wsk5 = torch.empty(max(nbk5, 1), dtype=torch.uint8, device=dev)

        def c5_knn():
            _lib.check(L.fpsg_knn_fwd(xk5.data_ptr(), B5, 3, N, 20, ik5.data_ptr(), wsk5.data_ptr(), nbk5, sp))
        barrier()
        t_ch = _timed_us(torch, stream, dev, c5_step, 30, 5)
        barrier()
        t_kn = _timed_us(torch, stream, dev, c5_knn, 10, 3)
        barrier()
        tt = torch.tensor([t_ch, t_kn], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ch, t_kn = float(tt[0]), float(tt[1])
        c5 = {"scaling": "strong", "total_pairs": 256, "pairs_per_gpu": B5,
              "chamfer_fwd_bwd": {"us_per_step_max_over_ranks": t_ch, "pairs_per_s": 256 / t_ch * 1e6,
                                  "frac_of_fp32_peak_per_gpu": 16.0 * B5 * N * M / t_ch / 1e6 / fp32_peak_tflops},
              "knn_k20_xyz_N2048": {"us_per_call_max_over_ranks": t_kn, "clouds_per_s": 256 / t_kn * 1e6},
              "chamfer_plus_knn_pairs_per_s": 256 / (t_ch + t_kn) * 1e6}
        del p1s, p2s, gp5, ix5, ws5, xk5, ik5, wsk5

    secondary = None
    if rank == 0 and not args.no_secondary:
        secondary = secondary_configs(torch, L, _lib, dev, stream, fp32_peak_tflops, ptrs, N, M)
        secondary["c5_strong_B256"] = c5
    if world > 1:
        dist.barrier()

    # ---- end to end, headline: the host-buffer call of the C ABI (fpsg_chamfer_pipeline_*, fpsg_b200.ChamferPipeline): every
    # step copies ITS inputs from pinned host memory (H2D, 1.57 MB), runs the fused loss step and reads its losses back
    # (D2H); up to DEPTH steps are in flight (one lane each), so the copies of step i+1 overlap the kernels of step i ----
    host = []
    for s_ in range(8):
        h1, h2 = make_inputs(B, N, M, seed=777 + 1000 * rank + s_, pin=True)
        host.append((h1, h2))
    DEPTH = int(os.environ.get("FPSG_BENCH_DEPTH", "6"))  # steps in flight INCLUDING those whose host->device copy is still running: four in the kernels + two copying
    # (measured: 261 / 494 / 640 / 715 / 732 k pairs/s at depth 1 / 2 / 3 / 4 / 6, profiles/r02b_e2e_depth.json; depth 4 varies 450-735 k from run to run, depth 6 737-750 k:
    # profiles/r02d_e2e_trigger.jsonl)
    pipe = fpsg_b200.ChamferPipeline(B, N, M, device=dev, depth=DEPTH)
    e2e_exchange = None
    if exchange is not None:  # N > 1: every step of the pipeline takes part in the in-kernel scalar all-reduce (lag = DEPTH)
        from fpsg_b200.sharded import PeerScalarExchange
        e2e_exchange = PeerScalarExchange(dev)
        pipe.set_exchange(e2e_exchange)
    e2e_steps = max(20, min(args.steps, 1000))

    def e2e_loop(nsteps):
        sub = 0
        acc = 0.0
        for i in range(nsteps):
            while sub < nsteps and pipe.in_flight < DEPTH:
                pipe.submit(*host[sub % len(host)])
                sub += 1
            s_, _loss, _gp1 = pipe.result()   # this step's loss sum, per-pair losses (host) and gradient (device)
            acc += s_
            if e2e_exchange is not None and i >= 1 and (i & 63) == 0:
                acc += 0.0 * pipe.previous_global_sum()  # the all-reduced loss of the previous step, as a logger would read it
        return acc
    # The loop is host-driven and one pass lasts only ~45 ms: three passes, the median is reported and all three are listed
    # (a single pass right after the other sections of this script varied between 0.56 and 0.73 M pairs/s from run to run)
    e2e_loop(200)
    e2e_reps = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        barrier()
        e2e_reps.append(time.perf_counter() - t0)
    e2e_s = sorted(e2e_reps)[1]
    pipe.close()

    # ---- the floor under e2e: the same host->device copies alone (all ranks at once, nothing else running). On an 8-GPU box the
    # host side delivers ~200 GB/s in total (25 GB/s per GPU), so the copies of a step can take longer than its kernels; the line reports both ----
    h2d_dst = [torch.empty_like(host[0][0], device=dev), torch.empty_like(host[0][1], device=dev)]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_iters = 300

    def h2d_only(n):
        with torch.cuda.stream(copy_stream):
            for i_ in range(n):
                h1_, h2_ = host[i_ % len(host)]
                h2d_dst[0].copy_(h1_, non_blocking=True)
                h2d_dst[1].copy_(h2_, non_blocking=True)
        copy_stream.synchronize()
    h2d_only(50)
    barrier()
    t0 = time.perf_counter()
    h2d_only(h2d_iters)
    barrier()
    h2d_s = time.perf_counter() - t0

    # ---- end to end through the reference-facing Python call with autograd: kaolin shim chamfer_distance(p1, p2).sum()
    # .backward(), one H2D copy per step from pinned memory, loss read back, stream synchronised every step ----
    fpsg_b200.shims.install(force=True)
    from kaolin.metrics.pointcloud import chamfer_distance as shim_chamfer
    hcat = [torch.cat([h1.reshape(-1), h2.reshape(-1)]).pin_memory() for h1, h2 in host]
    n1 = B * N * 3
    dbuf = torch.empty(hcat[0].numel(), device=dev)
    p1_leaf = dbuf[:n1].view(B, N, 3).requires_grad_(True)   # the decoder output's stand-in: a leaf that wants a gradient
    p2_view = dbuf[n1:].view(B, M, 3)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def autograd_step(i):
        with torch.no_grad():
            dbuf.copy_(hcat[i % len(hcat)], non_blocking=True)
        p1_leaf.grad = None
        val = shim_chamfer(p1_leaf, p2_view).sum()
        val.backward()
        loss_host.copy_(val.detach(), non_blocking=True)
        stream.synchronize()
        return float(loss_host)
    ag_steps = max(20, min(args.steps, 300))
    for i in range(5):
        autograd_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(ag_steps):
        autograd_step(5 + i)
    barrier()
    ag_s = time.perf_counter() - t0

    # ---- reduce over ranks (max time) ----
    if world > 1:
        t = torch.tensor([ms, e2e_s, ag_s, ms_nocoll, ms_lag0, ms_serial, h2d_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, ag_s, ms_nocoll, ms_lag0, ms_serial, h2d_s = (float(v) for v in t)

    if rank == 0:
        pairs = world * B
        value = pairs * args.steps / (ms * 1e-3)
        flops = 16.0 * B * N * M  # 2*N*M evaluations x 8 FLOP per pair
        achieved = flops / (sweep_us * 1e-6) / 1e12
        step_us = ms / args.steps * 1e3
        serial_us = ms_serial / args.steps * 1e3
        step_tf = flops / (step_us * 1e-6) / 1e12
        bwd_bytes = (36 * N + 24 * M) * B
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "global_pairs": pairs, "n_points": N, "m_points": M,
                       "parallelism": f"batch-sharded x{world}", "collective": collective,
                       "l2": f"inputs rotate over {RING_SETS} distinct resident batches ({RING_SETS * B * (N + M) * 12 / 1e6:.0f} MB > 126 MB L2)",
                       "launch": f"C ABI via ctypes (fpsg_chamfer_multi_loss_fwd_bwd[_dist]: sweep + finish kernels), no CUDA graph; the steps work on "
                                 f"independent batches and are issued round-robin over {LANES} streams, each with its own workspace and outputs, so the "
                                 f"latency-bound finish kernel of step i overlaps the sweep of step i+1 (stream order inside a lane); "
                                 f"value_single_stream / latency_us_per_step = the same steps strictly one after the other on one stream",
                       "lanes": LANES,
                       "parity": "Chamfer oracle is parity-unpinned (Kaolin 0.9.0 absent): bit-exact vs the restated algorithm; kNN and input pipeline pinned"},
            "e2e": {"value": pairs * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": B * (N + M) * 12, "d2h_bytes_per_step": (2 + B + _lib.PEER_RING) * 4,
                    "steps": e2e_steps, "passes_pairs_per_s": [pairs * e2e_steps / t_ for t_ in e2e_reps], "reported": "median of three passes",
                    "h2d_copies_alone": {"gb_per_s_per_gpu": B * (N + M) * 12 * h2d_iters / h2d_s / 1e9, "us_per_step": h2d_s / h2d_iters * 1e6,
                                         "pairs_per_s_ceiling": pairs * h2d_iters / h2d_s,
                                         "what": "the same pinned-host -> device copies with nothing else running, all ranks at once (max over ranks): "
                                                 "when this exceeds ms_per_step the host link, not the kernels, bounds e2e"},
                    "api": "fpsg_chamfer_pipeline_submit / _result (include/fpsg_b200.h; fpsg_b200.ChamferPipeline): every step copies its own inputs "
                           "from pinned HOST memory, runs sweep + finish and reads its per-pair losses and their sum back to the host (with the ring of all-reduced sums); depth " + str(DEPTH) + ", so the "
                           "H2D copies of step i+1 overlap the kernels of step i; the gradient stays on the device as in the reference"},
            "e2e_autograd": {"value": pairs * ag_steps / ag_s, "unit": "pairs/s", "h2d_bytes_per_step": B * (N + M) * 12, "d2h_bytes_per_step": 4,
                             "steps": ag_steps,
                             "api": "kaolin.metrics.pointcloud.chamfer_distance(p1, p2).sum().backward() (fpsg_b200 shim, torch.autograd); every step: one "
                                    "H2D copy from pinned host memory, forward, backward, loss read back, stream synchronised (host-bound)"},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": {"bound": "fp32_fma", "kernel": "chamfer_sweep_kernel<16>", "achieved": achieved, "peak": fp32_peak_tflops,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops, "traffic": SWEEP_DRAM_TRAFFIC_BYTES,
                         "traffic_source": SWEEP_DRAM_TRAFFIC_SOURCE,
                         "us_per_launch": sweep_us, "algorithmic_flop_per_launch": flops,
                         "peak_source": f"FFMA micro-kernel in this run (fpsg_probe_fp32_peak) at {probe_mhz:.0f} MHz; MEASURED_PEAKS.json has no FP32 row; "
                                        f"nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                         "how": "sweep kernel launched alone (fpsg_chamfer_sweep_only) over the same input ring, CUDA events on the launch stream",
                         "overlapped": {"us_per_launch": sweep_lanes_us, "achieved": flops / (sweep_lanes_us * 1e-6) / 1e12,
                                        "frac": flops / (sweep_lanes_us * 1e-6) / 1e12 / fp32_peak_tflops,
                                        "how": f"the same launches alternating over the {LANES} lanes of the timed step: ramp and tail of "
                                               f"consecutive launches overlap (total time / launches)"}},
            "roofline_step": {"bound": "fp32_fma", "what": "the whole fwd+bwd step (the north-star target: >= 0.60), throughput over independent "
                                                             f"batches issued over {LANES} lanes", "achieved": step_tf,
                              "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": step_tf / fp32_peak_tflops, "frac_of_nominal_74.4": step_tf / 74.4,
                              "us_per_step": step_us, "algorithmic_flop_per_step": flops,
                              "single_stream": {"us_per_step": serial_us, "achieved": flops / (serial_us * 1e-6) / 1e12,
                                                "frac": flops / (serial_us * 1e-6) / 1e12 / fp32_peak_tflops,
                                                "what": "one step at a time on one stream (step i+1 starts when step i has finished): the step's latency"}},
            "value_single_stream": pairs * args.steps / (ms_serial * 1e-3),
            "latency_us_per_step": serial_us,
            "roofline_bwd": {"bound": "hbm", "kernel": "chamfer_bwd_kernel (autograd path; the bench step runs the backward inside the finish kernel)",
                             "achieved": bwd_bytes / (bwd_us * 1e-6) / 1e9, "peak": _hbm_peak(),
                             "unit": "GB/s", "us_per_launch": bwd_us, "algorithmic_bytes_per_launch": bwd_bytes,
                             "note": "3.9 MB working set is L2-resident: latency-, not bandwidth-bound (SURVEY.md 8d)"},
            "clocks": clocks,
            "secondary": secondary,
        }
        if world > 1:
            line["value_without_collective"] = pairs * args.steps / (ms_nocoll * 1e-3)
            line["value_collective_in_step"] = pairs * args.steps / (ms_lag0 * 1e-3)
        line["roofline_bwd"]["frac"] = line["roofline_bwd"]["achieved"] / line["roofline_bwd"]["peak"]
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sec = cpu_formulation_pairs_per_s(sample_pairs=8, iters=3)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"8 of the 32 pairs, best of 3 iterations ({sec:.2f} s each), pure-PyTorch difference-form "
                                              f"Chamfer fwd+bwd (oracle/torch_formulation.py; Kaolin 0.9.0 is absent), os.cpu_count()={os.cpu_count()}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0  # the profiling recipe's fallback


if __name__ == "__main__":
    main()
