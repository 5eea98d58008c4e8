#!/usr/bin/env python
"""bench.py -- headline benchmark of fpsg_b200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): Chamfer fwd+bwd pairs/s at B=32, N=M=2048 per GPU (weak scaling over 1/2/4/8 B200),
plus the dominant kernel's fraction of the FP32 FMA roofline.

A *step* = one pass of the hot path over one batch of 32 synthetic cloud pairs: fused Chamfer loss forward
(sweep + finalize + loss kernels), backward to p1 (the ground truth has no gradient, as in the reference's
training loss, src/models/few_shot.py:109-117), and -- for N > 1 -- the scalar loss all-reduce over NCCL.

  value : whole-job pairs/s, inputs resident in HBM, the C ABI called directly, CUDA events, max over ranks.
  e2e   : the same metric through the reference-facing Python call (the kaolin shim's
          chamfer_distance(p1, p2).sum().backward()) with HOST (pinned) inputs: H2D copies and the D2H read of
          the loss are inside the timed region.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_PTS, M_PTS = 32, 2048, 2048
WORKLOAD = "c1: Chamfer loss fwd+bwd(p1), B=32 per GPU, N=M=2048, synthetic xyz float32"
METRIC = "chamfer_fwd_bwd_pairs_per_s"
RING_SETS = 96  # distinct input batches cycled through: 96 x 1.57 MB = 151 MB > 126 MB L2
# dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch (profiles/r01b_chamfer_kernels_raw.csv, r01_summary.md)
SWEEP_DRAM_TRAFFIC_BYTES = 1.599e6


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
        "config": {"workload": WORKLOAD, "note": f"each step is a bounded sample of {sample} of the 32 pairs; value = pairs/s"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} pairs x {args.steps} steps of the c1 workload, pure-PyTorch difference-form Chamfer fwd+bwd "
                                   f"(oracle/torch_formulation.py; Kaolin 0.9.0 itself is absent), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the bounded CPU-baseline leg")
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
    dist1 = torch.empty(B, N, device=dev)
    dist2 = torch.empty(B, M, device=dev)
    idx1 = torch.empty(B, N, dtype=torch.int64, device=dev)
    idx2 = torch.empty(B, M, dtype=torch.int64, device=dev)
    loss = torch.empty(B, device=dev)
    gloss = torch.ones(B, device=dev)
    gp1 = torch.empty(B, N, 3, device=dev)
    # N > 1: the per-step sums of the local losses (logged values; the backward does not need them) are collected in a
    # device-side ring of KLOG slots and all-reduced KLOG steps at a time -- torch.distributed's host-side cost of one
    # collective (~45 us) would otherwise make the 64-us step host-bound. Two rings alternate.
    KLOG = 8
    totals = [torch.zeros(KLOG, 2, device=dev) for _ in range(2)]  # [:, 0] = sum over the local pairs of one step
    pending = [None, None]
    nbytes = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in ring]
    o = (loss.data_ptr(), dist1.data_ptr(), dist2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(), ws.data_ptr(), gloss.data_ptr(), gp1.data_ptr())

    # N > 1: the local sum of the per-pair losses comes out of the finalize kernel itself (the one-segment form of
    # fpsg_chamfer_multi_loss_fwd), so no reduction kernel sits between the forward and the backward
    segs = []
    for a_, b_ in ptrs:
        sg = (_lib.ChamferSegment * 1)()
        sg[0].p1, sg[0].p2, sg[0].gp1, sg[0].gp2, sg[0].batch, sg[0].weight = a_, b_, None, None, B, 1.0
        segs.append(sg)
    tot_ptrs = [[t.data_ptr() + 8 * j for j in range(KLOG)] for t in totals]

    def step(i):
        a, b = ptrs[i % RING_SETS]
        if world == 1:
            _lib.check(L.fpsg_chamfer_loss_fwd(a, b, B, N, M, 1.0, 1.0, o[0], o[1], o[2], o[3], o[4], o[5], nbytes, sp))
        else:
            k, j = (i // KLOG) & 1, i % KLOG
            if j == 0 and pending[k] is not None:
                # the all-reduce that last used this ring (KLOG steps ago) has normally finished long ago: a host-side
                # query then suffices and no wait goes into the compute stream
                if not pending[k].is_completed():
                    pending[k].wait()
                pending[k] = None
            _lib.check(L.fpsg_chamfer_multi_loss_fwd(segs[i % RING_SETS], 1, N, M, 1.0, 1.0, o[0], tot_ptrs[k][j], o[3], o[4], o[5], nbytes, sp))
        _lib.check(L.fpsg_chamfer_loss_bwd(a, b, o[3], o[4], o[6], 1.0, 1.0, B, N, M, o[7], None, None, 0, sp))
        if world > 1:
            # issued after the backward launch: an event record between the finalize and the backward kernels would break
            # their programmatic (PDL) chaining
            if j == KLOG - 1:
                pending[k] = dist.all_reduce(totals[k], op=dist.ReduceOp.SUM, async_op=True)

    kernels_per_step = 3  # sweep, finalize (+ fused loss and batch sum), backward (the NCCL kernel for N > 1 is not ours)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)  # same parity/ring phase as the warm-up left it
    e1.record(stream)
    for k in (0, 1):
        if pending[k] is not None:
            pending[k].wait()
            pending[k] = None
    barrier()
    sampler.mark(tw0, time.perf_counter())
    ms = e0.elapsed_time(e1)

    # ---- dominant kernel alone (roofline), same ring, same stream ----
    def sweep(i):
        a, b = ptrs[i % RING_SETS]
        _lib.check(L.fpsg_chamfer_sweep_only(a, b, B, N, M, o[5], nbytes, sp))
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

    # ---- backward kernel alone (HBM view) ----
    def bwd(i):
        a, b = ptrs[i % RING_SETS]
        _lib.check(L.fpsg_chamfer_loss_bwd(a, b, o[3], o[4], o[6], 1.0, 1.0, B, N, M, o[7], None, None, 0, sp))
    for i in range(5):
        bwd(i)
    torch.cuda.synchronize(dev)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for i in range(n_sweeps):
        bwd(5 + i)
    b1.record(stream)
    torch.cuda.synchronize(dev)
    bwd_us = b0.elapsed_time(b1) / n_sweeps * 1e3
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary configurations of BASELINE.json (not the metric; reported for context, rank 0 only) ----
    secondary = None
    if rank == 0:
        def timed(fn, iters=50, warm=5):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize(dev)
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record(stream)
            for _ in range(iters):
                fn()
            t1_.record(stream)
            torch.cuda.synchronize(dev)
            return t0_.elapsed_time(t1_) / iters * 1e3
        secondary = {}
        gk = torch.Generator(device="cpu").manual_seed(1234 + 3)
        for C_ in (3, 64):  # c3: DGCNN kNN graph, k=20, B=32, N=1024 (xyz layer; a 64-channel layer for scale)
            xk = torch.randn(32, C_, 1024, generator=gk).to(dev)
            ik = torch.empty(32, 1024, 20, dtype=torch.int64, device=dev)
            nbk = L.fpsg_knn_workspace_bytes(32, C_, 1024, 20)
            wsk = torch.empty(max(nbk, 1), dtype=torch.uint8, device=dev)
            us = timed(lambda: _lib.check(L.fpsg_knn_fwd(xk.data_ptr(), 32, C_, 1024, 20, ik.data_ptr(), wsk.data_ptr(), nbk, sp)), 20, 3)
            secondary[f"knn_k20_B32_N1024_C{C_}"] = {"us_per_call": us, "tflops": (2 * C_ + 2) * 32 * 1024 * 1024 / us / 1e6,
                                                    "frac_of_fp32_peak": (2 * C_ + 2) * 32 * 1024 * 1024 / us / 1e6 / fp32_peak_tflops}
            del xk, ik, wsk
        # c2: one training episode (5 query + 32 support pairs) as the reference issues it (two calls) and as one call
        qa, qb = ptrs[0]
        sg2 = (_lib.ChamferSegment * 2)()
        gp37 = torch.empty(37, N, 3, device=dev)
        sg2[0].p1, sg2[0].p2, sg2[0].gp1, sg2[0].gp2, sg2[0].batch, sg2[0].weight = qa, qb, gp37.data_ptr(), None, 5, 1.0
        sg2[1].p1, sg2[1].p2, sg2[1].gp1, sg2[1].gp2, sg2[1].batch, sg2[1].weight = (ptrs[1][0], ptrs[1][1], gp37.data_ptr() + 5 * N * 12, None,
                                                                                   32, 1.0)
        nb37 = L.fpsg_chamfer_fwd_workspace_bytes(37, N, M)
        ws37 = torch.empty(nb37, dtype=torch.uint8, device=dev)
        i37 = torch.empty(37 * (N + M), dtype=torch.int64, device=dev)
        l37 = torch.empty(37 + 3, device=dev)
        g1_ = torch.ones((), device=dev)

        def two_calls():
            for (a_, b_), nb_ in ((ptrs[0], 5), (ptrs[1], 32)):
                _lib.check(L.fpsg_chamfer_loss_fwd(a_, b_, nb_, N, M, 1.0, 1.0, o[0], None, None, o[3], o[4], o[5], nbytes, sp))
                _lib.check(L.fpsg_chamfer_loss_bwd(a_, b_, o[3], o[4], o[6], 1.0, 1.0, nb_, N, M, o[7], None, None, 0, sp))

        def one_call():
            _lib.check(L.fpsg_chamfer_multi_loss_fwd(sg2, 2, N, M, 1.0, 1.0, l37.data_ptr() + 12, l37.data_ptr(), i37.data_ptr(),
                                                     i37.data_ptr() + 8 * 37 * N, ws37.data_ptr(), nb37, sp))
            _lib.check(L.fpsg_chamfer_multi_loss_bwd(sg2, 2, i37.data_ptr(), i37.data_ptr() + 8 * 37 * N, g1_.data_ptr(), 1.0, 1.0, N, M, sp))
        secondary["c2_episode_5q_32s"] = {"two_calls_us": timed(two_calls), "multi_segment_us": timed(one_call)}

    # ---- end to end: the reference-facing call, host (pinned) inputs, loss read back every step ----
    fpsg_b200.shims.install(force=True)
    from kaolin.metrics.pointcloud import chamfer_distance as shim_chamfer
    # Every step copies its own inputs from pinned host memory (one buffer per batch holding p1 then p2: a single H2D
    # copy), runs the reference-facing call, and reads its own loss back before the next step is issued. The copy
    # goes on the compute stream: the step is bound by the host side of autograd (bench_micro/e2e_profile.py), so a
    # separate copy stream only adds host work.
    host = []
    for s_ in range(8):
        h1, h2 = make_inputs(B, N, M, seed=777 + 1000 * rank + s_)
        host.append(torch.cat([h1.reshape(-1), h2.reshape(-1)]).pin_memory())
    n1 = B * N * 3
    dbuf = torch.empty(host[0].numel(), device=dev)
    p1_leaf = dbuf[:n1].view(B, N, 3).requires_grad_(True)   # the decoder output's stand-in: a leaf that wants a gradient
    p2_view = dbuf[n1:].view(B, M, 3)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    cur = torch.cuda.current_stream(dev)

    # N > 1: the global sum of the loss is a logged value only (the backward does not need it): as in the device-timed
    # loop above, the per-step local sums are collected in a device-side ring and all-reduced KLOG steps per collective
    # (one torch.distributed collective costs ~45 us of host time, the whole step ~160); ranks never wait for each other
    # inside a step. What every step reads back and synchronises on is its own loss.
    e2e_ring = [torch.zeros(KLOG, device=dev) for _ in range(2)]
    e2e_work = [None, None]

    def e2e_step(i):
        with torch.no_grad():
            dbuf.copy_(host[i % len(host)], non_blocking=True)
        p1_leaf.grad = None
        val = shim_chamfer(p1_leaf, p2_view).sum()
        val.backward()
        loss_host.copy_(val.detach(), non_blocking=True)
        if world > 1:
            kk, j = (i // KLOG) & 1, i % KLOG
            if j == 0 and e2e_work[kk] is not None:
                if not e2e_work[kk].is_completed():
                    e2e_work[kk].wait()
                e2e_work[kk] = None
            e2e_ring[kk][j].copy_(val.detach())
            if j == KLOG - 1:
                e2e_work[kk] = dist.all_reduce(e2e_ring[kk], op=dist.ReduceOp.SUM, async_op=True)
        cur.synchronize()  # the step's result (and its gradient) exist before the next step is issued
        return float(loss_host)

    e2e_steps = max(20, min(args.steps, 400))
    for i in range(5):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(5 + i)
    for kk in (0, 1):
        if e2e_work[kk] is not None:
            e2e_work[kk].wait()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- the same step through the C ABI itself with host buffers (what a native binding of the reference gets, without
    # the autograd engine): H2D copy of the step's inputs from pinned memory, fused loss forward with the batch sum,
    # backward to p1, D2H read of the scalar, stream synchronised -- every step. (Copying the next step's inputs on a
    # second stream was tried and changes nothing: the step is host issue time + device time, not copy time.) ----
    cabi_tot = torch.zeros(2, device=dev)
    cabi_seg = (_lib.ChamferSegment * 1)()
    cabi_seg[0].p1, cabi_seg[0].p2, cabi_seg[0].gp1, cabi_seg[0].gp2 = dbuf.data_ptr(), dbuf.data_ptr() + 4 * n1, None, None
    cabi_seg[0].batch, cabi_seg[0].weight = B, 1.0
    p1d, p2d = dbuf.data_ptr(), dbuf.data_ptr() + 4 * n1

    def cabi_step(i):
        dbuf.copy_(host[i % len(host)], non_blocking=True)
        _lib.check(L.fpsg_chamfer_multi_loss_fwd(cabi_seg, 1, N, M, 1.0, 1.0, o[0], cabi_tot.data_ptr(), o[3], o[4], o[5], nbytes, sp))
        _lib.check(L.fpsg_chamfer_loss_bwd(p1d, p2d, o[3], o[4], o[6], 1.0, 1.0, B, N, M, o[7], None, None, 0, sp))
        loss_host.copy_(cabi_tot[0], non_blocking=True)
        cur.synchronize()
        return float(loss_host)

    with torch.no_grad():
        for i in range(5):
            cabi_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            cabi_step(5 + i)
        barrier()
        cabi_s = time.perf_counter() - t0

    # ---- reduce over ranks (max time) ----
    if world > 1:
        t = torch.tensor([ms, e2e_s, cabi_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, cabi_s = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        pairs = world * B
        value = pairs * args.steps / (ms * 1e-3)
        e2e_value = pairs * e2e_steps / e2e_s
        flops = 16.0 * B * N * M  # 2*N*M evaluations x 8 FLOP per pair
        achieved = flops / (sweep_us * 1e-6) / 1e12
        bwd_bytes = (36 * N + 24 * M) * B
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "global_pairs": pairs, "n_points": N, "m_points": M,
                       "parallelism": f"batch-sharded x{world}" + (", scalar loss all-reduce (NCCL, 8 steps per collective)" if world > 1 else ""),
                       "l2": f"inputs rotate over {RING_SETS} distinct resident batches ({RING_SETS * B * (N + M) * 12 / 1e6:.0f} MB > 126 MB L2)",
                       "launch": "C ABI via ctypes, stream-ordered, no CUDA graph"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": B * (N + M) * 12, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "api": "kaolin.metrics.pointcloud.chamfer_distance(p1, p2).sum().backward() (fpsg_b200 shim); every step: one H2D copy of its inputs from pinned host memory, forward, backward, loss read back to the host, stream synchronised"},
            "e2e_c_abi": {"value": pairs * e2e_steps / cabi_s, "unit": "pairs/s", "h2d_bytes_per_step": B * (N + M) * 12, "d2h_bytes_per_step": 4,
                          "steps": e2e_steps,
                          "api": "fpsg_chamfer_multi_loss_fwd + fpsg_chamfer_loss_bwd through ctypes (no autograd); every step: one H2D copy of "
                                 "its inputs from pinned host memory, forward, backward, loss sum read back to the host, stream synchronised"},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": {"bound": "fp32_fma", "kernel": "chamfer_sweep_kernel<16,false>", "achieved": achieved, "peak": fp32_peak_tflops,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops, "traffic": SWEEP_DRAM_TRAFFIC_BYTES,
                         "us_per_launch": sweep_us, "algorithmic_flop_per_launch": flops,
                         "peak_source": f"FFMA micro-kernel in this run (fpsg_probe_fp32_peak) at {probe_mhz:.0f} MHz; MEASURED_PEAKS.json has no FP32 row; "
                                        f"nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                         "how": "sweep kernel launched alone (fpsg_chamfer_sweep_only) over the same input ring, CUDA events on the launch stream"},
            "roofline_bwd": {"bound": "hbm", "kernel": "chamfer_bwd_kernel", "achieved": bwd_bytes / (bwd_us * 1e-6) / 1e9, "peak": _hbm_peak(),
                             "unit": "GB/s", "us_per_launch": bwd_us, "algorithmic_bytes_per_launch": bwd_bytes,
                             "note": "3.9 MB working set is L2-resident: latency-, not bandwidth-bound (SURVEY.md 8d)"},
            "clocks": clocks,
            "secondary": secondary,
        }
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
