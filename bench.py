#!/usr/bin/env python
"""Headline benchmark: 1080p warp + DeformConv2d path, frames/s (BASELINE.json `metric`, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path (one JSON line)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...              # one rank per GPU, weak scaling

A "step" is one pass of the hot path (/root/reference/src/models/ema_vfi.py:130-138: warp, concat, 3 x DCNv2 with the
offset/mask split) over one batch of 8 synthetic 1080p frame pairs per GPU, bf16 tensors.

* `value`      frames/s over all ranks with every input already resident in HBM (CUDA events, max over ranks).
* `e2e`        the same metric through HotPath.run_host: pinned HOST buffers in and out, H2D/D2H inside the timed region.
* `roofline`   the dominant kernel (DCNv2 forward): algorithmic 2*P*603*67 FLOP per launch / mean launch duration measured
               with CUDA events around every launch inside the timed region, against MEASURED_PEAKS.json.
* `cpu_baseline` the reference's CPU path (stock torch grid_sample + torchvision deform_conv2d CPU kernels driven by
               oracle/torch_ref.py) on a bounded stripe of the same workload, rank 0, N=1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (batch per GPU, H, W, flow sigma px, offset sigma px, description)
    "cfg2": (8, 1080, 1920, 8.0, 1.5, "cfg2: 1080p (1920x1080) frame pairs, batch 8 per GPU, bf16, warp + concat + 3 x DCNv2 fwd"),
    "cfg4": (1, 2160, 3840, 64.0, 1.5, "cfg4: 4K (3840x2160) frame pair, batch 1, bf16, large-displacement flow"),
    "cfg1": (1, 256, 256, 8.0, 1.5, "cfg1 geometry: 256x256, batch 1 (hot path only)"),
    # cfg2's shapes; the e2e leg streams STREAM_PAIRS independent pairs from pinned host memory, sharded over the ranks
    "cfg5": (8, 1080, 1920, 8.0, 1.5, "cfg5: stream of 256 independent 1080p frame pairs in batches of 8, sharded over the ranks, bf16"),
}
# SURVEY 8d (iii), worst-case locality: per-pixel independent displacements (not timed in round 1)
# BASELINE config 3: training step (forward + backward + gradient all-reduce), handled by run_cfg3 below
WORKLOADS["cfg3"] = (16, 256, 256, 2.0, 1.5, "cfg3: training step at crop 256x256, global batch 16: warp + concat + 3 x (offset_conv, DCNv2) "
                     "forward + backward, gradient all-reduce over the ranks")
WORKLOADS["cfg4_iid"] = (1, 2160, 3840, 64.0, 1.5, "cfg4, incoherent variant: 4K frame pair, batch 1, bf16, iid N(0, 64^2) px flow")
FLOW_KIND = {"cfg4_iid": "iid"}          # everything else: the smooth field of hotpath.synthetic_inputs
STREAM_PAIRS = 256
FLOP_PER_PX = 2 * 603 * 67          # one DCNv2 layer, algorithmic; padding not counted
WARP_BYTES_PER_PX_BF16 = (3 + 2 + 3) * 2


def csrc_digest() -> str:
    """sha256 over the kernel sources: what a committed ncu capture must have been taken from to describe this build."""
    import hashlib

    h = hashlib.sha256()
    for f in sorted((ROOT / "video-frame-interpolation_b200" / "csrc").glob("*.cu*")) + sorted((ROOT / "video-frame-interpolation_b200" / "csrc").glob("*.h")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per launch of the hot kernels from the committed `ncu --set full` captures (profiles/traffic.json, written by
    scripts/update_traffic.py).  An entry counts only while the kernel sources are the ones it was captured from; otherwise the
    bench line says null + why (the number is a property of a build, not of the benchmark run)."""
    f = ROOT / "profiles" / "traffic.json"
    d = json.loads(f.read_text()) if f.exists() else {}
    now, out = csrc_digest(), {}
    for k, v in d.items():
        if isinstance(v, dict) and v.get("csrc_digest") == now:
            out[k] = v["dram_bytes_per_launch"]
            out[k + "_source"] = f"{v.get('report')} @ {v.get('commit')} (csrc {now})"
        elif isinstance(v, dict):
            out[k + "_source"] = f"stale: captured from csrc {v.get('csrc_digest')} @ {v.get('commit')}, tree is {now}"
    return out


def measured_gather_floor():
    """scripts/microbench/gather_floor.cu on this pool (profiles/r02_*gather_floor*.jsonl): the producers' work of the DCN forward
    alone -- entry reads, 36 x LDS.128 per pixel, blend, tcgen05.st, box copies -- in clocks per output pixel."""
    best = None
    for f in sorted((ROOT / "profiles").glob("r02_*gather_floor*.jsonl")):
        for ln in f.read_text().splitlines():
            try:
                r = json.loads(ln)
            except ValueError:
                continue
            if r.get("mode") == "gather+entry+lerp+tmem_st+box_copy" and (best is None or r["clk_per_px"] < best["clk_per_px"]):
                best = dict(r, file=f.name)
    return best


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]),
                    tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, 20 ms period)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop, self._thr = [], set(), None, threading.Event(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))))
                try:
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def stop(self, t0=None, t1=None):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        vals = [v for (t, v) in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        if len(vals) < 3:
            vals = [v for _, v in self.samples]
        if not vals:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(statistics.median(vals)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(vals)}


# --------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_factory(H_rows: int, W: int, seed: int = 1234):
    """The reference's CPU path on a bounded stripe (H_rows x W, batch 1, fp32 as the reference runs on CPU)."""
    import torch

    from oracle import torch_ref
    from vfi_b200.hotpath import synthetic_inputs, synthetic_weights

    torch.set_num_threads(os.cpu_count() or 1)
    frame2, flow, feat, convs = synthetic_inputs(1, H_rows, W, dtype=torch.float32, device="cpu", seed=seed)
    ws, bs = synthetic_weights(dtype=torch.float32, device="cpu")

    def step():
        with torch.no_grad():
            return torch_ref.hot_path(frame2, flow, feat, convs, ws, bs)

    return step


def time_cpu(step, reps: int, warm: int):
    for _ in range(warm):
        step()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t)
    return ts


def run_reference_arm(args, desc, H, W):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    rows = 32 if (args.steps + args.warmup) <= 30 else 16
    step = cpu_reference_step_factory(rows, W)
    ts = time_cpu(step, args.steps, args.warmup)
    frames_per_step = rows * W / float(H * W)
    total = sum(ts)
    value = frames_per_step * len(ts) / total
    cores = os.cpu_count() or 1
    sample = (f"{rows}x{W} stripe of one {W}x{H} frame (batch 1, fp32, {rows * W} px = {frames_per_step:.5f} frame) per step; "
              "stock torch grid_sample + torchvision deform_conv2d CPU kernels (the reference's own third-party ops) "
              "driven by oracle/torch_ref.py; frames/s scaled linearly in pixels "
              "(checked against one full 1080p frame on this pool: 50.1 s, the extrapolation is 4 % conservative -- profiles/r02_s_cpu_full_frame.json)")
    line = {"impl": "reference", "metric": "1080p warp+DeformConv path frames/sec", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- cfg5: frames in, frames out
def e2e_stream(args, topo, dev, H, W):
    """BASELINE config 5 end to end: STREAM_PAIRS independent 1080p frame pairs as uint8 frames in pinned host memory -> interpolated
    uint8 frames back on the host, through vfi_b200.stream.PairStreamer and the network (vfi_b200.refmodel.StockInterpolator: the
    reference's layers as stock PyTorch, hot path through install(fuse=True)), the pairs sharded contiguously over the ranks in
    batches of 8.  uint8 frames are the only PCIe traffic (6.2 MB per frame each way)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import vfi_b200
    from vfi_b200 import stream
    from vfi_b200.refmodel import StockInterpolator

    torch.manual_seed(2026)
    model = StockInterpolator().eval()
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for blk in model.attention_blocks:       # offset_conv is zero-initialised; randomise it so the gather is exercised
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
        model.motion_estimation[-1].weight.mul_(40.0)
    model = model.to(dev)
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    n = STREAM_PAIRS + 1
    frames = [np.roll(base, 3 * i, axis=1) for i in range(n)]
    vfi_b200.install(StockInterpolator, fuse=True)
    try:
        ps = stream.PairStreamer(model, dev, batch_pairs=8, topology=topo, autocast_dtype=torch.float16)
        list(ps.run(frames[:17]))                            # warm-up
        torch.cuda.synchronize(dev)
        if topo.world > 1:
            dist.barrier()
        h0, d0 = ps.stats["h2d_bytes"], ps.stats["d2h_bytes"]
        t0 = time.perf_counter()
        written = sum(1 for _ in ps.run(frames))
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        calls = vfi_b200.dropin.call_counts()
    finally:
        vfi_b200.uninstall()
    h2d, d2h = ps.stats["h2d_bytes"] - h0, ps.stats["d2h_bytes"] - d0
    if topo.world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    steps = max(1, -(-(STREAM_PAIRS // topo.world) // 8))
    return {"value": STREAM_PAIRS / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d // steps), "d2h_bytes_per_step": int(d2h // steps),
            "steps": steps, "ms_per_step": 1e3 * dt / steps, "pairs": STREAM_PAIRS, "written_this_rank": written,
            "api": "vfi_b200.stream.PairStreamer over vfi_b200.refmodel.StockInterpolator with install(fuse=True): uint8 frames in pinned "
                   "host memory in, interpolated uint8 frames out (inference.py:160-199), no_grad + fp16 autocast",
            "seam_calls": calls}


# --------------------------------------------------------------------------------------------------- cfg3: training step
def run_cfg3(args, desc):
    """Strong scaling (the global batch of 16 is fixed): samples/s over all ranks, CUDA events, max over ranks; the exposed part of
    the gradient all-reduce is measured on rank 0 with events around the wait that follows backward."""
    import torch
    import torch.distributed as dist

    import vfi_b200
    from vfi_b200 import shard
    from vfi_b200.trainstep import TrainStep

    topo = shard.init_distributed()
    dev = torch.device("cuda", topo.local_rank)
    torch.cuda.set_device(dev)
    math = "bf16_tc" if args.math == "auto" else args.math
    ts = TrainStep(topo, dev, math=math, overlap=not args.no_overlap, fused=args.fused_blocks)
    if args.cuda_graph:
        ts.capture()
    for _ in range(args.warmup):
        ts.step()
    if topo.world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index or 0).start() if topo.is_root else None
    vfi_b200.reset_launch_count()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ts.step()
    e1.record()
    if topo.world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    launches = vfi_b200.launch_count()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms = e0.elapsed_time(e1) / args.steps
    if topo.world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    for _ in range(0 if args.cuda_graph else 5):         # separate, synchronising pass: how long the step waits for the exchange
        ts.step(measure=True)
    exposed_us = 1e3 * statistics.fmean(ts.exposed_ms) if ts.exposed_ms else 0.0
    ok = bool(torch.isfinite(ts.flow.grad).all()) and bool(torch.isfinite(ts.bucket.flat).all())
    if topo.is_root:
        line = {"metric": "cfg3 training step samples/sec", "value": ts.global_batch / (ms * 1e-3), "unit": "samples/s", "n_gpus": topo.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if math == "bf16_tc" else "f32", "data": "synthetic",
                "config": {"workload": desc, "global_batch": ts.global_batch, "batch_per_gpu": ts.B, "height": ts.H, "width": ts.W, "math": math,
                           "exchange": ("3 group all-reduces launched from autograd hooks (block 3 first), overlapped with backward"
                                        if ts.overlap else "one flat-bucket all-reduce after backward"),
                           "bucket_bytes": ts.bucket.numel * 4,
                           "launch": "one CUDA graph replay per step" if args.cuda_graph else "eager (one launch per kernel)",
                           "blocks": ("record activations [B,H,W,72] + fused DCN block (vfi_dcn_fwd_fused / vfi_dcn_bwd_*_fused)" if ts.fused
                                      else "the reference's glue as stock ops (chunk / cat / sigmoid) around vfi_dcn_fwd / vfi_dcn_bwd_*")},
                "gpu_launches": int(launches) if not args.cuda_graph else int(ts.graph_launches * args.steps), "allreduce_exposed_us": None if args.cuda_graph else exposed_us, "gradients_finite": ok, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if topo.world > 1:
        if args.cuda_graph:
            # Tearing the process group down while a captured graph still references its NCCL kernels hung for the full
            # timeout on an 8-GPU box (profiles/r02_w_*: the line above had been printed).  Drop the graph, drain the device and
            # leave without the collective teardown.
            ts.graph = None
            torch.cuda.synchronize(dev)
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vfi_b200", choices=["vfi_b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="auto", choices=["auto", "fp32", "bf16_tc"])
    ap.add_argument("--dcn-kernel", default="", choices=["", "v6"], help="A/B switch: v6 = the round-1 tcgen05 DCN kernel (default: v7)")
    ap.add_argument("--conv27-layout", default="nchw", choices=["nchw", "channels_last"],
                    help="memory format of the three offset_conv outputs the DCN layers read")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--fused-blocks", action="store_true",
                    help="cfg3: 72-channel record activations + ops.deform_conv2d_block (chunk / cat / sigmoid, layout and padding passes folded away, forward and backward)")
    ap.add_argument("--cuda-graph", action="store_true", help="cfg3: capture the whole step in a CUDA graph and replay it (SURVEY H7)")
    ap.add_argument("--no-overlap", action="store_true", help="cfg3: one all-reduce after backward instead of the overlapped group exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.dcn_kernel:
        os.environ["VFI_DCN_KERNEL"] = args.dcn_kernel      # read once by libvfi_b200 at the first DCN launch
    B, H, W, flow_sigma, off_sigma, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, desc, H, W)
        return
    if args.workload == "cfg3":
        run_cfg3(args, desc)
        return

    import torch
    import torch.distributed as dist

    import vfi_b200
    from vfi_b200 import shard
    from vfi_b200.hotpath import HotPath, synthetic_inputs, synthetic_weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    topo = shard.init_distributed()
    world = topo.world
    dev = torch.device("cuda", topo.local_rank)
    torch.cuda.set_device(dev)
    affinity = shard.bind_to_gpu(topo.local_rank)          # before any pinned allocation: host buffers land on this GPU's node
    dtype = torch.bfloat16
    P = B * H * W

    ws, bs = synthetic_weights(dtype=dtype, device=dev)
    path = HotPath(ws, bs, math=args.math)
    frame2, flow, feat, convs = synthetic_inputs(B, H, W, dtype=dtype, device=dev, seed=1234 + topo.rank,
                                                 flow_sigma=flow_sigma, offset_sigma=off_sigma,
                                                 flow_kind=FLOW_KIND.get(args.workload, "smooth"))
    feat = feat.contiguous(memory_format=torch.channels_last)
    if args.conv27_layout == "channels_last":
        convs = [c.contiguous(memory_format=torch.channels_last) for c in convs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # kernel-level events: every DCN forward launch and the warp launch inside the timed region
    dcn_ev, warp_ev = [], []
    orig_dcn, orig_dcn_fused, orig_warp = vfi_b200.ops.deform_conv2d, vfi_b200.ops.deform_conv2d_fused, vfi_b200.ops.warp
    timing = {"on": False}

    def timed(fn, bucket):
        def wrapper(*a, **k):
            if not timing["on"]:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            bucket.append((e0, e1))
            return r
        return wrapper

    vfi_b200.ops.deform_conv2d = timed(orig_dcn, dcn_ev)
    vfi_b200.ops.deform_conv2d_fused = timed(orig_dcn_fused, dcn_ev)
    vfi_b200.ops.warp = timed(orig_warp, warp_ev)

    out = None
    for _ in range(args.warmup):
        out = path.run(frame2, flow, feat, convs)
    barrier()
    sampler = ClockSampler(dev.index if dev.index is not None else 0).start() if topo.is_root else None
    vfi_b200.reset_launch_count()
    timing["on"] = True
    barrier()
    t_wall0 = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out = path.run(frame2, flow, feat, convs)
    end.record()
    barrier()
    t_wall1 = time.perf_counter()
    timing["on"] = False
    launches = vfi_b200.launch_count()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    elapsed_ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * B * args.steps / (elapsed_ms / 1e3)
    dcn_ms = statistics.fmean(a.elapsed_time(b) for a, b in dcn_ev) if dcn_ev else float("nan")
    warp_ms = statistics.fmean(a.elapsed_time(b) for a, b in warp_ev) if warp_ev else float("nan")
    del out

    # the warp kernel on its own (planar NCHW in/out, the form EMA_VFI.warp is called in): 20 launches, inputs 266 MB > L2
    planar_ms = {}
    for name, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
        f2, fl = frame2.to(dt), flow.to(dt)
        for _ in range(3):
            orig_warp(f2, fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(20):
            orig_warp(f2, fl)
        e1.record()
        torch.cuda.synchronize(dev)
        planar_ms[name] = e0.elapsed_time(e1) / 20
        del f2, fl

    # same kernel, SURVEY section 8d flow class (i): the flow a random-init EMA_VFI produces, N(0, 0.03^2) px.  The
    # headline flow above (class ii, 0.25 px/px of shear) spreads one request over ~7 source rows; this one streams.
    model_like = {}
    try:
        gm = torch.Generator(device=dev).manual_seed(77)
        fl32 = 0.03 * torch.randn(B, 2, H, W, device=dev, generator=gm)
        for name, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
            f2, fl = frame2.to(dt), fl32.to(dt)
            for _ in range(3):
                orig_warp(f2, fl)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(20):
                orig_warp(f2, fl)
            e1.record()
            torch.cuda.synchronize(dev)
            model_like[name] = e0.elapsed_time(e1) / 20
            del f2, fl
        del fl32
    except Exception as exc:   # an extra evidence line must never break the headline
        print(f"bench: model-like-flow warp timing skipped: {exc}", file=sys.stderr)
        model_like = {}

    # the fused warp + blend extension (no reference counterpart): two frames, two flows, one mask, bf16
    blend_ms = None
    try:
        fb2 = frame2.flip(0).contiguous()
        flb = (-flow).contiguous()
        mm = torch.rand(B, 1, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).to(dtype)
        for _ in range(3):
            vfi_b200.warp_blend(frame2, flow, fb2, flb, mm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(20):
            vfi_b200.warp_blend(frame2, flow, fb2, flb, mm)
        e1.record()
        torch.cuda.synchronize(dev)
        blend_ms = e0.elapsed_time(e1) / 20
        del fb2, flb, mm
    except Exception as exc:   # never let the extension break the headline line
        print(f"bench: warp_blend timing skipped: {exc}", file=sys.stderr)

    # ---------------------------------------------------------------- e2e: pinned host buffers through the public API
    e2e = None
    if not args.no_e2e and args.workload == "cfg5":
        e2e = e2e_stream(args, topo, dev, H, W)
    elif not args.no_e2e:
        from vfi_b200.hotpath import HostBatch

        src = synthetic_inputs(B, H, W, dtype=dtype, device="cpu", seed=99 + topo.rank, flow_sigma=flow_sigma, offset_sigma=off_sigma,
                               flow_kind=FLOW_KIND.get(args.workload, "smooth"))
        hb = HostBatch(B, H, W, dtype=dtype, conv27_channels_last=args.conv27_layout == "channels_last").copy_from(*src)
        del src
        h2d, d2h = hb.h2d_bytes, hb.d2h_bytes
        n_e2e = args.e2e_steps or min(args.steps, 5)
        path.run_host_batch(hb)                                 # warm-up (stream creation, allocator)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            path.run_host_batch(hb)                             # returns after the D2H copy has completed
        torch.cuda.synchronize(dev)
        dt_rank = time.perf_counter() - t0
        dt = dt_rank
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * n_e2e / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e,
               "h2d_gbs_this_rank": h2d * n_e2e / dt_rank / 1e9, "d2h_gbs_this_rank": d2h * n_e2e / dt_rank / 1e9,
               "copies_per_step": {"h2d": B, "d2h": B}, "cpu_affinity": affinity,
               "api": "vfi_b200.HotPath.run_host_batch (frame-major pinned HostBatch: one H2D and one D2H copy per frame, 3-stream pipeline)",
               "bound": "PCIe: the step moves feat (265 MB / frame) and three conv27 from the host -- tensors that are born on the device "
                        "in the real pipeline (see --workload cfg5 for the frame-in / frame-out number)"}
        del hb

    if not topo.is_root:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    tr = measured_traffic() if args.workload == "cfg2" else {}
    dcn_tflops = P * FLOP_PER_PX / (dcn_ms * 1e-3) / 1e12
    warp_gbs = P * WARP_BYTES_PER_PX_BF16 / (warp_ms * 1e-3) / 1e9
    line = {
        "metric": "1080p warp+DeformConv path frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "batch_per_gpu": B, "height": H, "width": W, "math": args.math,
                   "tensors": f"frame2/flow NCHW bf16, conv27 {args.conv27_layout} bf16, feat channels_last bf16, DCN weights [67,67,3,3] bf16",
                   "l2": "per-step working set (~5.6 GB) exceeds the 126 MB L2, no flush between iterations",
                   "sharding": "independent frame pairs per rank, no data-path collective"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "DCNv2 forward (one launch per layer, 3 per step)",
                     # the timed region is ~0.1-0.3 s: the BURST bf16 figure is the applicable peak (the sustained one belongs to
                     # kernels timed inside multi-second steps); both fractions are reported
                     "achieved": dcn_tflops, "peak": pk["tensor_burst"], "unit": "TFLOP/s",
                     "frac": dcn_tflops / pk["tensor_burst"], "traffic": tr.get("dcn_fwd"), "traffic_source": tr.get("dcn_fwd_source"), "peak_source": pk["source"] + " burst bf16",
                     "frac_of_burst_peak": dcn_tflops / pk["tensor_burst"], "frac_of_sustained_peak": dcn_tflops / pk["tensor_sustained"],
                     "ms_per_launch": dcn_ms,
                     "algorithmic_flop_per_launch": P * FLOP_PER_PX,
                     # what actually bounds the operator on this SM: building one A row reads
                     # 9 taps x 4 corners x 128 B of activations through the 128 B/clk/SM load/store data path
                     "lsu_bound": (lambda f_hz: {"bytes_per_px": 36 * 128, "floor_ms": 1e3 * P * 36 / 148 / f_hz,
                                                 "frac": (1e3 * P * 36 / 148 / f_hz) / dcn_ms,
                                                 "note": "36 x 128 B gather per output pixel / (128 B/clk/SM x 148 SMs) at the sampled SM clock",
                                                 "measured": (lambda m: None if m is None else {
                                                     "clk_per_px": m["clk_per_px"], "ms": 1e3 * P * m["clk_per_px"] / 148 / f_hz,
                                                     "frac": (1e3 * P * m["clk_per_px"] / 148 / f_hz) / dcn_ms, "producer_warps": m["producer_warps"],
                                                     "source": "scripts/microbench/gather_floor.cu, " + m["file"]})(measured_gather_floor())})(
                         1e6 * float((clocks or {}).get("sm_mhz") or 1965.0))},
        "roofline_warp": {"bound": "hbm", "kernel": "warp_fwd (planar NCHW bf16 in/out, timed alone, 20 launches)",
                          "achieved": P * WARP_BYTES_PER_PX_BF16 / (planar_ms["bf16"] * 1e-3) / 1e9, "peak": pk["hbm"],
                          "unit": "GB/s", "frac": P * WARP_BYTES_PER_PX_BF16 / (planar_ms["bf16"] * 1e-3) / 1e9 / pk["hbm"],
                          "traffic": tr.get("warp_fwd_planar_bf16"), "traffic_source": tr.get("warp_fwd_planar_bf16_source"), "ms_per_launch": planar_ms["bf16"],
                          "algorithmic_bytes_per_launch": P * WARP_BYTES_PER_PX_BF16, "peak_source": pk["source"],
                          "f32": {"ms_per_launch": planar_ms["f32"],
                                  "achieved": 2 * P * WARP_BYTES_PER_PX_BF16 / (planar_ms["f32"] * 1e-3) / 1e9,
                                  "frac": 2 * P * WARP_BYTES_PER_PX_BF16 / (planar_ms["f32"] * 1e-3) / 1e9 / pk["hbm"]},
                          "in_step": {"kernel": "warp_fwd_rec (writes the DCN tail records in place of torch.cat)",
                                      "ms_per_launch": warp_ms, "achieved": warp_gbs, "frac": warp_gbs / pk["hbm"],
                                      # the record form writes 16 bytes per pixel (3 channels + pad + mirrored half) where the planar
                                      # form writes 6: 26 bytes actually cross HBM per pixel against the 16 algorithmic ones above
                                      "dram_bytes_per_px": 26, "achieved_dram": warp_gbs * 26 / 16, "frac_dram": warp_gbs * 26 / 16 / pk["hbm"]}},
        "clocks": clocks,
    }
    if model_like.get("bf16") and model_like.get("f32"):
        wb = P * WARP_BYTES_PER_PX_BF16
        line["roofline_warp"]["model_like_flow"] = {
            "flow": "iid N(0, 0.03^2) px (SURVEY 8d class i: what a random-init model produces), 20 launches each",
            "bf16": {"ms_per_launch": model_like["bf16"], "achieved": wb / (model_like["bf16"] * 1e-3) / 1e9,
                     "frac": wb / (model_like["bf16"] * 1e-3) / 1e9 / pk["hbm"]},
            "f32": {"ms_per_launch": model_like["f32"], "achieved": 2 * wb / (model_like["f32"] * 1e-3) / 1e9,
                    "frac": 2 * wb / (model_like["f32"] * 1e-3) / 1e9 / pk["hbm"]}}
    if blend_ms:
        bl_bytes = P * (2 * 3 + 2 * 2 + 1 + 3) * 2
        line["roofline_blend"] = {"bound": "hbm", "kernel": "warp_blend (two planar bf16 frames + two flows + mask -> blended frame, 20 launches)",
                                  "achieved": bl_bytes / (blend_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": bl_bytes / (blend_ms * 1e-3) / 1e9 / pk["hbm"], "ms_per_launch": blend_ms,
                                  "algorithmic_bytes_per_launch": bl_bytes, "traffic": None, "peak_source": pk["source"]}
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        rows = 32
        ts = time_cpu(cpu_reference_step_factory(rows, W), reps=3, warm=1)
        fps = (rows * W / float(H * W)) / min(ts)
        line["cpu_baseline"] = {
            "value": fps, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{rows}x{W} stripe (batch 1, fp32) of the same path, min of 3 after 1 warm-up, scaled linearly in pixels (checked against one full 1080p frame on this pool: 50.1 s, the extrapolation is 4 % conservative -- profiles/r02_s_cpu_full_frame.json); "
                      "stock torch grid_sample + torchvision deform_conv2d CPU kernels via oracle/torch_ref.py"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
