"""The hot path inside the network (BASELINE configs 2 and 5): the structure-identical stand-in of the
reference model (vfi_b200.refmodel, stock PyTorch layers; the reference checkout does not travel to the GPU box) at 1080p in
the reference's GPU inference mode (no_grad + CUDA autocast, inference.py:158-159), timed

  stock          no drop-in: stock aten grid_sample + torchvision deform_conv2d CUDA kernels
  dropin         vfi_b200.install(): both seams on libvfi_b200 (tensor cores under autocast)
  fused          vfi_b200.install(fuse=True): + seams 3 / 4 (one [B,H,W,72] buffer between the blocks, no cat / chunk / sigmoid)

and through vfi_b200.stream.PairStreamer (uint8 frames in pinned memory -> uint8 frames out, the loop of inference.py:160-199).
usage: model_bench.py [--batch 1 4] [--stream-frames 65] [--out FILE]     (one JSON line per measurement)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vfi_b200
from vfi_b200 import shard, stream
from vfi_b200.refmodel import StockInterpolator


def build_model(dev):
    torch.manual_seed(2026)
    m = StockInterpolator().eval()
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for blk in m.attention_blocks:          # F4: offset_conv is zero-initialised; randomise it so the gather is exercised
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
        m.motion_estimation[-1].weight.mul_(40.0)
    return m.to(dev)


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 4])
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--stream-frames", type=int, default=65)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    topo = shard.init_distributed()
    dev = torch.device("cuda", topo.local_rank)
    torch.cuda.set_device(dev)
    model = build_model(dev)
    H, W = args.height, args.width
    lines = []

    def emit(d):
        d.update(n_gpus=topo.world, height=H, width=W)
        if topo.is_root:
            print(json.dumps(d), flush=True)
        lines.append(d)

    modes = {"stock": None, "dropin": dict(fuse=False), "fused": dict(fuse=True)}
    ref_out = {}
    for B in args.batch:
        g = torch.Generator(device=dev).manual_seed(5)
        mean = torch.tensor(stream.MEAN, device=dev).view(1, 3, 1, 1)
        std = torch.tensor(stream.STD, device=dev).view(1, 3, 1, 1)
        a = (torch.rand(B, 3, H, W, device=dev, generator=g) - mean) / std
        b = (torch.rand(B, 3, H, W, device=dev, generator=g) - mean) / std
        for name, kw in modes.items():
            if kw is not None:
                vfi_b200.install(StockInterpolator, **kw)
            try:
                def step():
                    with torch.no_grad(), torch.autocast("cuda"):
                        return model(a, b)
                try:
                    out = step().float()
                    ms = timed(step)
                except torch.OutOfMemoryError as exc:       # stock torchvision materialises 603 x P columns
                    emit({"what": "model_forward", "mode": name, "batch": B, "error": "out of memory", "detail": str(exc)[:120]})
                    continue
                rec = {"what": "model_forward", "mode": name, "batch": B, "ms": ms, "frames_per_s": B / (ms * 1e-3),
                       "autocast": "cuda fp16 (inference.py:158-159)", "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
                if name == "stock":
                    ref_out[B] = out
                elif B in ref_out:
                    rec["max_abs_vs_stock"] = float((out - ref_out[B]).abs().max())
                    rec["psnr_vs_stock_db"] = float(-10 * torch.log10(((out - ref_out[B]) ** 2).mean()))
                if kw is not None:
                    rec["calls"] = vfi_b200.dropin.call_counts()
                emit(rec)
                del out
            finally:
                if kw is not None:
                    vfi_b200.uninstall()
            torch.cuda.reset_peak_memory_stats(dev)
        ref_out.pop(B, None)
        del a, b
        torch.cuda.empty_cache()

    # ---- the frame loop: uint8 1080p frames through PairStreamer, each rank its contiguous share of the pairs
    n = args.stream_frames
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    frames = [np.roll(base, 3 * i, axis=1) for i in range(n)]
    for name, kw in (("stock", None), ("fused", dict(fuse=True))):
        if kw is not None:
            vfi_b200.install(StockInterpolator, **kw)
        try:
            for bp in (1, 8):
                ps = stream.PairStreamer(model, dev, batch_pairs=bp, topology=topo, autocast_dtype=torch.float16)
                try:
                    list(ps.run(frames[:9]))                     # warm-up
                    torch.cuda.synchronize()
                    if topo.world > 1:
                        import torch.distributed as dist
                        dist.barrier()
                    t0 = time.perf_counter()
                    written = sum(1 for _ in ps.run(frames))
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                except torch.OutOfMemoryError:
                    emit({"what": "pair_streamer", "mode": name, "batch_pairs": bp, "error": "out of memory"})
                    continue
                if topo.world > 1:
                    import torch.distributed as dist
                    t = torch.tensor([dt], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                emit({"what": "pair_streamer", "mode": name, "batch_pairs": bp, "frames_in": n, "pairs": n - 1, "written_this_rank": written,
                      "seconds": dt, "pairs_per_s": (n - 1) / dt, "h2d_bytes": ps.stats["h2d_bytes"], "d2h_bytes": ps.stats["d2h_bytes"],
                      "note": "uint8 frames from pinned host memory in, uint8 frames out; model under no_grad + fp16 autocast"})
        finally:
            if kw is not None:
                vfi_b200.uninstall()
    if args.out and topo.is_root:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
