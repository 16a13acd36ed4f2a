#!/usr/bin/env python
"""BASELINE config 3: training step of the hot path at crop 256x256, global batch 16 -- warp and 3 x DCNv2 forward + backward
(fp32, the 1e-5 parity kernels) with one flat-bucket NCCL gradient all-reduce when launched under torchrun.

    python scripts/train_step_bench.py [--steps K] [--stock]        # --stock: the same step on stock torch / torchvision CUDA ops

Prints one JSON line (samples/s over all ranks, CUDA events, max over ranks).  Not the headline metric (bench.py is)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vfi_b200
from vfi_b200 import shard


from vfi_b200.trainstep import FusionBlockParams as Block, stock_warp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stock", action="store_true")
    ap.add_argument("--math", default="fp32", choices=["fp32", "bf16_tc"],
                    help="fp32: parity kernels; bf16_tc: bf16 tensors, tcgen05 forward + weight gradient (bf16 tolerances)")
    args = ap.parse_args()
    topo = shard.init_distributed()
    dev = torch.device("cuda", topo.local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(7)
    B, H, W = 16 // topo.world, 256, 256
    if args.stock:
        import torchvision

        dcn = torchvision.ops.deform_conv2d

        def warp(f, fl):   # ema_vfi.py:149-171 on the device
            ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
            g = torch.stack((xs, ys), 0).float()[None] + fl
            g = torch.stack((2 * g[:, 0] / (W - 1) - 1, 2 * g[:, 1] / (H - 1) - 1), -1)
            return torch.nn.functional.grid_sample(f, g, mode="bilinear", padding_mode="zeros", align_corners=True)
    else:
        import functools

        dcn, warp = functools.partial(vfi_b200.deform_conv2d, math=args.math), vfi_b200.warp
    dt = torch.bfloat16 if (args.math == "bf16_tc" and not args.stock) else torch.float32
    blocks = torch.nn.ModuleList([Block(dcn) for _ in range(3)]).to(dev)
    bucket = shard.GradBucket(blocks.parameters())
    g = torch.Generator(device=dev).manual_seed(100 + topo.rank)
    frame2 = torch.randn(B, 3, H, W, device=dev, generator=g).to(dt)
    feat = torch.randn(B, 64, H, W, device=dev, generator=g).to(dt)
    flow = (2.0 * torch.randn(B, 2, H, W, device=dev, generator=g)).to(dt).requires_grad_(True)

    def step():
        bucket.zero()
        flow.grad = None
        x = torch.cat((feat, warp(frame2, flow)), 1)
        for blk in blocks:
            x = blk(x)
        x.float().square().mean().backward()
        bucket.attach()
        bucket.allreduce_mean()

    for _ in range(args.warmup):
        step()
    if topo.world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    if topo.world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if topo.is_root:
        print(json.dumps({"workload": "cfg3: training step, crop 256x256, global batch 16, fp32: warp + 3 x (offset_conv, DCNv2) "
                          "forward + backward, flat-bucket gradient all-reduce", "impl": "stock torch/torchvision CUDA" if args.stock
                          else f"vfi_b200 ({args.math})", "n_gpus": topo.world, "ms_per_step": ms,
                          "samples_per_s": 16 / (ms * 1e-3), "grad_flow_finite": bool(torch.isfinite(flow.grad).all())}))
    if topo.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
