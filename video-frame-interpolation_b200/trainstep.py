"""BASELINE config 3: one training step of the hot path -- warp, concat and 3 x (offset_conv, DCNv2) forward + backward at crop
256 x 256, global batch 16 -- as a reusable object for ``bench.py --workload cfg3`` and ``scripts/train_step_bench.py``.

Reference lines: the forward is /root/reference/src/models/ema_vfi.py:130-138 with the 64-channel feature map, the frame and
the flow as inputs; backward is what ``loss.backward()`` (train.py:125) runs through those lines; the gradient exchange is the
one collective BASELINE config 3 adds between train.py:125 and :128 (the reference itself is single-process).

Data parallel over ``world`` ranks: every rank holds the same parameters and 16 / world samples; the parameter gradients are
all-reduced group by group (block 3 first -- its gradients are final first) from autograd hooks, so the exchange of blocks 3
and 2 hides under the backward of blocks 2 and 1 and only block 1's segment is exposed (``shard.GradBucket``).
"""
from __future__ import annotations

import functools
from typing import Optional

import torch

from . import ops, shard


class FusionBlockParams(torch.nn.Module):
    """offset_conv + modulated DCN as in ema_vfi.py:41-60 (stock conv for the offsets, the op under test for the DCN), with
    fp32 master parameters: in the bf16 mode they are cast per step and autograd returns fp32 gradients to them."""

    def __init__(self, dcn):
        super().__init__()
        self.offset_conv = torch.nn.Conv2d(67, 27, 3, padding=1)
        self.weight = torch.nn.Parameter(torch.empty(67, 67, 3, 3).uniform_(-1, 1) / 603 ** 0.5)
        self.bias = torch.nn.Parameter(torch.empty(67).uniform_(-1, 1) / 603 ** 0.5)
        torch.nn.init.normal_(self.offset_conv.weight, std=0.02)
        torch.nn.init.normal_(self.offset_conv.bias, std=0.5)
        self.dcn = dcn

    def forward(self, x):
        dt = x.dtype
        c27 = torch.nn.functional.conv2d(x, self.offset_conv.weight.to(dt), self.offset_conv.bias.to(dt), padding=1)
        o1, m, o2 = c27.chunk(3, dim=1)
        return self.dcn(x, torch.cat((o1, o2), 1), self.weight.to(dt), self.bias.to(dt), stride=1, padding=1, dilation=1,
                        mask=torch.sigmoid(m))

    def forward_records(self, x72):
        """The same block on 72-channel record activations (ops.records_buffer) with the glue folded into the kernels, forward
        and backward (ops.deform_conv2d_block): the stock offset_conv runs on the record buffer as it lies (zero weights for
        the pad / mirror channels 67..71, so cuDNN neither pads nor converts layouts), its raw 27-channel output goes straight
        to the DCN kernels, and the fp32 master parameters of the DCN are passed as they are."""
        w = self.offset_conv.weight
        w72 = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, x72.shape[1] - w.shape[1])).to(torch.bfloat16)
        w72 = w72.contiguous(memory_format=torch.channels_last)
        c27 = torch.nn.functional.conv2d(x72, w72, self.offset_conv.bias.to(torch.bfloat16), padding=1)
        return ops.deform_conv2d_block(x72, c27, self.weight, self.bias)


def stock_warp(frame, flow):
    """ema_vfi.py:149-171 with stock ops on the device (the --stock arm)."""
    _, _, H, W = frame.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=frame.device), torch.arange(W, device=frame.device), indexing="ij")
    g = torch.stack((xs, ys), 0).float()[None] + flow
    g = torch.stack((2 * g[:, 0] / (W - 1) - 1, 2 * g[:, 1] / (H - 1) - 1), -1)
    return torch.nn.functional.grid_sample(frame, g.to(frame.dtype), mode="bilinear", padding_mode="zeros", align_corners=True)


class TrainStep:
    def __init__(self, topo: shard.Topology, device, *, math: str = "bf16_tc", stock: bool = False, global_batch: int = 16,
                 size: int = 256, overlap: bool = True, seed: int = 7, fused: bool = False):
        self.topo, self.dev = topo, torch.device(device)
        if global_batch % topo.world:
            raise ValueError(f"global batch {global_batch} does not divide over {topo.world} ranks")
        self.B, self.H, self.W = global_batch // topo.world, size, size
        self.global_batch = global_batch
        torch.manual_seed(seed)                              # identical replicas
        if stock:
            import torchvision

            dcn, self.warp = torchvision.ops.deform_conv2d, stock_warp
        else:
            dcn, self.warp = functools.partial(ops.deform_conv2d, math=math), ops.warp
        self.dtype = torch.bfloat16 if (math == "bf16_tc" and not stock) else torch.float32
        # fused: record activations + ops.deform_conv2d_block (tensor-core math only)
        self.fused = fused and not stock and math == "bf16_tc"
        self.blocks = torch.nn.ModuleList([FusionBlockParams(dcn) for _ in range(3)]).to(self.dev)
        groups = [list(b.parameters()) for b in reversed(self.blocks)]      # backward order: block 3's gradients are final first
        self.bucket = shard.GradBucket(self.blocks.parameters(), groups=groups)
        self.overlap = overlap and topo.world > 1
        self.bucket.zero()
        self.bucket.attach()
        if self.overlap:
            self.bucket.enable_overlap()
        g = torch.Generator(device=self.dev).manual_seed(100 + topo.rank)
        B, H, W, dt = self.B, self.H, self.W, self.dtype
        self.frame2 = torch.randn(B, 3, H, W, device=self.dev, generator=g).to(dt)
        self.feat = torch.randn(B, 64, H, W, device=self.dev, generator=g).to(dt)
        self.flow = (2.0 * torch.randn(B, 2, H, W, device=self.dev, generator=g)).to(dt).requires_grad_(True)
        if self.fused:
            self.feat = self.feat.contiguous(memory_format=torch.channels_last)
            self._zero1 = torch.zeros(B, 1, H, W, device=self.dev, dtype=dt)
        self.exposed_ms = []
        self.graph = None
        self.graph_launches = 0

    # ---------------------------------------------------------------------------------------------- CUDA graph
    def capture(self, warmup: int = 3) -> None:
        """Capture one whole step -- zero the bucket, forward, backward with its hook-launched group all-reduces, the final wait --
        into a CUDA graph; ``step()`` then replays it.  At 16 / world samples per GPU the step is ~200 launches of a few
        microseconds each (half of them the stock convolutions): launch latency, not the kernels, bounds it from N = 4 on.  All
        tensors of a step are static (inputs, the flat gradient bucket) or live in the graph's private pool (activations,
        workspaces, TMA tensor maps are encoded from those fixed addresses at capture time)."""
        if self.graph is not None:
            return
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):                           # eager warm-up on the side stream the capture will use
            for _ in range(warmup):
                self._body(measure=False)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g, stream=s):
            self._body(measure=False)
        self.graph_launches = ops.launch_count() - n0        # kernels of this library inside one replay
        self.graph = g

    def step(self, measure: bool = False) -> Optional[float]:
        if self.graph is not None and not measure:
            self.graph.replay()
            return None
        return self._body(measure)

    def _body(self, measure: bool = False) -> Optional[float]:
        self.bucket.zero()
        self.flow.grad = None
        if self.overlap:
            self.bucket.begin_step()
        if self.fused:
            # [feat | warped | 0 | warped | 0] = the record layout (channel 67 zero, 68..71 mirror 64..67); the mirror's gradient
            # is zero by construction (deform_conv2d_block returns zeros for channels 67..71)
            w = self.warp(self.frame2, self.flow)
            z = self._zero1
            x = torch.cat((self.feat, w, z, w, z), 1).contiguous(memory_format=torch.channels_last)
            for blk in self.blocks:
                x = blk.forward_records(x)
            x = x[:, :67]
        else:
            x = torch.cat((self.feat, self.warp(self.frame2, self.flow)), 1)
            for blk in self.blocks:
                x = blk(x)
        # mean over the GLOBAL batch: every rank contributes sum over its shard / (global count), summed by the all-reduce
        (x.float().square().sum() / (x.numel() * self.topo.world)).backward()
        if self.overlap:
            exposed = self.bucket.finish_overlap(measure=measure, average=False)   # the loss above is already the global mean
            if measure:
                self.exposed_ms.append(exposed)
            return exposed
        self.bucket.attach()
        if self.topo.world > 1:
            import torch.distributed as dist

            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM)
            e1.record()
            if measure:
                e1.synchronize()
                self.exposed_ms.append(float(e0.elapsed_time(e1)))
        return None
