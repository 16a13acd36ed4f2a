"""The hot path as one callable: warp -> concat -> 3 x (offset/mask split, DCNv2), i.e. lines 130-138 of
/root/reference/src/models/ema_vfi.py with the tensors that stock convolutions produce (feat, flow, the three
27-channel offset_conv outputs) taken as inputs.  This is the public API `bench.py` measures:

* :meth:`HotPath.run`       -- inputs already resident in HBM (device tensors in, device tensor out)
* :meth:`HotPath.run_host`  -- pinned HOST buffers in, pinned host buffer out; the call does the H2D copies, the
  kernels and the D2H copy, pipelined frame by frame over three CUDA streams so PCIe in, compute and PCIe out overlap.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import ops


def pack_split(conv27: torch.Tensor):
    """ema_vfi.py:57-59 (stock PyTorch glue): thirds (0, 2) -> offset, sigmoid(third 1) -> mask."""
    a, m, b = conv27.split(9, dim=1)
    return torch.cat((a, b), dim=1), torch.sigmoid(m)


class HotPath:
    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], *, math: str = "auto"):
        if len(weights) != len(biases):
            raise ValueError("one bias per DCN block")
        self.weights = list(weights)
        self.biases = list(biases)
        self.math = math
        self._streams = None
        self._bufs = {}

    # ---------------------------------------------------------------------------------------------- device API
    def run(self, frame2: torch.Tensor, flow: torch.Tensor, feat: torch.Tensor,
            convs27: Sequence[torch.Tensor], out: "ops.Planes | None" = None) -> torch.Tensor:
        """Tensor-core form (bf16 channels-last ``feat``): returns :class:`ops.Planes`.  The activation planes between the
        layers are cached per (shape, device, CUDA stream) and reused by the next call ON THAT STREAM; the RESULT is written
        to ``out`` when given, else to a plane set of the same cache -- in that case **it is overwritten by the next ``run``
        with the same key** (the benchmark's steady state: no allocation in the timed region).  Pass ``out=ops.Planes(...)``
        to keep results across calls.  Any other input takes the generic path and returns a fresh [B,67,H,W] tensor."""
        if self._fast(frame2, feat):
            # Tensor-core path with the reference's glue folded away: the warp writes its 3 channels into the 16-byte
            # tail record of each pixel, the first DCN layer gathers from (feat, tail) directly (no torch.cat), every
            # layer reads the raw 27-channel offset_conv output (no chunk/cat/sigmoid) and writes the two planes the
            # next layer gathers from.  Returns ops.Planes (``.to_nchw()`` gives the logical [B,67,H,W] tensor).
            B, C, H, W = feat.shape
            # keyed by stream as well: buffers allocated under one stream are only ever reused in that stream's order
            key = (B, H, W, feat.device, torch.cuda.current_stream(feat.device).cuda_stream)
            bufs = self._bufs.get(key) if isinstance(self._bufs, dict) else None
            if bufs is None:
                if not isinstance(self._bufs, dict) or len(self._bufs) >= 4:
                    self._bufs = {}
                # ping-pong activation planes, allocated once; tail pad channels of `src` are zeroed once and stay zero
                bufs = self._bufs[key] = (ops.Planes(B, H, W, feat.device, zero_tail=True), ops.Planes(B, H, W, feat.device),
                                          ops.Planes(B, H, W, feat.device))
            src, ping, pong = bufs
            ops.warp(frame2, flow, out=src.tail_nchw(frame2.shape[1]), tail_record=True)
            math = self.math if self.math != "auto" else "bf16_tc"
            x_main, x_tail = feat, src.tail_nchw(frame2.shape[1])
            dst = ping
            last = len(self.weights) - 1
            for i, (w, b, c27) in enumerate(zip(self.weights, self.biases, convs27)):
                y = ops.deform_conv2d_fused(x_main, x_tail, c27, w, b, math=math, out=out if (i == last and out is not None) else dst)
                x_main, x_tail = y.main_nchw, y.tail_nchw()
                dst = pong if dst is ping else ping
            return y
        warped = ops.warp(frame2, flow)
        x = torch.cat((feat, warped), dim=1)
        for w, b, c27 in zip(self.weights, self.biases, convs27):
            offset, mask = pack_split(c27)
            x = ops.deform_conv2d(x, offset, w, b, stride=1, padding=1, dilation=1, mask=mask, math=self.math)
        return x

    def _fast(self, frame2, feat) -> bool:
        return (self.math != "fp32" and feat.dtype == torch.bfloat16 and feat.shape[1] == 64 and frame2.shape[1] <= 8
                and feat.is_contiguous(memory_format=torch.channels_last))

    # ---------------------------------------------------------------------------------------------- host API
    def run_host(self, frame2: torch.Tensor, flow: torch.Tensor, feat: torch.Tensor, convs27: Sequence[torch.Tensor],
                 out: torch.Tensor, chunk: int = 1) -> torch.Tensor:
        """All arguments are pinned host tensors ([B,...]); ``out`` [B,O,H,W] receives the result.  Frames are
        processed in chunks of ``chunk`` so that copy-in of chunk i+1, compute of chunk i and copy-out of chunk i-1
        run concurrently.  Returns ``out`` after the last copy has completed (stream-synchronised)."""
        dev = self.weights[0].device
        for t in (frame2, flow, feat, out, *convs27):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("run_host takes pinned host tensors")
        if self._streams is None:
            self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        s_in, s_run, s_out = self._streams
        cur = torch.cuda.current_stream(dev)
        for s in self._streams:
            s.wait_stream(cur)
        B = frame2.shape[0]
        # device memory stays O(chunk): every per-chunk tensor is handed to the caching allocator as soon as Python drops it,
        # record_stream() keeps the memory from being reused before the streams that touch it have passed
        for i in range(0, B, chunk):
            sl = slice(i, min(i + chunk, B))
            with torch.cuda.stream(s_in):
                d = [t[sl].to(dev, non_blocking=True) for t in (frame2, flow, feat, *convs27)]
                ready = torch.cuda.Event()
                ready.record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ready)
                y = self.run(d[0], d[1], d[2], d[3:])
                if isinstance(y, ops.Planes):
                    y = y.to_nchw()
                for t in d:
                    t.record_stream(s_run)
                done = torch.cuda.Event()
                done.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                out[sl].copy_(y, non_blocking=True)
                y.record_stream(s_out)
        s_out.synchronize()
        return out


    def run_host_batch(self, hb: "HostBatch", chunk: int = 1) -> torch.Tensor:
        """``run_host`` on a frame-major :class:`HostBatch`: one H2D copy (the chunk's whole input record) and one D2H copy per
        chunk, pipelined over three streams.  Returns ``hb.out`` after the last copy has completed."""
        dev = self.weights[0].device
        if self._streams is None:
            self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        s_in, s_run, s_out = self._streams
        cur = torch.cuda.current_stream(dev)
        for s in self._streams:
            s.wait_stream(cur)
        for i in range(0, hb.B, chunk):
            sl = slice(i, min(i + chunk, hb.B))
            with torch.cuda.stream(s_in):
                d_arena = hb.arena[sl].to(dev, non_blocking=True)           # one cudaMemcpyAsync
                ready = torch.cuda.Event()
                ready.record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ready)
                f2, fl, ft, cv = hb.views(d_arena)
                ys = []
                for j in range(d_arena.shape[0]):                            # a frame's planes are dense: one frame per launch set
                    yj = self.run(f2[j:j + 1], fl[j:j + 1], ft[j:j + 1], [c[j:j + 1] for c in cv])
                    ys.append(yj.to_nchw() if isinstance(yj, ops.Planes) else yj)
                y = ys[0] if len(ys) == 1 else torch.cat(ys, 0)
                d_arena.record_stream(s_run)
                done = torch.cuda.Event()
                done.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                hb.out[sl].copy_(y, non_blocking=True)                       # one cudaMemcpyAsync
                y.record_stream(s_out)
        s_out.synchronize()
        return hb.out


class HostBatch:
    """Pinned host staging of one batch, FRAME-MAJOR: frame i's inputs -- frame2 [3,H,W], flow [2,H,W], feat [64,H,W]
    (channels-last inside the record), three conv27 [27,H,W] -- are one contiguous record, so a chunk of frames crosses PCIe
    as ONE cudaMemcpyAsync in each direction (``HotPath.run_host_batch``) instead of one per tensor.  The fields are ordinary
    strided views of the arena: fill them in place (``copy_from``) or hand them to whatever produces the tensors."""

    def __init__(self, B: int, H: int, W: int, *, dtype=torch.bfloat16, out_channels: int = 67, conv27_channels_last: bool = False):
        hw = H * W
        sizes = [3 * hw, 2 * hw, 64 * hw, 27 * hw, 27 * hw, 27 * hw]
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.B, self.H, self.W, self.dtype, self.conv_cl = B, H, W, dtype, conv27_channels_last
        self.arena = torch.empty((B, self.offsets[-1]), dtype=dtype).pin_memory()
        self.out = torch.empty((B, out_channels, H, W), dtype=dtype).pin_memory()
        self.frame2, self.flow, self.feat, self.convs = self.views(self.arena)

    def views(self, arena: torch.Tensor):
        """The six tensors of an arena laid out like this one (host arena or its device copy)."""
        o, n, H, W = self.offsets, arena.shape[0], self.H, self.W
        frame2 = arena[:, o[0]:o[1]].view(n, 3, H, W)
        flow = arena[:, o[1]:o[2]].view(n, 2, H, W)
        feat = arena[:, o[2]:o[3]].view(n, H, W, 64).permute(0, 3, 1, 2)
        convs = [arena[:, o[3 + i]:o[4 + i]].view(n, H, W, 27).permute(0, 3, 1, 2) if self.conv_cl else arena[:, o[3 + i]:o[4 + i]].view(n, 27, H, W)
                 for i in range(3)]
        return frame2, flow, feat, convs

    def copy_from(self, frame2, flow, feat, convs) -> "HostBatch":
        self.frame2.copy_(frame2); self.flow.copy_(flow); self.feat.copy_(feat)
        for d, c in zip(self.convs, convs):
            d.copy_(c)
        return self

    @property
    def h2d_bytes(self) -> int:
        return self.arena.numel() * self.arena.element_size()

    @property
    def d2h_bytes(self) -> int:
        return self.out.numel() * self.out.element_size()


def synthetic_inputs(B: int, H: int, W: int, *, dtype=torch.bfloat16, device="cuda", seed: int = 1234,
                     flow_sigma: float = 8.0, offset_sigma: float = 1.5, pinned_host: bool = False,
                     flow_kind: str = "smooth"):
    """Synthetic tensors of the shapes the path sees inside EMA_VFI.forward: ImageNet-normalised
    uniform frames, a flow field (``smooth``: N(0,1) at 1/32 resolution, bilinearly up-sampled, times ``flow_sigma`` px --
    the locality real optical flow has; ``iid``: per-pixel N(0, sigma^2), the worst-case gather), N(0,1) features, and 27-channel offset_conv outputs whose offset thirds are
    N(0, sigma^2) px and whose mask third is N(0,1) (-> sigmoid)."""
    g = torch.Generator(device="cpu" if pinned_host else device).manual_seed(seed)
    where = "cpu" if pinned_host else device

    def fin(t):
        t = t.to(dtype)
        return t.pin_memory() if pinned_host else t

    mean = torch.tensor([0.485, 0.456, 0.406], device=where).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=where).view(1, 3, 1, 1)
    frame2 = fin((torch.rand(B, 3, H, W, generator=g, device=where) - mean) / std)
    if flow_kind == "smooth":
        coarse = torch.randn(B, 2, max(H // 32, 2), max(W // 32, 2), generator=g, device=where)
        flow = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True) * flow_sigma
    elif flow_kind == "iid":
        flow = flow_sigma * torch.randn(B, 2, H, W, generator=g, device=where)
    else:
        raise ValueError(f"unknown flow_kind {flow_kind!r}")
    flow = fin(flow)
    feat = fin(torch.randn(B, 64, H, W, generator=g, device=where))
    convs = []
    for _ in range(3):
        c = torch.randn(B, 27, H, W, generator=g, device=where)
        c[:, :9] *= offset_sigma
        c[:, 18:] *= offset_sigma
        convs.append(fin(c))
    return frame2, flow, feat, convs


def synthetic_weights(C: int = 67, *, dtype=torch.bfloat16, device="cuda", seed: int = 4321, blocks: int = 3):
    """torchvision's DeformConv2d default init: kaiming-uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for weight and bias."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    bound = 1.0 / (C * 9) ** 0.5
    ws = [((torch.rand(C, C, 3, 3, generator=g) * 2 - 1) * bound).to(device=device, dtype=dtype) for _ in range(blocks)]
    bs = [((torch.rand(C, generator=g) * 2 - 1) * bound).to(device=device, dtype=dtype) for _ in range(blocks)]
    return ws, bs
