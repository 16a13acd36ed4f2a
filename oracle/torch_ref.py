"""Stock-PyTorch restatement of the reference call sequence.  TEST INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box, but the third-party kernels its hot path runs do
(torch 2.11 ``aten::grid_sampler_2d`` and torchvision 0.26 ``torchvision::deform_conv2d``).  This module drives
those kernels with the same arguments the reference passes, so that

* ``bench.py --impl reference`` / ``cpu_baseline`` can time the reference's CPU path on the box's host cores, and
* GPU tests can cross-check sizes the C oracle would take minutes for.

It is validated against the unmodified reference in ``tests/test_oracle.py`` (bit-exact on CPU when
``/root/reference`` is present) and against ``tests/golden/*.npz`` everywhere.

Reference lines followed: ``src/models/ema_vfi.py:149-171`` (warp), ``:53-60`` (pack), ``:136-138`` (3 blocks).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def warp(frame2: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """Backward warp of ``frame2`` by ``flow`` (pixels; channel 0 = x, 1 = y), as ``EMA_VFI.warp`` computes it.

    Pixel grid + flow (ema_vfi.py:153-162), scale each axis with ``2*v/max(n-1,1) - 1`` (:165-166), channels-last
    view (:168), bilinear ``grid_sample`` with zero padding and ``align_corners=True`` (:169).
    """
    B, _, H, W = frame2.shape
    cols = torch.arange(W, device=flow.device, dtype=torch.float32).expand(H, W)
    rows = torch.arange(H, device=flow.device, dtype=torch.float32).unsqueeze(1).expand(H, W)
    vx = cols.unsqueeze(0) + flow[:, 0]
    vy = rows.unsqueeze(0) + flow[:, 1]
    gx = 2.0 * vx / max(W - 1, 1) - 1.0
    gy = 2.0 * vy / max(H - 1, 1) - 1.0
    grid = torch.stack((gx, gy), dim=-1)
    return F.grid_sample(frame2, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


class WarpHost:
    """Minimal stand-in for the reference model class: exposes the seam ``warp(self, frame2, feature, flow)``
    (``src/models/ema_vfi.py:149``) so drop-in tests can patch a class on machines without ``/root/reference``."""

    def warp(self, frame2, feature, flow):  # noqa: D401 - same signature as the reference seam
        return warp(frame2, flow)


def pack_split(conv27: torch.Tensor):
    """ema_vfi.py:57-59: thirds (0, 2) of the 27-channel tensor -> 18-channel offset, sigmoid(third 1) -> mask."""
    a, m, b = conv27.split(9, dim=1)
    return torch.cat((a, b), dim=1), torch.sigmoid(m)


def dcn(x, offset, mask, weight, bias):
    """The call ``DeformConv2d.forward`` makes for the reference geometry (3x3, stride 1, pad 1, dilation 1)."""
    import torchvision.ops.deform_conv as tv  # resolved at call time so an installed drop-in is honoured

    return tv.deform_conv2d(x, offset, weight, bias, stride=(1, 1), padding=(1, 1), dilation=(1, 1), mask=mask)


def dcn_stock(x, offset, mask, weight, bias):
    """Same as :func:`dcn` but always the stock torchvision kernel, even when the drop-in is installed."""
    import torchvision  # noqa: F401  (registers torch.ops.torchvision)

    return torch.ops.torchvision.deform_conv2d(x, weight, offset, mask, bias, 1, 1, 1, 1, 1, 1, 1, 1, True)


class FusionBlock(torch.nn.Module):
    """Test stand-in with the structure of the reference's ModulatedDeformConvPack (ema_vfi.py:23-60): a
    3x3 conv producing 27 channels, the split above, and a torchvision ``DeformConv2d`` called ``dcn_v2``."""

    def __init__(self, channels: int = 67):
        super().__init__()
        from torchvision.ops import DeformConv2d

        self.offset_conv = torch.nn.Conv2d(channels, 27, 3, 1, 1)
        self.dcn_v2 = DeformConv2d(channels, channels, kernel_size=3, stride=1, padding=1, dilation=1, bias=True)

    def forward(self, x):
        offset, mask = pack_split(self.offset_conv(x))
        return self.dcn_v2(x, offset, mask)


def hot_path(frame2, flow, feat, convs27, weights, biases):
    """warp + cat + 3 x DCN exactly as ema_vfi.py:130-138 orders them, with the 27-channel offset_conv outputs
    supplied as inputs (they come from stock convolutions that are outside the path)."""
    x = torch.cat((feat, warp(frame2, flow)), dim=1)
    for c27, w, b in zip(convs27, weights, biases):
        off, m = pack_split(c27)
        x = dcn_stock(x, off, m, w, b)
    return x
