"""Structure-identical stand-in of the reference's interpolation network, built from stock ``torch.nn`` layers.

The reference checkout (``/root/reference``) does not travel to the GPU box, but measuring the hot path *inside* the network
(BASELINE config 5, ``inference.py``'s frame loop, the PSNR gate of the north star) needs the network around it.  This module
provides it without touching or copying the reference's sources: the same layers with the same attribute names, so that a
reference ``state_dict`` loads key for key (``attention_blocks.N.dcn_v2.{weight,bias}`` etc.) and a seeded construction
consumes the random stream in the same order, and the same two seams the drop-in patches:

* ``StockInterpolator.warp(self, frame2, feature, flow)``      -- /root/reference/src/models/ema_vfi.py:149
* ``FusionPack.forward(self, x)`` -> ``self.dcn_v2(x, offset, mask)``  -- ema_vfi.py:53-60 (``ModulatedDeformConvPack``)

Everything here is stock PyTorch / torchvision (the convolution trunk "stays in stock PyTorch", north star); the hot path is
whatever ``vfi_b200.install`` routes the two seams to.  ``tests/test_oracle.py`` checks on CPU, where the reference is
present, that this module reproduces ``EMA_VFI.forward`` bit for bit from the reference's own ``state_dict``.

Layer map (ema_vfi.py line of the layer it mirrors):
  feat_ext_conv1 6->64 (:73), feat_ext_blocks 3 x 64->64 (:74-76), context_encoding 64->128/s2 ->256/s2 ->256, GAP, Linear 256->64
  (:79-86), motion_estimation 128->64->64->2 (:89-93), attention_blocks 3 x FusionPack(67) (:96-99), reconstruction 67->64->32->3
  + tanh (:102-107); forward order :110-147.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


def _cr(cin: int, cout: int, stride: int = 1) -> nn.Sequential:
    """3x3 convolution + ReLU as a two-element Sequential (parameter keys ``<name>.0.weight`` / ``<name>.0.bias``)."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1), nn.ReLU())


class FusionPack(nn.Module):
    """offset_conv (zero-initialised, 27 channels) -> thirds (0, 2) = offsets, sigmoid(third 1) = mask -> DCNv2 (ema_vfi.py:22-60)."""

    def __init__(self, channels: int):
        super().__init__()
        from torchvision.ops import DeformConv2d

        self.offset_conv = nn.Conv2d(channels, 27, 3, 1, 1)
        nn.init.zeros_(self.offset_conv.weight)
        nn.init.zeros_(self.offset_conv.bias)
        self.dcn_v2 = DeformConv2d(channels, channels, kernel_size=3, stride=1, padding=1, dilation=1, bias=True)

    def forward(self, x):
        first, logits, last = torch.chunk(self.offset_conv(x), 3, dim=1)
        return self.dcn_v2(x, torch.cat((first, last), dim=1), torch.sigmoid(logits))


class StockInterpolator(nn.Module):
    def __init__(self, in_channels: int = 3, mid_channels: int = 64, num_blocks: int = 3):
        super().__init__()
        c = mid_channels
        # construction order = the reference's, so torch.manual_seed(s) gives the same initial weights
        self.feat_ext_conv1 = _cr(2 * in_channels, c)
        self.feat_ext_blocks = nn.Sequential(OrderedDict((f"conv_block_{i}", _cr(c, c)) for i in range(num_blocks)))
        self.context_encoding = nn.Sequential(_cr(c, 2 * c, 2), _cr(2 * c, 4 * c, 2), _cr(4 * c, 4 * c), nn.AdaptiveAvgPool2d(1),
                                              nn.Flatten(), nn.Linear(4 * c, c))
        self.motion_estimation = nn.Sequential(_cr(2 * c, c), _cr(c, c), nn.Conv2d(c, 2, 3, 1, 1))
        self.attention_blocks = nn.ModuleList(FusionPack(c + in_channels) for _ in range(num_blocks))
        self.reconstruction = nn.Sequential(_cr(c + in_channels, c), _cr(c, c // 2), nn.Conv2d(c // 2, in_channels, 3, 1, 1), nn.Tanh())

    # ---- the trunk in three pieces, so that tests and the streamer can run the hot path between them
    def features_and_flow(self, frame1, frame2):
        feat = self.feat_ext_blocks(self.feat_ext_conv1(torch.cat((frame1, frame2), dim=1)))
        context = self.context_encoding(feat)
        spread = context[:, :, None, None].repeat(1, 1, feat.size(2), feat.size(3))
        return feat, self.motion_estimation(torch.cat((feat, spread), dim=1))

    def fuse(self, frame2, feat, flow):
        x = torch.cat((feat, self.warp(frame2, feat, flow)), dim=1)
        for block in self.attention_blocks:
            x = block(x)
        return x

    def reconstruct(self, fused):
        return (self.reconstruction(fused) + 1) / 2

    def forward(self, frame1, frame2):
        feat, flow = self.features_and_flow(frame1, frame2)
        return self.reconstruct(self.fuse(frame2, feat, flow))

    def warp(self, frame2, feature, flow):
        """Seam 1 (same signature as the reference's helper): backward warp of ``frame2`` by ``flow`` in pixels."""
        B, _, H, W = frame2.shape
        xs = torch.arange(W).view(1, 1, 1, W).expand(B, 1, H, W)
        ys = torch.arange(H).view(1, 1, H, 1).expand(B, 1, H, W)
        grid = torch.cat((xs, ys), dim=1).float()
        if feature.is_cuda:
            grid = grid.cuda()
        v = grid + flow
        gx = 2.0 * v[:, 0] / max(W - 1, 1) - 1.0
        gy = 2.0 * v[:, 1] / max(H - 1, 1) - 1.0
        return F.grid_sample(frame2, torch.stack((gx, gy), dim=-1), align_corners=True)
