"""vfi_b200.refmodel.StockInterpolator -- the stock-PyTorch stand-in that hosts the hot path on machines without the
reference checkout -- is pinned against the unmodified reference (when /root/reference is present) and against the golden
fixture recorded from it (everywhere).  CPU only: the stand-in is stock torch / torchvision, none of the library's kernels."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from vfi_b200.refmodel import StockInterpolator

REF = os.environ.get("VFI_REFERENCE", "/root/reference")


def digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().float().numpy().tobytes())
    return h.hexdigest()


def psnr_model(z, device="cpu"):
    """The network of the model_psnr_256 fixture: seeded default init + the recorded offset_conv weights + scaled motion conv."""
    torch.manual_seed(int(z["seed"]))
    m = StockInterpolator().eval()
    assert digest(m.state_dict()) == str(z["state_digest"]), "seeded construction differs from the reference's on this machine"
    with torch.no_grad():
        for i, blk in enumerate(m.attention_blocks):
            blk.offset_conv.weight.copy_(torch.from_numpy(z[f"offset_conv_weight_{i}"]))
            blk.offset_conv.bias.copy_(torch.from_numpy(z[f"offset_conv_bias_{i}"]))
        m.motion_estimation[-1].weight.mul_(float(z["motion_scale"]))
    return m.to(device)


def psnr_inputs(z):
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    to_t = lambda a: ((torch.from_numpy(a.transpose(2, 0, 1).copy()).float() / 255.0)[None] - mean) / std  # noqa: E731
    gt = torch.from_numpy(z["frame_gt"].transpose(2, 0, 1).copy()).float()[None] / 255.0
    return to_t(z["frame_a"]), to_t(z["frame_b"]), gt


def psnr(a, b):
    return float(-10.0 * torch.log10(((a.double() - b.double()) ** 2).mean()))


def test_seeded_construction_matches_the_recorded_digest():
    z = load_golden("model_psnr_256")
    psnr_model(z)          # asserts the digest


def test_standin_reproduces_the_recorded_reference_output_on_a_crop():
    """Full 256 x 256 forward of the stand-in on CPU takes ~4 s; a 64 x 64 corner of the recorded output is not comparable
    (global average pooling), so the whole frame is run once."""
    z = load_golden("model_psnr_256")
    m = psnr_model(z)
    a, b, gt = psnr_inputs(z)
    with torch.no_grad():
        out = m(a, b)
    assert float((out - torch.from_numpy(z["model_out"])).abs().max()) <= 1e-6
    assert abs(psnr(out, gt) - psnr(torch.from_numpy(z["model_out"]), gt)) < 1e-6


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "models")), reason="reference checkout not present")
def test_standin_is_the_reference_key_for_key_and_bit_for_bit():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from src.models.ema_vfi import EMA_VFI
    finally:
        sys.path.remove(REF)
    torch.manual_seed(11)
    ref = EMA_VFI().eval()
    torch.manual_seed(11)
    mine = StockInterpolator().eval()
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr) == list(sm)
    assert all(torch.equal(sr[k], sm[k]) for k in sr)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for blk in ref.attention_blocks:
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
    mine.load_state_dict(ref.state_dict())
    a, b = torch.rand(2, 3, 40, 56), torch.rand(2, 3, 40, 56)
    with torch.no_grad():
        assert torch.equal(ref(a, b), mine(a, b))
    flow = 3.0 * torch.randn(2, 2, 40, 56)
    assert torch.equal(ref.warp(a, a, flow), mine.warp(a, a, flow))
