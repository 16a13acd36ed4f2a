"""Pins oracle/vfi_oracle.c (and oracle/torch_ref.py) against outputs of the unmodified reference.

The fixtures in tests/golden were produced by tests/golden/make_golden.py, which runs
/root/reference/src/models/ema_vfi.py (EMA_VFI.warp, ModulatedDeformConvPack, EMA_VFI.forward) and
torchvision.ops.deform_conv2d on CPU.  Tolerance: the north star's fp32 bar, max-abs 1e-5 (scaled by the
magnitude of the tensor for gradients that reach |g| >> 1).
"""
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import load_golden

WARP_CASES = ["warp_rand", "warp_integer", "warp_tiny_flow", "warp_w1", "warp_h1", "warp_urban2"]
DCN_CASES = ["dcn_c67_sigma3", "dcn_c67_zero", "dcn_c5_o7_sigma1", "dcn_c67_sigma16"]


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if a.size else 0.0


def tol(ref, base=1e-5):
    return base * max(1.0, float(np.max(np.abs(ref)))) if ref.size else base


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_fwd_matches_reference(name):
    z = load_golden(name)
    out = oracle.warp_fwd(z["src"], z["flow"])
    assert maxabs(out, z["out"]) <= 1e-5


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_bwd_matches_reference(name):
    z = load_golden(name)
    if "grad_src" in z:
        gflow, gsrc = oracle.warp_bwd(z["grad_out"], z["src"], z["flow"], need_grad_src=True)
        assert maxabs(gsrc, z["grad_src"]) <= tol(z["grad_src"])
    else:
        gflow = oracle.warp_bwd(z["grad_out"], z["src"], z["flow"])
    assert maxabs(gflow, z["grad_flow"]) <= tol(z["grad_flow"])


@pytest.mark.parametrize("name", DCN_CASES)
def test_dcn_fwd_matches_reference(name):
    z = load_golden(name)
    out = oracle.dcn_fwd(z["x"], z["offset"], z["mask"], z["weight"], z["bias"])
    assert maxabs(out, z["out"]) <= 1e-5


@pytest.mark.parametrize("name", DCN_CASES)
def test_dcn_bwd_matches_reference(name):
    z = load_golden(name)
    gx, goff, gmask, gw, gb = oracle.dcn_bwd(z["grad_out"], z["x"], z["offset"], z["mask"], z["weight"])
    for got, key in ((gx, "grad_x"), (goff, "grad_offset"), (gmask, "grad_mask"), (gw, "grad_weight"), (gb, "grad_bias")):
        assert maxabs(got, z[key]) <= tol(z[key]), key


def test_pack_split_and_block_match_reference():
    z = load_golden("pack_c67")
    off, msk = oracle.pack_split(z["conv27"])
    assert maxabs(off, z["offset"]) == 0.0
    assert maxabs(msk, z["mask"]) <= 1e-6
    out = oracle.dcn_fwd(z["x"], off, msk, z["weight"], z["bias"])
    assert maxabs(out, z["out"]) <= 1e-5


def test_model_hot_path_matches_reference():
    """warp -> cat -> 3 x (split, DCN) replayed by the oracle on tensors recorded inside EMA_VFI.forward."""
    z = load_golden("model_24x32")
    warped = oracle.warp_fwd(z["frame2"], z["flow"])
    assert maxabs(warped, z["warped"]) <= 1e-5
    x = np.concatenate([z["feat"], warped], axis=1)
    for i in range(3):
        off, msk = oracle.pack_split(z[f"conv27_{i}"])
        x = oracle.dcn_fwd(x, off, msk, z[f"dcn_weight_{i}"], z[f"dcn_bias_{i}"])
        # each block is compared on the reference's own input to keep the tolerance per-op
        assert maxabs(x, z[f"block_out_{i}"]) <= 1e-5, i
        x = z[f"block_out_{i}"]


def test_warp_blend_is_composition_of_two_warps():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((2, 1, 3, 9, 13), dtype=np.float32)
    fa, fb = 2 * rng.standard_normal((2, 1, 2, 9, 13), dtype=np.float32)
    m = rng.random((1, 1, 9, 13), dtype=np.float32)
    out = oracle.warp_blend_fwd(a, fa, b, fb, m)
    ref = m * oracle.warp_fwd(a, fa) + (1 - m) * oracle.warp_fwd(b, fb)
    assert maxabs(out, ref) <= 1e-6


def test_round_trip_matters():
    """sampling at x+flow directly is NOT the reference; the oracle must replay the round trip."""
    z = load_golden("warp_tiny_flow")
    W = z["src"].shape[-1]
    x = np.arange(W, dtype=np.float32) + z["flow"][0, 0, 0]
    g = np.float32(2.0) * x / np.float32(W - 1) - np.float32(1.0)
    ix = ((g + np.float32(1.0)) / np.float32(2.0)) * np.float32(W - 1)
    assert np.any(ix != x)


# ---- torch_ref (stock torch / torchvision CPU kernels driven the way the reference drives them) ----------

def test_torch_ref_matches_golden():
    torch = pytest.importorskip("torch")
    pytest.importorskip("torchvision")
    from oracle import torch_ref

    for name in WARP_CASES:
        z = load_golden(name)
        out = torch_ref.warp(torch.from_numpy(z["src"]), torch.from_numpy(z["flow"])).numpy()
        assert maxabs(out, z["out"]) <= 1e-6, name
    z = load_golden("model_24x32")
    t = {k: torch.from_numpy(v) for k, v in z.items()}
    out = torch_ref.hot_path(t["frame2"], t["flow"], t["feat"], [t[f"conv27_{i}"] for i in range(3)],
                             [t[f"dcn_weight_{i}"] for i in range(3)], [t[f"dcn_bias_{i}"] for i in range(3)])
    assert maxabs(out.numpy(), z["block_out_2"]) <= 2e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_torch_ref_is_bit_exact_with_live_reference():
    torch = pytest.importorskip("torch")
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.ema_vfi import EMA_VFI
    finally:
        sys.path.remove("/root/reference")
    from oracle import torch_ref

    g = torch.Generator().manual_seed(5)
    src = torch.randn(2, 3, 31, 45, generator=g)
    flow = 4 * torch.randn(2, 2, 31, 45, generator=g)
    ref = EMA_VFI.warp(None, src, src, flow)
    assert torch.equal(torch_ref.warp(src, flow), ref)
    assert maxabs(oracle.warp_fwd(src.numpy(), flow.numpy()), ref.numpy()) <= 1e-5


@pytest.mark.parametrize("C,O,H,W,sigma", [(67, 67, 9, 13, 1.5), (5, 7, 12, 10, 6.0), (67, 67, 8, 8, 0.0)])
def test_column_gradient_split_matches_oracle(C, O, H, W, sigma):
    """The split vfi_dcn_bwd_data_cols is built on -- gcol = grad_out x W as a dense GEMM in the column order of
    ops.cols_weight_matrix, then the position-dependent half (oracle/cols_ref.py) -- reproduces the oracle's grad_x /
    grad_offset / grad_mask, including samples outside the image (sigma = 6) and integer positions (sigma = 0)."""
    import torch

    from oracle import cols_ref
    from vfi_b200.ops import cols_weight_matrix

    g = torch.Generator().manual_seed(7 + H)
    B = 2
    x = torch.randn(B, C, H, W, generator=g)
    off = sigma * torch.randn(B, 18, H, W, generator=g)
    m = torch.sigmoid(torch.randn(B, 9, H, W, generator=g))
    w = torch.randn(O, C, 3, 3, generator=g) / (9 * C) ** 0.5
    go = torch.randn(B, O, H, W, generator=g)
    M = cols_weight_matrix(w, torch.float64)                                       # [72, 9 * 72]
    assert M.shape == (72, 648)
    rows = torch.zeros(B * H * W, 72, dtype=torch.float64)
    rows[:, :O] = go.permute(0, 2, 3, 1).reshape(-1, O)
    gcol = (rows @ M).reshape(B, H, W, 9, 72)
    assert float(gcol[..., C:].abs().max()) == 0.0                                 # pad columns of every tap stay zero
    want = torch.einsum("bohw,ock->bhwkc", go.double(), w.double().reshape(O, C, 9))
    assert float((gcol[..., :C] - want).abs().max()) <= 1e-12
    gx, goff, gmask = cols_ref.dcn_bwd_data_from_cols(gcol[..., :C].numpy(), x.numpy(), off.numpy(), m.numpy())
    ref = oracle.dcn_bwd(go.numpy(), x.numpy(), off.numpy(), m.numpy(), w.numpy())
    for got, r, name in ((gx, ref[0], "grad_x"), (goff, ref[1], "grad_offset"), (gmask, ref[2], "grad_mask")):
        assert np.max(np.abs(got - r)) <= 2e-5 * max(1.0, float(np.max(np.abs(r)))), name


def test_fused_block_gradient_formula_matches_autograd_through_the_reference_glue():
    """The backward of the fused block (vfi_dcn_bwd_data_cols_fused; checked on the GPU against exactly this composition in
    smoke() and tests/test_gpu_parity.py): d loss / d conv27 = [grad_offset[:, :9] | grad_mask * m * (1 - m) | grad_offset[:, 9:]]
    with (grad_offset, grad_mask) from the oracle's DCN backward and m = sigmoid(conv27[:, 9:18]).  Pinned here against autograd
    through the reference's glue (chunk / cat / sigmoid, ema_vfi.py:57-59) around the stock torchvision op on CPU -- and through
    the unmodified ModulatedDeformConvPack itself where the reference tree is present."""
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")

    g = torch.Generator().manual_seed(21)
    B, C, H, W = 1, 67, 9, 11
    x = torch.randn(B, C, H, W, generator=g)
    c27 = torch.randn(B, 27, H, W, generator=g)
    c27[:, :9] *= 1.5
    c27[:, 18:] *= 1.5
    w = (torch.rand(C, C, 3, 3, generator=g) * 2 - 1) / 603 ** 0.5
    b = (torch.rand(C, generator=g) * 2 - 1) / 603 ** 0.5
    gy = torch.randn(B, C, H, W, generator=g)

    ct = c27.clone().requires_grad_(True)
    xt = x.clone().requires_grad_(True)
    o1, m, o2 = torch.chunk(ct, 3, dim=1)
    y = tv.ops.deform_conv2d(xt, torch.cat((o1, o2), dim=1), w, b, stride=1, padding=1, dilation=1, mask=torch.sigmoid(m))
    y.backward(gy)

    off, msk = oracle.pack_split(c27.numpy())
    gx, goff, gmask, _, _ = oracle.dcn_bwd(gy.numpy(), x.numpy(), off, msk, w.numpy())
    g27 = np.concatenate([goff[:, :9], gmask * msk * (1.0 - msk), goff[:, 9:]], axis=1)
    assert maxabs(g27, ct.grad.numpy()) <= tol(ct.grad.numpy())
    assert maxabs(gx, xt.grad.numpy()) <= tol(xt.grad.numpy())

    if os.path.isdir("/root/reference/src"):
        sys.dont_write_bytecode = True
        sys.path.insert(0, "/root/reference")
        try:
            from src.models.ema_vfi import ModulatedDeformConvPack
        finally:
            sys.path.remove("/root/reference")
        blk = ModulatedDeformConvPack(C, C)                      # ema_vfi.py:24: 3x3, stride 1, padding 1
        with torch.no_grad():
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
        xr = x.clone().requires_grad_(True)
        seen = {}

        def keep(mod, inp, out):                                 # returns None: the output is not replaced
            out.retain_grad()
            seen["c27"] = out

        h = blk.offset_conv.register_forward_hook(keep)
        try:
            blk(xr).backward(gy)
        finally:
            h.remove()
        cr = seen["c27"]
        off, msk = oracle.pack_split(cr.detach().numpy())
        _, goff, gmask, _, _ = oracle.dcn_bwd(gy.numpy(), x.numpy(), off, msk, blk.dcn_v2.weight.detach().numpy())
        g27 = np.concatenate([goff[:, :9], gmask * msk * (1.0 - msk), goff[:, 9:]], axis=1)
        assert maxabs(g27, cr.grad.numpy()) <= tol(cr.grad.numpy())
