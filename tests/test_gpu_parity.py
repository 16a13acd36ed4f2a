"""GPU parity tests (run on the B200 box with `-m gpu`).  Every op is called through the C ABI (ctypes -> libvfi_b200.so)
and compared with

* the committed golden vectors produced by the unmodified reference (tests/golden/*.npz),
* the C oracle (oracle/vfi_oracle.c) on seeded random inputs at sizes it finishes in seconds,
* the stock torch / torchvision kernels (oracle/torch_ref.py) at larger sizes, and
* size-independent properties at BASELINE.json's full sizes (1080p batch 8, 4K).

Tolerances: fp32 max-abs 1e-5 (north star), scaled by max|ref| for tensors whose magnitude exceeds 1 (gradients);
bf16: max|delta| / max|ref| <= 1e-2 against fp32 arithmetic on bf16-rounded inputs.
"""
import numpy as np
import pytest
import torch

import oracle
import vfi_b200
from conftest import load_golden
from oracle import torch_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
WARP_CASES = ["warp_rand", "warp_integer", "warp_tiny_flow", "warp_w1", "warp_h1", "warp_urban2"]
DCN_CASES = ["dcn_c67_sigma3", "dcn_c67_zero", "dcn_c5_o7_sigma1", "dcn_c67_sigma16"]


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def maxabs(a, b):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


def tol(ref, base=1e-5):
    ref = ref.detach().float().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    return base * max(1.0, float(np.max(np.abs(ref)))) if ref.size else base


def relerr(a, ref):
    ref = ref.detach().float().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    return maxabs(a, ref) / max(float(np.max(np.abs(ref))), 1e-30)


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def test_library_runs_on_this_device():
    assert torch.cuda.get_device_capability(0)[0] == 10
    assert vfi_b200._lib.load().vfi_check_device() == 0


# ------------------------------------------------------------------------------------------------------------ warp
@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_fwd_golden(name):
    z = load_golden(name)
    out = vfi_b200.warp(cu(z["src"]), cu(z["flow"]))
    assert maxabs(out, z["out"]) <= 1e-5


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_bwd_golden(name):
    z = load_golden(name)
    src = cu(z["src"]).requires_grad_("grad_src" in z)
    flow = cu(z["flow"]).requires_grad_(True)
    vfi_b200.warp(src, flow).backward(cu(z["grad_out"]))
    assert maxabs(flow.grad, z["grad_flow"]) <= tol(z["grad_flow"])
    if "grad_src" in z:
        assert maxabs(src.grad, z["grad_src"]) <= tol(z["grad_src"])


@pytest.mark.parametrize("shape,sigma", [((2, 3, 64, 96), 0.03), ((1, 3, 270, 480), 8.0), ((3, 3, 33, 61), 20.0),
                                         ((1, 5, 47, 52), 3.0)])
def test_warp_fwd_bwd_vs_oracle(shape, sigma):
    g = torch.Generator().manual_seed(11)
    B, C, H, W = shape
    src = torch.randn(shape, generator=g)
    flow = sigma * torch.randn(B, 2, H, W, generator=g)
    go = torch.randn(shape, generator=g)
    f = flow.to(DEV).requires_grad_(True)
    out = vfi_b200.warp(src.to(DEV), f)
    out.backward(go.to(DEV))
    assert maxabs(out, oracle.warp_fwd(src.numpy(), flow.numpy())) <= 1e-5
    ref = oracle.warp_bwd(go.numpy(), src.numpy(), flow.numpy())
    assert maxabs(f.grad, ref) <= tol(ref)


def test_warp_strided_and_channels_last_inputs():
    g = torch.Generator().manual_seed(12)
    src = torch.randn(2, 3, 40, 64, generator=g)
    flow = 4 * torch.randn(2, 2, 40, 64, generator=g)
    ref = oracle.warp_fwd(src.numpy(), flow.numpy())
    # channels_last source (NHWC in memory) and a flow that is a slice of a wider tensor
    s_cl = src.to(DEV).contiguous(memory_format=torch.channels_last)
    wide = torch.zeros(2, 5, 40, 64, device=DEV)
    wide[:, 1:3] = flow.to(DEV)
    out = vfi_b200.warp(s_cl, wide[:, 1:3])
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert maxabs(out, ref) <= 1e-5
    # odd width -> scalar (non-vectorised) kernel
    out2 = vfi_b200.warp(src[..., :63].contiguous().to(DEV), flow[..., :63].contiguous().to(DEV))
    assert maxabs(out2, oracle.warp_fwd(src[..., :63].numpy(), flow[..., :63].numpy())) <= 1e-5


def test_warp_into_tail_plane_records():
    """bf16 warp written straight into the 16-byte tail records the tensor-core DCN gathers from (no torch.cat)."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(15)
    src = torch.randn(2, 3, 40, 64, generator=g).to(torch.bfloat16)
    flow = (5 * torch.randn(2, 2, 40, 64, generator=g)).to(torch.bfloat16)
    ref = oracle.warp_fwd(src.float().numpy(), flow.float().numpy())
    pl = ops.Planes(2, 40, 64, DEV)
    pl.tail.fill_(7.0)                                            # garbage: the kernel must overwrite pad channels with 0
    out = vfi_b200.warp(src.to(DEV), flow.to(DEV), out=pl.tail_nchw(3), tail_record=True)
    assert out.data_ptr() == pl.tail.data_ptr()
    assert relerr(pl.tail_nchw(3), ref) <= 1e-2
    assert float(pl.tail[..., 3].abs().max()) == 0.0 and torch.equal(pl.tail[..., 4:], pl.tail[..., :4])   # mirrored halves
    assert maxabs(pl.tail_nchw(3), vfi_b200.warp(src.to(DEV), flow.to(DEV))) == 0.0   # same values as the planar kernel


def test_warp_into_a_slice_of_an_8_channel_tensor_keeps_the_other_channels():
    """Without VFI_WARP_OUT_TAIL_RECORD the warp writes exactly the channels `out` describes, whatever its strides: a
    [:, :3] slice of an ordinary channels-last bf16 tensor has the strides of a tail plane but is not one."""
    g = torch.Generator().manual_seed(16)
    src = torch.randn(1, 3, 24, 40, generator=g).to(torch.bfloat16).to(DEV)
    flow = (3 * torch.randn(1, 2, 24, 40, generator=g)).to(torch.bfloat16).to(DEV)
    buf = torch.full((1, 8, 24, 40), 7.0, dtype=torch.bfloat16, device=DEV).contiguous(memory_format=torch.channels_last)
    out = vfi_b200.warp(src, flow, out=buf[:, :3])
    assert out.data_ptr() == buf.data_ptr()
    assert maxabs(buf[:, :3], vfi_b200.warp(src, flow)) == 0.0
    assert float((buf[:, 3:] - 7.0).abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="TAIL_RECORD"):
        vfi_b200.warp(src, flow, out=torch.empty_like(src), tail_record=True)


def test_warp_empty_batch_and_errors():
    out = vfi_b200.warp(torch.zeros(0, 3, 8, 8, device=DEV), torch.zeros(0, 2, 8, 8, device=DEV))
    assert out.shape == (0, 3, 8, 8)
    with pytest.raises(RuntimeError, match="flow must be"):
        vfi_b200.warp(torch.zeros(1, 3, 8, 8, device=DEV), torch.zeros(1, 3, 8, 8, device=DEV))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_warp_low_precision(dtype):
    g = torch.Generator().manual_seed(13)
    src = torch.randn(2, 3, 72, 128, generator=g).to(dtype)
    flow = (6 * torch.randn(2, 2, 72, 128, generator=g)).to(dtype)
    ref = oracle.warp_fwd(src.float().numpy(), flow.float().numpy())
    out = vfi_b200.warp(src.to(DEV), flow.to(DEV))
    assert out.dtype == dtype
    assert relerr(out, ref) <= 1e-2
    out32 = vfi_b200.warp(src.to(DEV), flow.float().to(DEV))      # fp32 flow with low-precision frames
    assert relerr(out32, ref) <= 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("division", ["ieee", "reciprocal"])
def test_warp_staged_kernel_is_bit_identical_to_the_l1_kernel(dtype, division):
    """The TMA-staged warp (source window in shared memory, zero fill by the copy engine) against round 1's L1-gather
    kernel on every flow class the staging decision distinguishes: near-identity (48 x 40 window), sheared / translated
    (64-wide window, anchored at the tile's north-west corner wherever it lies, also outside the frame), incoherent large
    displacement (tiles fall back to L1 inside the same launch), NaN / inf flow, frames whose size is not a multiple of the
    tile.  Results must not depend on the route: exact equality, and the tile counters prove both routes ran."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(77)
    for (B, H, W), kind in [((2, 72, 136), "tiny"), ((1, 100, 232), "shear"), ((2, 64, 96), "translate"), ((1, 96, 160), "iid"),
                            ((1, 40, 72), "nan"), ((3, 33, 40), "edge")]:
        src = torch.randn(B, 3, H, W, generator=g).to(dtype).to(DEV)
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        if kind == "tiny":
            flow = 0.03 * torch.randn(B, 2, H, W, generator=g)
        elif kind == "shear":
            flow = torch.stack([0.3 * yy - 0.2 * xx + 5, 0.25 * xx - 9], 0)[None].repeat(B, 1, 1, 1) + 0.2 * torch.randn(B, 2, H, W, generator=g)
        elif kind == "translate":
            flow = torch.tensor([-70.3, 41.6]).view(1, 2, 1, 1) + 0.5 * torch.randn(B, 2, H, W, generator=g)   # window partly / wholly outside
        elif kind == "iid":
            flow = 40.0 * torch.randn(B, 2, H, W, generator=g)
        elif kind == "nan":
            flow = 2.0 * torch.randn(B, 2, H, W, generator=g)
            flow[0, 0, 3, 5] = float("nan"); flow[0, 1, 20, 40] = float("inf"); flow[0, 0, 39, 71] = -float("inf")
        else:
            flow = 1.5 * torch.randn(B, 2, H, W, generator=g)
        flow = flow.to(DEV)
        ops.warp_tile_counts(reset=True)
        a = vfi_b200.warp(src, flow, division=division, count_tiles=True)
        staged, direct = ops.warp_tile_counts(reset=True)
        b = vfi_b200.warp(src, flow, division=division, staging=False)
        tiles = B * ((H + 31) // 32) * ((W + 31) // 32)
        assert staged + direct == tiles, (kind, staged, direct, tiles)
        if kind in ("tiny", "translate", "edge") or (kind == "shear" and dtype != torch.float32):
            assert direct == 0, (kind, staged, direct)
        if kind == "shear":                                        # fp32 frames have a 64 x 32 big window: some tiles do not fit
            assert staged > 0, (kind, staged, direct)
        if kind == "iid":
            assert direct == tiles
        assert torch.equal(torch.nan_to_num(a.float(), nan=123.0), torch.nan_to_num(b.float(), nan=123.0)), kind
    # tail-record output and a source the copy engine cannot map (row pitch not a multiple of 16 bytes): L1 kernels, same values
    src = torch.randn(2, 3, 48, 64, generator=g).to(torch.bfloat16).to(DEV)
    flow = (0.5 * torch.randn(2, 2, 48, 64, generator=g)).to(DEV)
    pa, pb = ops.Planes(2, 48, 64, DEV), ops.Planes(2, 48, 64, DEV)
    vfi_b200.warp(src, flow, out=pa.tail_nchw(3), tail_record=True, division=division)
    vfi_b200.warp(src, flow, out=pb.tail_nchw(3), tail_record=True, division=division, staging=False)
    assert torch.equal(pa.tail, pb.tail)
    wide = torch.randn(1, 3, 40, 75, generator=g).to(dtype).to(DEV)
    sl, fl = wide[..., 1:73], (2.0 * torch.randn(1, 2, 40, 72, generator=g)).to(DEV)
    ops.warp_tile_counts(reset=True)
    a = vfi_b200.warp(sl, fl, division=division, count_tiles=True)
    assert ops.warp_tile_counts() == (0, 0)                       # not staged at all: the whole launch is the L1 kernel
    assert torch.equal(a, vfi_b200.warp(sl.contiguous(), fl, division=division))


def test_warp_blend_matches_composition():
    g = torch.Generator().manual_seed(14)
    a, b = torch.randn(2, 2, 3, 50, 70, generator=g)
    fa, fb = 3 * torch.randn(2, 2, 2, 50, 70, generator=g)
    m = torch.rand(2, 1, 50, 70, generator=g)
    out = vfi_b200.warp_blend(a.to(DEV), fa.to(DEV), b.to(DEV), fb.to(DEV), m.to(DEV))
    ref = oracle.warp_blend_fwd(a.numpy(), fa.numpy(), b.numpy(), fb.numpy(), m.numpy())
    assert maxabs(out, ref) <= 1e-5


def test_warp_full_size_properties_1080p_batch8_and_4k():
    """BASELINE configs 2 and 4 at full size.  Properties that need no oracle: the op is exactly linear in the frame
    (scaling by 2 commutes bit for bit), a constant frame stays constant wherever all four corners are inside, and
    every batch entry agrees with the stock aten::grid_sampler_2d kernel driven the way the reference drives it.
    (Zero flow is NOT the identity: the reference's normalise/un-normalise round trip moves samples by ~1e-4 px.)"""
    g = torch.Generator(device=DEV).manual_seed(1234)
    for shape, sigma in [((8, 3, 1080, 1920), 8.0), ((1, 3, 2160, 3840), 64.0)]:
        B, C, H, W = shape
        src = torch.randn(shape, device=DEV, generator=g)
        flow = sigma * torch.randn(B, 2, H, W, device=DEV, generator=g)       # incoherent displacement field
        a = vfi_b200.warp(src, flow)
        assert torch.equal(vfi_b200.warp(2.0 * src, flow), 2.0 * a)
        # the reference as it runs on CPU (true division): one batch entry through the stock CPU kernel
        ref = torch_ref.warp(src[:1].cpu(), flow[:1].cpu())
        assert maxabs(a[:1], ref) <= 1e-5
        # the reference as it runs on CUDA (aten multiplies by the reciprocal): every batch entry, stock CUDA kernel
        r = vfi_b200.warp(src, flow, division="reciprocal")
        for b in range(B):
            ref = torch_ref.warp(src[b:b + 1], flow[b:b + 1])
            assert maxabs(r[b:b + 1], ref) <= 1e-5
        ones = vfi_b200.warp(torch.ones(1, 1, H, W, device=DEV), 0.4 * torch.ones(1, 2, H, W, device=DEV))
        assert float((ones[:, :, : H - 1, : W - 1] - 1.0).abs().max()) <= 1e-6
        del src, flow, a, r, ref, ones
        torch.cuda.empty_cache()


def test_warp_backward_config3_full_size_vs_stock_cuda():
    """BASELINE config 3 at full size (16 x 3 x 256 x 256, forward + backward): grad_flow and grad_frame against autograd
    through the stock aten::grid_sampler_2d kernels driven the way the reference drives them (oracle/torch_ref.warp on
    CUDA).  division="reciprocal" replays the CUDA path's coordinate arithmetic, so the bar is the fp32 one; the flow
    gradient is a difference of neighbouring pixels times grad_out, scaled by max|ref|."""
    g = torch.Generator(device=DEV).manual_seed(77)
    src = torch.randn(16, 3, 256, 256, device=DEV, generator=g)
    flow = 6.0 * torch.randn(16, 2, 256, 256, device=DEV, generator=g)
    go = torch.randn(16, 3, 256, 256, device=DEV, generator=g)
    s1, f1 = src.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    out = vfi_b200.warp(s1, f1, division="reciprocal")
    out.backward(go)
    s2, f2 = src.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    ref = torch_ref.warp(s2, f2)
    ref.backward(go)
    assert maxabs(out, ref) <= 1e-5
    assert maxabs(f1.grad, f2.grad) <= tol(f2.grad, 2e-5)
    assert maxabs(s1.grad, s2.grad) <= tol(s2.grad, 2e-5)
    # the model path differentiates with respect to the flow only: that takes the fast kernel (clamped 2 x 2 patch,
    # re-slotted corners) -- bit-identical to the generic kernel's grad_flow, edges included (|flow| up to ~25 px here)
    f3 = flow.clone().requires_grad_(True)
    vfi_b200.warp(src, f3, division="reciprocal").backward(go)
    assert torch.equal(f3.grad, f1.grad)


# ------------------------------------------------------------------------------------------------------------ DCN
def run_dcn(z, dtype=torch.float32, math="auto", grads=True):
    t = {k: cu(z[k], dtype if k in ("x", "offset", "mask", "weight", "bias", "grad_out") else None) for k in z}
    leaves = [t[k].requires_grad_(grads) for k in ("x", "offset", "mask", "weight", "bias")]
    out = vfi_b200.deform_conv2d(t["x"], t["offset"], t["weight"], t["bias"], stride=1, padding=1, dilation=1,
                                 mask=t["mask"], math=math)
    if grads:
        out.backward(t["grad_out"])
    return out, leaves


@pytest.mark.parametrize("name", DCN_CASES)
def test_dcn_fwd_bwd_golden(name):
    z = load_golden(name)
    out, (x, off, m, w, b) = run_dcn(z)
    assert maxabs(out, z["out"]) <= 1e-5
    for t, key in ((x, "grad_x"), (off, "grad_offset"), (m, "grad_mask"), (w, "grad_weight"), (b, "grad_bias")):
        assert maxabs(t.grad, z[key]) <= tol(z[key]), key


def rand_dcn(B, C, O, H, W, sigma, seed=0):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (C * 9) ** 0.5
    return dict(x=torch.randn(B, C, H, W, generator=g).numpy(),
                offset=(sigma * torch.randn(B, 18, H, W, generator=g)).numpy(),
                mask=torch.sigmoid(torch.randn(B, 9, H, W, generator=g)).numpy(),
                weight=((torch.rand(O, C, 3, 3, generator=g) * 2 - 1) * bound).numpy(),
                bias=((torch.rand(O, generator=g) * 2 - 1) * bound).numpy(),
                grad_out=torch.randn(B, O, H, W, generator=g).numpy())


@pytest.mark.parametrize("B,H,W,sigma", [(1, 37, 53, 1.5), (2, 16, 130, 8.0), (1, 64, 64, 0.0)])
def test_dcn_fwd_bwd_vs_oracle(B, H, W, sigma):
    z = rand_dcn(B, 67, 67, H, W, sigma, seed=B * 100 + H)
    out, (x, off, m, w, b) = run_dcn(z)
    assert maxabs(out, oracle.dcn_fwd(z["x"], z["offset"], z["mask"], z["weight"], z["bias"])) <= 1e-5
    ref = oracle.dcn_bwd(z["grad_out"], z["x"], z["offset"], z["mask"], z["weight"])
    for t, r, key in zip((x, off, m, w, b), ref, ("grad_x", "grad_offset", "grad_mask", "grad_weight", "grad_bias")):
        assert maxabs(t.grad, r) <= tol(r, 2e-5), key


def test_dcn_config1_size_vs_stock_torchvision_cuda():
    """BASELINE config 1 geometry (256x256, batch 1, fp32) against torchvision's own CUDA kernel."""
    z = rand_dcn(1, 67, 67, 256, 256, 1.5, seed=5)
    out, (x, off, m, w, b) = run_dcn(z)
    t = {k: cu(v).requires_grad_(k != "grad_out") for k, v in z.items()}
    ref = torch_ref.dcn_stock(t["x"], t["offset"], t["mask"], t["weight"], t["bias"])
    ref.backward(t["grad_out"])
    assert maxabs(out, ref) <= 1e-5
    for mine, key in ((x, "x"), (off, "offset"), (m, "mask"), (w, "weight"), (b, "bias")):
        assert maxabs(mine.grad, t[key].grad) <= tol(t[key].grad, 3e-5), key


def test_dcn_strided_inputs_and_mixed_dtypes():
    z = rand_dcn(2, 67, 67, 24, 40, 2.0, seed=7)
    ref = oracle.dcn_fwd(z["x"], z["offset"], z["mask"], z["weight"], z["bias"])
    x_cl = cu(z["x"]).contiguous(memory_format=torch.channels_last)
    out = vfi_b200.deform_conv2d(x_cl, cu(z["offset"]), cu(z["weight"]), cu(z["bias"]), stride=1, padding=1, dilation=1,
                                 mask=cu(z["mask"]))
    assert maxabs(out, ref) <= 1e-5
    # what CUDA autocast hands over: fp32 activations, half-precision offsets and mask
    off16, m16 = cu(z["offset"], torch.float16), cu(z["mask"], torch.float16)
    ref16 = oracle.dcn_fwd(z["x"], off16.float().cpu().numpy(), m16.float().cpu().numpy(), z["weight"], z["bias"])
    out16 = vfi_b200.deform_conv2d(cu(z["x"]), off16, cu(z["weight"]), cu(z["bias"]), stride=1, padding=1, dilation=1, mask=m16)
    assert out16.dtype == torch.float32 and maxabs(out16, ref16) <= 1e-5


def test_dcn_zero_offset_unit_mask_is_a_plain_convolution_1080p():
    """Full 1080p frame: with offsets 0 and mask 1 DCNv2 degenerates to conv2d (checked against cuDNN)."""
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(1, 67, 1080, 1920, device=DEV, generator=g)
    w = (torch.rand(67, 67, 3, 3, device=DEV, generator=g) * 2 - 1) / 603 ** 0.5
    b = torch.randn(67, device=DEV, generator=g) * 0.01
    off = torch.zeros(1, 18, 1080, 1920, device=DEV)
    m = torch.ones(1, 9, 1080, 1920, device=DEV)
    out = vfi_b200.deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m, math="fp32")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x, w, b, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert maxabs(out, ref) <= 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_dcn_low_precision_tensors_fp32_math(dtype):
    """Low-precision tensors through the fp32-arithmetic kernels (forward and backward): reference = fp32 arithmetic
    on the rounded inputs (there is no native bf16 torchvision kernel)."""
    z = rand_dcn(2, 67, 67, 40, 72, 1.5, seed=9)
    zr = {k: torch.from_numpy(v).to(dtype).float().numpy() for k, v in z.items()}
    ref = oracle.dcn_fwd(zr["x"], zr["offset"], zr["mask"], zr["weight"], zr["bias"])
    out, leaves = run_dcn(z, dtype=dtype, math="fp32", grads=True)
    gref = oracle.dcn_bwd(zr["grad_out"], zr["x"], zr["offset"], zr["mask"], zr["weight"])
    for t, r in zip(leaves, gref):
        assert relerr(t.grad, r) <= 2e-2
    assert out.dtype == dtype
    assert relerr(out, ref) <= 1e-2


# ------------------------------------------------------------------------------------------------------------ tcgen05
def test_umma_selftest_matches_matmul():
    """The tensor-core plumbing alone (descriptors, SWIZZLE_128B layout, tcgen05.mma/commit/ld, TMEM) as a plain GEMM."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(21)
    for K in (64, 192, 704):
        a = torch.randn(128, K, generator=g).to(torch.bfloat16).to(DEV)
        b = torch.randn(80, K, generator=g).to(torch.bfloat16).to(DEV)
        d = ops.selftest_umma(a, b)
        ref = a.float() @ b.float().t()
        assert maxabs(d, ref) <= 1e-3 * max(1.0, float(ref.abs().max())), K


@pytest.mark.parametrize("B,H,W,sigma", [(1, 8, 16, 0.0), (1, 8, 16, 1.5), (2, 37, 53, 1.5), (2, 20, 40, 1.5), (1, 64, 96, 8.0)])
def test_dcn_tensor_core_path(B, H, W, sigma):
    """bf16 operands / fp32 accumulate through tcgen05: reference = fp32 arithmetic on the bf16-rounded inputs."""
    z = rand_dcn(B, 67, 67, H, W, sigma, seed=31 + H)
    zr = {k: bf16_round(v) for k, v in z.items()}
    ref = oracle.dcn_fwd(zr["x"], zr["offset"], zr["mask"], zr["weight"], zr["bias"])
    t = {k: cu(z[k], torch.bfloat16) for k in ("x", "offset", "mask", "weight", "bias")}
    out = vfi_b200.deform_conv2d(t["x"], t["offset"], t["weight"], t["bias"], stride=1, padding=1, dilation=1,
                                 mask=t["mask"], math="bf16_tc")
    assert out.shape == (B, 67, H, W) and out.dtype == torch.bfloat16
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert relerr(out, ref) <= 1e-2
    # chained with fp32 offsets/mask
    out2 = vfi_b200.deform_conv2d(out, cu(z["offset"]), t["weight"], t["bias"], stride=1, padding=1, dilation=1,
                                  mask=cu(z["mask"]), math="bf16_tc")
    ref2 = oracle.dcn_fwd(out.float().cpu().numpy(), z["offset"], z["mask"], zr["weight"], zr["bias"])
    assert relerr(out2, ref2) <= 1e-2
    # fp32 tensors routed through the tensor-core path (inputs rounded to bf16 inside), NCHW fp32 result
    out3 = vfi_b200.deform_conv2d(cu(z["x"]), cu(z["offset"]), cu(z["weight"]), cu(z["bias"]), stride=1, padding=1,
                                  dilation=1, mask=cu(z["mask"]), math="bf16_tc")
    assert out3.dtype == torch.float32 and out3.is_contiguous()
    ref3 = oracle.dcn_fwd(zr["x"], z["offset"], z["mask"], zr["weight"], z["bias"])
    assert relerr(out3, ref3) <= 1e-2


def test_dcn_tensor_core_path_fp16_offsets_under_autocast_dtypes():
    """What the unmodified model hands over under fp16 autocast: fp32 input, fp16 offset / mask.  With
    math="bf16_tc" this takes the v6 kernel (fp16 geometry inputs, input packed to planes inside the call)."""
    z = rand_dcn(2, 67, 67, 24, 48, 1.5, seed=77)
    off16 = torch.from_numpy(z["offset"]).to(torch.float16)
    m16 = torch.from_numpy(z["mask"]).to(torch.float16)
    out = vfi_b200.deform_conv2d(cu(z["x"]), off16.to(DEV), cu(z["weight"]), cu(z["bias"]), stride=1, padding=1, dilation=1,
                                 mask=m16.to(DEV), math="bf16_tc")
    assert out.dtype == torch.float32 and out.shape == (2, 67, 24, 48)
    ref = oracle.dcn_fwd(bf16_round(z["x"]), off16.float().numpy(), m16.float().numpy(), bf16_round(z["weight"]), z["bias"])
    assert relerr(out, ref) <= 1e-2


@pytest.mark.parametrize("math,bar", [("bf16_tc", 1e-2)])
def test_dcn_fused_split_input_and_conv27(math, bar):
    """vfi_dcn_fwd_fused: (feat 64ch channels-last, 3-channel tail) + raw 27-channel offset_conv output, against the
    oracle chain cat -> split/sigmoid -> DCN on the same bf16 tensors."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(51)
    B, H, W = 2, 24, 40
    feat = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16)
    tail3 = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16)
    c27 = torch.randn(B, 27, H, W, generator=g)
    c27[:, :9] *= 1.5
    c27[:, 18:] *= 1.5
    c27 = c27.to(torch.bfloat16)
    w = ((torch.rand(67, 67, 3, 3, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16)
    b = ((torch.rand(67, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16)
    src = ops.Planes(B, H, W, DEV, zero_tail=True)
    src.main.copy_(feat.permute(0, 2, 3, 1))
    src.set_tail(tail3)
    out = ops.deform_conv2d_fused(src.main_nchw, src.tail_nchw(3), c27.to(DEV), w.to(DEV), b.to(DEV), math=math)
    off, m = oracle.pack_split(c27.float().numpy())
    ref = oracle.dcn_fwd(torch.cat([feat, tail3], 1).float().numpy(), off, bf16_round(m), w.float().numpy(), b.float().numpy())
    assert relerr(out.to_nchw(), ref) <= bar
    # channel 3 of the tail plane is zero padding, the upper half of every record mirrors the lower one
    assert float(out.tail[..., 3].abs().max()) == 0.0 and torch.equal(out.tail[..., 4:], out.tail[..., :4])
    # planes in -> planes out (what layers 2 and 3 of the path do), and a plain NCHW tensor in
    out2 = ops.deform_conv2d_fused(out.main_nchw, out.tail_nchw(), c27.to(DEV), w.to(DEV), b.to(DEV), math=math)
    ref2 = oracle.dcn_fwd(out.to_nchw().float().cpu().numpy(), off, bf16_round(m), w.float().numpy(), b.float().numpy())
    assert relerr(out2.to_nchw(), ref2) <= bar
    out3 = ops.deform_conv2d_fused(torch.cat([feat, tail3], 1).to(DEV), None, c27.to(DEV), w.to(DEV), b.to(DEV), math=math)
    assert relerr(out3.to_nchw(), ref) <= bar
    # a channels_last offset_conv output (what a channels_last model produces) is read as it lies: same bits out
    c27_cl = c27.to(DEV).contiguous(memory_format=torch.channels_last)
    out4 = ops.deform_conv2d_fused(src.main_nchw, src.tail_nchw(3), c27_cl, w.to(DEV), b.to(DEV), math=math)
    assert torch.equal(out4.main, out.main) and torch.equal(out4.tail, out.tail)


def test_umma_ts_selftest_matches_matmul():
    """The plumbing v6 adds: A written to tensor memory by tcgen05.st (16x256b with the producers' thread mapping, and
    32x32b) and consumed by the A-from-TMEM MMA form.  The raw TMEM image must be A itself (row r = lane r, 32-bit column c
    = K elements 2c, 2c+1)."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(22)
    a = torch.randn(128, 64, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(80, 64, generator=g).to(torch.bfloat16).to(DEV)
    d, raw = ops.selftest_umma_ts(a, b)
    want_raw = a.contiguous().view(torch.int32).reshape(128, 32)
    assert torch.equal(raw, want_raw), "tcgen05.st.16x256b thread <-> (lane, column) mapping differs from the one assumed"
    ref = 2.0 * (a.float() @ b.float().t())
    assert maxabs(d, ref) <= 1e-3 * max(1.0, float(ref.abs().max()))


def test_dcn_kernel_variants_agree():
    """The default forward (v7: TMA tensor maps, tail channels gathered by the geometry warps) against VFI_DCN_KERNEL=v6 (the
    round-1 kernel, kept as one instantiation for exactly this check) in fresh processes -- the variant is read once per
    process.  Same K order, same packed-bf16 blend, same accumulation order: the bar below is one bf16 rounding, the kernels
    are in fact bit-identical.  sigma = 6 px exercises the out-of-box global path, 44 x 88 partial tiles and image borders."""
    import os
    import subprocess
    import sys

    code = r"""
import sys, torch
sys.path.insert(0, %r)
from vfi_b200 import ops
g = torch.Generator().manual_seed(77)
B, H, W = 2, 44, 88
feat = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
src = ops.Planes(B, H, W, 'cuda', zero_tail=True)
src.set_tail(torch.randn(B, 3, H, W, generator=g))
c27 = torch.randn(B, 27, H, W, generator=g)
c27[:, :9] *= 6.0
c27[:, 18:] *= 6.0
c27 = c27.to(torch.bfloat16).cuda()
w = ((torch.rand(67, 67, 3, 3, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16).cuda()
b = (torch.randn(67, generator=g) * 0.01).to(torch.bfloat16).cuda()
y = ops.deform_conv2d_fused(feat, src.tail_nchw(3), c27, w, b).to_nchw()
torch.save(y.cpu(), sys.argv[1])
""" % str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    import tempfile

    outs = []
    with tempfile.TemporaryDirectory() as d:
        for variant in ("v7", "v6"):                             # anything but "v6" selects the default kernel
            path = os.path.join(d, variant + ".pt")
            env = dict(os.environ, VFI_DCN_KERNEL=variant)
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(torch.load(path).float())
    assert torch.isfinite(outs[0]).all() and torch.isfinite(outs[1]).all()
    assert relerr(outs[1], outs[0]) <= 5e-3                       # one bf16 ulp of the largest value
    assert float((outs[0] != outs[1]).float().mean()) <= 0.05


@pytest.mark.parametrize("B,H,W,sigma,gdt", [(1, 8, 16, 1.5, torch.bfloat16), (2, 24, 40, 1.5, torch.bfloat16),
                                            (1, 40, 64, 6.0, torch.float32), (3, 64, 96, 0.0, torch.bfloat16),
                                            (2, 256, 256, 1.5, torch.bfloat16)])
def test_dcn_weight_grad_tensor_core(B, H, W, sigma, gdt):
    """grad_weight / grad_bias as a pixel-reduction GEMM on tcgen05 (A = grad_out^T in tensor memory, B = the producers'
    sample tile read MN-major, accumulators persistent in TMEM over the CTA's tiles): against the fp32 oracle on the
    bf16-rounded operands.  24 x 40 has partial tiles; sigma = 6 exercises the out-of-box global path; 2 x 256 x 256 is
    1024 tiles, i.e. seven per CTA to accumulate over in tensor memory."""
    from vfi_b200 import ops

    z = rand_dcn(B, 67, 67, H, W, sigma, seed=91 + H)
    xr, offr, mr = bf16_round(z["x"]), bf16_round(z["offset"]), bf16_round(z["mask"])
    g_in = torch.from_numpy(z["grad_out"]).to(gdt)
    ref = oracle.dcn_bwd(bf16_round(z["grad_out"]), xr, offr, mr, z["weight"])
    gw, gb = ops.dcn_weight_grad_tc(g_in.to(DEV), cu(z["x"], torch.bfloat16), cu(z["offset"], torch.bfloat16),
                                    cu(z["mask"], torch.bfloat16), 67)
    assert relerr(gw, ref[3]) <= 1e-2
    assert relerr(gb, ref[4]) <= 1e-2
    # channels_last grad_out (what a cuDNN backward hands over after a channels_last forward) takes the same kernel
    gw_cl, gb_cl = ops.dcn_weight_grad_tc(g_in.to(DEV).contiguous(memory_format=torch.channels_last),
                                          cu(z["x"], torch.bfloat16), cu(z["offset"], torch.bfloat16),
                                          cu(z["mask"], torch.bfloat16), 67)
    assert relerr(gw_cl, ref[3]) <= 1e-2 and relerr(gb_cl, ref[4]) <= 1e-2
    # autograd: a tensor-core forward uses the tensor-core weight gradient
    x = cu(z["x"], torch.bfloat16).requires_grad_(True)
    w = cu(z["weight"], torch.bfloat16).requires_grad_(True)
    b = cu(z["bias"], torch.bfloat16).requires_grad_(True)
    out = vfi_b200.deform_conv2d(x, cu(z["offset"], torch.bfloat16), w, b, stride=1, padding=1, dilation=1,
                                 mask=cu(z["mask"], torch.bfloat16), math="bf16_tc")
    out.backward(cu(z["grad_out"], torch.bfloat16))
    ref2 = oracle.dcn_bwd(bf16_round(z["grad_out"]), xr, offr, mr, bf16_round(z["weight"]))
    assert relerr(w.grad, ref2[3]) <= 2e-2 and relerr(b.grad, ref2[4]) <= 2e-2 and relerr(x.grad, ref2[0]) <= 2e-2


@pytest.mark.parametrize("B,H,W,sigma", [(1, 8, 16, 1.5), (2, 24, 40, 1.5), (1, 40, 64, 6.0), (2, 32, 48, 0.0), (1, 9, 13, 3.0),
                                         (2, 16, 24, 0.2)])
def test_dcn_data_grads_from_column_gradient(B, H, W, sigma):
    """grad_x / grad_offset / grad_mask of the bf16 training path: column gradient by a dense GEMM, then the
    channels-last gather / vector-reduction kernel (vfi_dcn_bwd_data_cols), against the fp32 oracle on bf16-rounded
    operands.  9 x 13 has a ragged last CTA and an odd width; sigma = 6 sends samples outside the image; sigma = 0 puts
    every sample on an integer position (lh = lw = 0 corner weights); sigma = 0 and 0.2 make neighbouring taps share corners,
    the case the kernel merges in registers."""
    z = rand_dcn(B, 67, 67, H, W, sigma, seed=131 + H)
    bf = torch.bfloat16
    x = cu(z["x"], bf).requires_grad_(True)
    off = cu(z["offset"], bf).requires_grad_(True)
    m = cu(z["mask"], bf).requires_grad_(True)
    w = cu(z["weight"], bf)
    out = vfi_b200.deform_conv2d(x, off, w, cu(z["bias"], bf), stride=1, padding=1, dilation=1, mask=m, math="bf16_tc")
    out.backward(cu(z["grad_out"], bf))
    ref = oracle.dcn_bwd(bf16_round(z["grad_out"]), bf16_round(z["x"]), bf16_round(z["offset"]), bf16_round(z["mask"]),
                         bf16_round(z["weight"]))
    assert x.grad.shape == x.shape and off.grad.shape == off.shape and m.grad.shape == m.shape
    assert relerr(x.grad, ref[0]) <= 2e-2
    assert relerr(off.grad, ref[1]) <= 2e-2
    assert relerr(m.grad, ref[2]) <= 2e-2
    # channels_last leaves (what layers 2 and 3 see in channels_last training): same gradients up to reduction order
    cl = torch.channels_last
    x3 = cu(z["x"], bf).contiguous(memory_format=cl).requires_grad_(True)
    off3 = cu(z["offset"], bf).contiguous(memory_format=cl).requires_grad_(True)
    m3 = cu(z["mask"], bf).contiguous(memory_format=cl).requires_grad_(True)
    out3 = vfi_b200.deform_conv2d(x3, off3, w, cu(z["bias"], bf), stride=1, padding=1, dilation=1, mask=m3, math="bf16_tc")
    assert torch.equal(out3, out)
    out3.backward(cu(z["grad_out"], bf).contiguous(memory_format=cl))
    assert relerr(x3.grad, x.grad) <= 1e-2 and relerr(off3.grad, off.grad) <= 1e-3 and relerr(m3.grad, m.grad) <= 1e-3
    # fp32 tensors with math="bf16_tc" (the unmodified fp32 model opting into the tensor cores): operands are rounded inside
    # (activations / weights to bf16, offsets / mask to fp16), gradients come back fp32 and take the same kernels
    leaves = [cu(z[k]).requires_grad_(True) for k in ("x", "offset", "mask", "weight", "bias")]
    out5 = vfi_b200.deform_conv2d(leaves[0], leaves[1], leaves[3], leaves[4], stride=1, padding=1, dilation=1, mask=leaves[2],
                                  math="bf16_tc")
    out5.backward(cu(z["grad_out"]))
    f16 = lambda a: torch.from_numpy(a).half().float().numpy()     # the oracle samples where the kernels sample
    ref5 = oracle.dcn_bwd(bf16_round(z["grad_out"]), bf16_round(z["x"]), f16(z["offset"]), f16(z["mask"]), bf16_round(z["weight"]))
    for t, r in zip(leaves, ref5):
        assert t.grad.dtype == torch.float32 and relerr(t.grad, r) <= 2e-2
    # the fp32 CUDA-core kernel on the same operands agrees to the same bar
    x2 = cu(z["x"], bf).float().requires_grad_(True)
    off2 = cu(z["offset"], bf).float().requires_grad_(True)
    m2 = cu(z["mask"], bf).float().requires_grad_(True)
    out2 = vfi_b200.deform_conv2d(x2, off2, w.float(), cu(z["bias"], bf).float(), stride=1, padding=1, dilation=1, mask=m2,
                                  math="fp32")
    out2.backward(cu(z["grad_out"], bf).float())
    assert relerr(x.grad, x2.grad) <= 2e-2 and relerr(off.grad, off2.grad) <= 2e-2 and relerr(m.grad, m2.grad) <= 2e-2


@pytest.mark.parametrize("B,H,W,cl,gdt", [(2, 24, 40, False, torch.bfloat16), (1, 9, 13, False, torch.float32), (3, 16, 24, True, torch.bfloat16),
                                          (1, 128, 130, True, torch.float32)])
def test_dcn_column_gradient_gemm_on_tcgen05(B, H, W, cl, gdt):
    """vfi_dcn_gcol: gcol[p, k * 72 + c] = sum_o grad_out[p, o] * W[o, c, k] on the tensor cores (grad_out read where it lies, A
    operand in tensor memory, weights resident in shared memory, TMA stores) against the same product in fp32 on the
    bf16-rounded operands.  P = 117 and 16,640 exercise the partial last tile; channels_last / fp32 grad_out the loaders."""
    import ctypes

    from vfi_b200 import _lib, ops

    g = torch.Generator(device=DEV).manual_seed(B * 1000 + H)
    go = torch.randn(B, 67, H, W, device=DEV, generator=g).to(gdt)
    if cl:
        go = go.contiguous(memory_format=torch.channels_last)
    w = ((torch.rand(67, 67, 3, 3, device=DEV, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16)
    P = B * H * W
    gcol = torch.full((P, 648), float("nan"), dtype=torch.bfloat16, device=DEV)
    lib = _lib.load()
    ws = torch.empty(int(lib.vfi_dcn_gcol_workspace_bytes()), dtype=torch.uint8, device=DEV)
    _lib.check(lib.vfi_dcn_gcol(_lib.ref(_lib.desc(go)), w.data_ptr(), _lib.dtype_code(w.dtype), 67, gcol.data_ptr(), 648, ws.data_ptr(),
                                ws.numel(), _lib.stream_handle(torch.device(DEV))), "vfi_dcn_gcol")
    rows = go.to(torch.bfloat16).float().permute(0, 2, 3, 1).reshape(P, 67)
    ref = rows @ ops.cols_weight_matrix(w, torch.float32)[:67]
    assert torch.isfinite(gcol.float()).all()
    assert relerr(gcol, ref) <= 5e-3                      # bf16 rounding of the result
    assert float(gcol.view(P, 9, 72)[:, :, 67:].abs().max()) == 0.0      # pad columns are zeros


def test_dcn_training_gradients_config3_full_size_vs_stock_torchvision_cuda():
    """BASELINE config 3 at full size (16 x 67 x 256 x 256, forward + backward of one DCNv2 layer): all five gradients of
    (a) the fp32 path and (b) the bf16 tensor-core training path (tcgen05 forward and weight gradient, column-gradient
    data backward, channels_last tensors) against stock torchvision's CUDA kernels in fp32 on the same bf16-rounded
    operands.  Stock grad_x uses fp32 atomics in a different order, hence the relative bars."""
    B, C, H, W = 16, 67, 256, 256
    g = torch.Generator(device=DEV).manual_seed(303)
    bf = torch.bfloat16
    x = torch.randn(B, C, H, W, device=DEV, generator=g).to(bf)
    off = (1.5 * torch.randn(B, 18, H, W, device=DEV, generator=g)).to(bf)
    m = torch.sigmoid(torch.randn(B, 9, H, W, device=DEV, generator=g)).to(bf)
    w = ((torch.rand(C, C, 3, 3, device=DEV, generator=g) * 2 - 1) / 603 ** 0.5).to(bf)
    b = ((torch.rand(C, device=DEV, generator=g) * 2 - 1) / 603 ** 0.5).to(bf)
    go = torch.randn(B, C, H, W, device=DEV, generator=g).to(bf)

    def grads(fn, dt, fmt):
        leaves = [t.to(dt).contiguous(memory_format=fmt).requires_grad_(True) if t.dim() == 4 and t.shape[-1] == W
                  else t.to(dt).requires_grad_(True) for t in (x, off, m, w, b)]
        out = fn(*leaves)
        out.backward(go.to(dt).contiguous(memory_format=fmt))
        return out.detach().float(), [t.grad.float() for t in leaves]

    ref_out, ref = grads(lambda x_, o_, m_, w_, b_: torch_ref.dcn_stock(x_, o_, m_, w_, b_), torch.float32, torch.contiguous_format)
    out32, g32 = grads(lambda x_, o_, m_, w_, b_: vfi_b200.deform_conv2d(x_, o_, w_, b_, stride=1, padding=1, dilation=1, mask=m_,
                                                                          math="fp32"), torch.float32, torch.contiguous_format)
    out16, g16 = grads(lambda x_, o_, m_, w_, b_: vfi_b200.deform_conv2d(x_, o_, w_, b_, stride=1, padding=1, dilation=1, mask=m_,
                                                                          math="bf16_tc"), bf, torch.channels_last)
    names = ("grad_x", "grad_offset", "grad_mask", "grad_weight", "grad_bias")
    assert maxabs(out32, ref_out) <= 1e-5 * max(1.0, float(ref_out.abs().max()))
    assert relerr(out16, ref_out) <= 1e-2
    for a, r, n in zip(g32, ref, names):
        assert relerr(a, r) <= 1e-4, n          # summation order over 1 M pixels (weights) / atomics (x)
    for a, r, n in zip(g16, ref, names):
        assert relerr(a, r) <= 2e-2, n


def test_dcn_tensor_core_path_1080p_vs_fp32_kernel():
    """Full 1080p frame: tcgen05 result against this library's fp32 parity kernel on the same bf16-rounded inputs."""
    g = torch.Generator(device=DEV).manual_seed(41)
    x = torch.randn(1, 67, 1080, 1920, device=DEV, generator=g).to(torch.bfloat16)
    off = (1.5 * torch.randn(1, 18, 1080, 1920, device=DEV, generator=g)).to(torch.bfloat16)
    m = torch.sigmoid(torch.randn(1, 9, 1080, 1920, device=DEV, generator=g)).to(torch.bfloat16)
    w = ((torch.rand(67, 67, 3, 3, device=DEV, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16)
    b = (torch.randn(67, device=DEV, generator=g) * 0.01).to(torch.bfloat16)
    out = vfi_b200.deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m, math="bf16_tc")
    ref = vfi_b200.deform_conv2d(x.float(), off.float(), w.float(), b.float(), stride=1, padding=1, dilation=1,
                                 mask=m.float(), math="fp32")
    err = float((out.float() - ref).abs().max()) / float(ref.abs().max())
    assert err <= 1e-2, err


def test_hot_path_bf16_tensor_core_vs_fp32_reference_psnr():
    """warp -> fused buffer -> 3 x DCN on the tensor-core path against the fp32 oracle chain on bf16-rounded inputs."""
    from vfi_b200.hotpath import HotPath, synthetic_inputs, synthetic_weights

    B, H, W = 1, 48, 80
    frame2, flow, feat, convs = synthetic_inputs(B, H, W, dtype=torch.bfloat16, device=DEV, seed=5)
    ws, bs = synthetic_weights(dtype=torch.bfloat16, device=DEV)
    out = HotPath(ws, bs, math="bf16_tc").run(frame2, flow, feat.contiguous(memory_format=torch.channels_last), convs)
    out = out.to_nchw()
    f = lambda t: t.float().cpu().numpy()  # noqa: E731
    x = np.concatenate([f(feat), oracle.warp_fwd(f(frame2), f(flow))], 1)
    for w, b, c in zip(ws, bs, convs):
        o, m = oracle.pack_split(f(c))
        x = oracle.dcn_fwd(x, bf16_round(o), bf16_round(m), f(w), f(b))
    assert relerr(out, x) <= 2e-2     # three chained bf16 layers


def _stock_layer_frame(x67_bf16, conv27_bf16, w, b):
    """One DCNv2 layer of one frame by STOCK torchvision CUDA in fp32 on the same bf16-rounded tensors (the bf16 oracle of
    SURVEY.md section 8c): offsets / mask from the 27-channel tensor as ema_vfi.py:57-59 computes them, sigmoid in the tensor's
    dtype (what torch.sigmoid on a bf16 tensor returns).  One frame at a time bounds torchvision's columns buffer (5 GB at
    1080p, 20 GB at 4K)."""
    off, m = torch_ref.pack_split(conv27_bf16)
    return torch_ref.dcn_stock(x67_bf16.float(), off.float(), m.float(), w.float(), b.float())


def test_hot_path_cfg2_full_size_vs_stock_torchvision_cuda():
    """BASELINE config 2 at full size -- 8 x 1080p, bf16, exactly what bench.py times (HotPath.run: warp into tail records,
    three fused tcgen05 layers on planes) -- against an INDEPENDENT oracle: stock aten grid_sample and stock torchvision
    deform_conv2d CUDA kernels in fp32.  Every layer of every frame is checked on the layer's own (bf16) input at the
    north star's bf16 bar, max|delta| / max|ref| <= 1e-2; the chain the bench runs is tied to those layers bit for bit."""
    from vfi_b200 import ops
    from vfi_b200.hotpath import HotPath, synthetic_inputs, synthetic_weights

    B, H, W = 8, 1080, 1920
    frame2, flow, feat, convs = synthetic_inputs(B, H, W, dtype=torch.bfloat16, device=DEV, seed=1234)
    feat = feat.contiguous(memory_format=torch.channels_last)
    ws, bs = synthetic_weights(dtype=torch.bfloat16, device=DEV)
    out = HotPath(ws, bs, math="bf16_tc").run(frame2, flow, feat, convs)
    # the same chain with every intermediate kept
    src = ops.Planes(B, H, W, DEV, zero_tail=True)
    ops.warp(frame2, flow, out=src.tail_nchw(3), tail_record=True)
    layers = [(feat, src.tail_nchw(3))]
    for w, b, c27 in zip(ws, bs, convs):
        y = ops.deform_conv2d_fused(layers[-1][0], layers[-1][1], c27, w, b, math="bf16_tc")
        layers.append((y.main_nchw, y.tail_nchw()))
    assert torch.equal(out.main, layers[-1][0].permute(0, 2, 3, 1)) and torch.equal(out.tail_nchw(), layers[-1][1])
    worst = 0.0
    for f in range(B):
        warped = torch_ref.warp(frame2[f:f + 1].float(), flow[f:f + 1].float())     # stock CUDA grid_sample (reciprocal division)
        got = src.tail_nchw(3)[f:f + 1].float()
        assert float((got - warped).abs().max()) <= 1e-2 * float(warped.abs().max())
        for i, (w, b, c27) in enumerate(zip(ws, bs, convs)):
            x67 = torch.cat((layers[i][0][f:f + 1], layers[i][1][f:f + 1]), dim=1)
            ref = _stock_layer_frame(x67, c27[f:f + 1], w, b)
            got = torch.cat((layers[i + 1][0][f:f + 1], layers[i + 1][1][f:f + 1]), dim=1).float()
            err = float((got - ref).abs().max()) / float(ref.abs().max())
            worst = max(worst, err)
            assert err <= 1e-2, (f, i, err)
            del ref, got, x67
    print(f"cfg2 full size: worst layer error {worst:.2e} of max|ref| (bar 1e-2)")


@pytest.mark.parametrize("flow_kind,sigma", [("smooth", 64.0), ("iid", 64.0)])
def test_hot_path_cfg4_4k_vs_stock_torchvision_cuda(flow_kind, sigma):
    """BASELINE config 4 (4K frame pair, batch 1, bf16, large-displacement flow -- smooth and incoherent): the warp's tail
    records against stock grid_sample, and the fused tcgen05 layers against stock torchvision CUDA fp32 on the same tensors."""
    from vfi_b200 import ops
    from vfi_b200.hotpath import synthetic_inputs, synthetic_weights

    B, H, W = 1, 2160, 3840
    frame2, flow, feat, convs = synthetic_inputs(B, H, W, dtype=torch.bfloat16, device=DEV, seed=4, flow_sigma=sigma, flow_kind=flow_kind)
    feat = feat.contiguous(memory_format=torch.channels_last)
    ws, bs = synthetic_weights(dtype=torch.bfloat16, device=DEV)
    src = ops.Planes(B, H, W, DEV, zero_tail=True)
    ops.warp(frame2, flow, out=src.tail_nchw(3), division="reciprocal", tail_record=True)
    warped = torch_ref.warp(frame2.float(), flow.float())
    assert float((src.tail_nchw(3).float() - warped).abs().max()) <= 1e-2 * float(warped.abs().max())
    del warped
    x = (feat, src.tail_nchw(3))
    for i, (w, b, c27) in enumerate(zip(ws, bs, convs)):
        if i == 2 and flow_kind == "iid":
            break                                        # two layers suffice for the second flow class (20 GB of columns each)
        y = ops.deform_conv2d_fused(x[0], x[1], c27, w, b, math="bf16_tc")
        ref = _stock_layer_frame(torch.cat(x, dim=1), c27, w, b)
        got = y.to_nchw().float()
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        assert err <= 1e-2, (i, err)
        del ref, got
        x = (y.main_nchw, y.tail_nchw())


def test_interpolated_frame_psnr_delta_bf16_hot_path():
    """North-star gate: interpolated-frame PSNR delta <= 0.01 dB.  The network of the model_psnr_256 golden (BASELINE config 1
    size; the unmodified reference's CPU output and Middlebury frame11 as ground truth are in the fixture) runs on the GPU
    three ways: stock ops, the fp32 drop-in, and the bf16 tensor-core drop-in; each output's PSNR against the ground-truth
    frame is compared with the PSNR of the reference's own output."""
    from test_refmodel import psnr, psnr_inputs, psnr_model
    from vfi_b200.refmodel import StockInterpolator

    z = load_golden("model_psnr_256")
    model = psnr_model(z, DEV)
    a, b, gt = (t.to(DEV) for t in psnr_inputs(z))
    ref_out = torch.from_numpy(z["model_out"]).to(DEV)
    ref_psnr = psnr(ref_out, gt)
    with torch.no_grad():
        stock = model(a, b)
    assert maxabs(stock, ref_out) <= 2e-4                      # cuDNN vs the reference's CPU kernels
    results = {"stock_cuda": psnr(stock, gt)}
    for math in ("fp32", "bf16_tc"):
        vfi_b200.install(StockInterpolator, math=math)
        try:
            with torch.no_grad():
                out = model(a, b)
        finally:
            vfi_b200.uninstall()
        results[math] = psnr(out, gt)
        if math == "fp32":
            assert maxabs(out, stock) <= 2e-5
    # the reference's own GPU inference mode (inference.py:158-159: no_grad + CUDA autocast), stock and with the fused drop-in
    with torch.no_grad(), torch.autocast("cuda"):
        stock_amp = model(a, b).float()
    results["stock_cuda_autocast"] = psnr(stock_amp, gt)
    before = dict(vfi_b200.dropin.call_counts())
    vfi_b200.install(StockInterpolator, fuse=True)
    try:
        with torch.no_grad(), torch.autocast("cuda"):
            fused = model(a, b).float()
    finally:
        vfi_b200.uninstall()
    after = vfi_b200.dropin.call_counts()
    assert (after["fused_block"] - before["fused_block"], after["fused_cat"] - before["fused_cat"],
            after["fused_warp"] - before["fused_warp"], after["dcn"] - before["dcn"]) == (3, 1, 1, 0)
    results["fused_autocast"] = psnr(fused, gt)
    assert maxabs(fused, stock_amp) <= 2e-2
    print(f"PSNR vs ground truth: reference {ref_psnr:.4f} dB, " + ", ".join(f"{k} {v:.4f} dB" for k, v in results.items()))
    for k, v in results.items():
        assert abs(v - ref_psnr) <= 0.01, (k, v, ref_psnr)


# ------------------------------------------------------------------------------------------------------------ path
def test_model_hot_path_fixture_through_the_dropin():
    """Replay of the tensors recorded inside the unmodified EMA_VFI.forward (golden model_24x32): warp -> cat -> 3 x
    (split, DCN) with both seams patched, compared with what the reference model computed."""
    import torchvision.ops

    z = load_golden("model_24x32")
    vfi_b200.install(torch_ref.WarpHost, division="ieee")     # the golden vectors are the reference's CPU output
    try:
        host = torch_ref.WarpHost()
        warped = host.warp(cu(z["frame2"]), cu(z["feat"]), cu(z["flow"]))
        assert maxabs(warped, z["warped"]) <= 1e-5
        x = torch.cat([cu(z["feat"]), warped], 1)
        for i in range(3):
            blk = torchvision.ops.DeformConv2d(67, 67, 3, padding=1).to(DEV)
            blk.load_state_dict({"weight": cu(z[f"dcn_weight_{i}"]), "bias": cu(z[f"dcn_bias_{i}"])})
            off, m = torch_ref.pack_split(cu(z[f"conv27_{i}"]))
            x = blk(x, off, m)
            assert maxabs(x, z[f"block_out_{i}"]) <= 2e-5, i
        assert vfi_b200.dropin.call_counts()["dcn"] >= 3
    finally:
        vfi_b200.uninstall()


def test_fusion_block_training_step_matches_stock_torchvision():
    """A ModulatedDeformConvPack-shaped block trained one step through the drop-in vs through stock torchvision."""
    torch.manual_seed(0)
    blk = torch_ref.FusionBlock(67).to(DEV)
    with torch.no_grad():
        blk.offset_conv.weight.normal_(0, 0.02)
        blk.offset_conv.bias.normal_(0, 0.5)
    x = torch.randn(2, 67, 32, 48, device=DEV)
    blk(x).square().mean().backward()
    ref = {k: p.grad.clone() for k, p in blk.named_parameters()}
    blk.zero_grad()
    vfi_b200.install()
    try:
        blk(x).square().mean().backward()
    finally:
        vfi_b200.uninstall()
    for k, p in blk.named_parameters():
        assert maxabs(p.grad, ref[k]) <= tol(ref[k], 3e-5), k


@pytest.mark.parametrize("c27_layout", ["nchw", "channels_last"])
def test_fused_training_block_on_records_matches_the_unfused_ops(c27_layout):
    """ops.deform_conv2d_block (record activations in / out, raw 27-channel offset_conv output in, glue folded into the kernels
    forward AND backward) against the same block built from the reference's glue as stock ops (chunk / cat / sigmoid, ema_vfi.py:
    57-59) around vfi_b200.deform_conv2d with the same bf16 tensor-core math, and against the fp32 oracle for the forward."""
    from vfi_b200 import ops

    g = torch.Generator().manual_seed(91)
    B, H, W, C = 2, 32, 48, 67
    x = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16)
    c27 = torch.randn(B, 27, H, W, generator=g)
    c27[:, :9] *= 1.5
    c27[:, 18:] *= 1.5
    c27 = c27.to(torch.bfloat16)
    w = ((torch.rand(C, C, 3, 3, generator=g) * 2 - 1) / 603 ** 0.5)
    b = ((torch.rand(C, generator=g) * 2 - 1) / 603 ** 0.5)
    gy = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16)

    # ---- unfused reference: stock glue + the drop-in op
    xr = x.to(DEV).requires_grad_(True)
    cr = c27.to(DEV).requires_grad_(True)
    wr, br = w.to(DEV).to(torch.bfloat16).requires_grad_(True), b.to(DEV).to(torch.bfloat16).requires_grad_(True)
    o1, m, o2 = cr.chunk(3, dim=1)
    yr = vfi_b200.deform_conv2d(xr, torch.cat((o1, o2), 1), wr, br, stride=1, padding=1, dilation=1, mask=torch.sigmoid(m), math="bf16_tc")
    yr.backward(gy.to(DEV))

    # ---- fused block on records (fp32 master parameters passed as they are)
    x72 = ops.records_buffer(B, H, W, DEV, zero=True)
    x72[:, :C] = x.to(DEV)
    x72[:, 68:71] = x[:, 64:67].to(DEV)
    x72.requires_grad_(True)
    cf = c27.to(DEV)
    if c27_layout == "channels_last":
        cf = cf.contiguous(memory_format=torch.channels_last)
    cf.requires_grad_(True)
    wf, bf = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y72 = ops.deform_conv2d_block(x72, cf, wf, bf)
    assert ops._is_records(y72)
    g72 = torch.zeros_like(y72)
    g72[:, :C] = gy.to(DEV)
    g72[:, 67:] = 3.0                                            # whatever reaches the derived channels must not matter
    y72.backward(g72)

    off, msk = oracle.pack_split(c27.float().numpy())
    ref = oracle.dcn_fwd(x.float().numpy(), off, bf16_round(msk), bf16_round(w.numpy()), b.numpy())
    assert relerr(y72[:, :C], ref) <= 1e-2
    assert relerr(y72[:, :C], yr) <= 1e-2
    assert float(y72[:, 67].abs().max()) == 0.0 and torch.equal(y72[:, 68:72], y72[:, 64:68])     # a valid record again
    assert relerr(x72.grad[:, :C], xr.grad) <= 2e-2 and float(x72.grad[:, 67:].abs().max()) == 0.0
    assert relerr(cf.grad, cr.grad) <= 2e-2
    assert relerr(wf.grad, wr.grad) <= 2e-2 and relerr(bf.grad, br.grad) <= 2e-2
    assert wf.grad.dtype == torch.float32


def test_training_step_fused_blocks_match_the_unfused_step():
    """TrainStep(fused=True) (record activations, ops.deform_conv2d_block, padded offset_conv) computes the gradients of the
    unfused step: same parameters, same inputs, bf16 tolerance."""
    from vfi_b200 import shard
    from vfi_b200.trainstep import TrainStep

    topo = shard.Topology(rank=0, world=1, local_rank=0)
    a = TrainStep(topo, DEV, math="bf16_tc", global_batch=2, size=64, fused=False)
    b = TrainStep(topo, DEV, math="bf16_tc", global_batch=2, size=64, fused=True)
    assert b.fused
    a.step()
    b.step()
    torch.cuda.synchronize()
    assert relerr(b.bucket.flat, a.bucket.flat) <= 2e-2
    assert relerr(b.flow.grad, a.flow.grad) <= 3e-2


def test_training_step_cuda_graph_replay_matches_eager():
    """TrainStep.capture(): the whole cfg3 step (bucket zeroing, forward, backward into the flat gradient bucket) replayed as
    one CUDA graph gives the gradients of the eager step -- same kernels, same order; the only non-determinism is the order of
    the fp32 atomics, hence a relative bar instead of equality."""
    from vfi_b200 import shard
    from vfi_b200.trainstep import TrainStep

    topo = shard.Topology(rank=0, world=1, local_rank=0)
    ts = TrainStep(topo, DEV, math="bf16_tc", global_batch=2, size=64)
    ts.step()
    torch.cuda.synchronize()
    ref_flat, ref_flow = ts.bucket.flat.clone(), ts.flow.grad.clone()
    ts.capture()
    assert ts.graph is not None and ts.graph_launches > 0
    for _ in range(2):
        ts.step()
    torch.cuda.synchronize()
    assert torch.isfinite(ts.bucket.flat).all()
    assert relerr(ts.bucket.flat, ref_flat) <= 1e-3
    assert relerr(ts.flow.grad, ref_flow) <= 1e-2      # a bf16 tensor: one ulp is 4e-3 of the value


def test_smoke_entry():
    import __graft_entry__

    __graft_entry__.smoke()
