#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Needs ``/root/reference`` (authoring container only; the GPU box does not have it, which is why the outputs are
committed).  Everything is computed on CPU in fp32 by the reference's own code path:

* ``EMA_VFI.warp``                       (/root/reference/src/models/ema_vfi.py:149-171)
* ``ModulatedDeformConvPack.forward``    (ema_vfi.py:53-60)  -> ``torchvision.ops.DeformConv2d``
* ``EMA_VFI.forward``                    (ema_vfi.py:110-147), with hooks that record the tensors entering and
  leaving the hot path (flow, feat, warped frame, the three 27-channel offset_conv outputs, each block's output)
* autograd through the above for every gradient the path produces.

torch / torchvision versions are recorded in each file (``meta``).
"""
from __future__ import annotations

import json
import os
import struct
import sys
from pathlib import Path

import numpy as np
import torch
import torchvision

REF = Path(os.environ.get("VFI_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF))
from src.models.ema_vfi import EMA_VFI, ModulatedDeformConvPack  # noqa: E402  (the unmodified reference)

OUT = Path(__file__).resolve().parent
META = json.dumps({"torch": torch.__version__, "torchvision": torchvision.__version__,
                   "reference": "424635328/video-frame-interpolation src/models/ema_vfi.py"})
torch.set_num_threads(1)  # fixed summation order inside the BLAS call


def save(name, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(OUT / f"{name}.npz", meta=np.array(META), **arrays)
    print(f"{name}.npz", {k: tuple(v.shape) for k, v in arrays.items()})


def ref_warp(frame2, flow):
    """Call the reference's bound method exactly as ema_vfi.py:130 does (feature is only probed for .is_cuda)."""
    return EMA_VFI.warp(None, frame2, frame2, flow)


def warp_case(name, src, flow, with_grad_src=False):
    flow = flow.clone().requires_grad_(True)
    src = src.clone().requires_grad_(with_grad_src)
    out = ref_warp(src, flow)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(99))
    out.backward(g)
    extra = {"grad_src": src.grad} if with_grad_src else {}
    save(name, src=src, flow=flow, out=out, grad_out=g, grad_flow=flow.grad, **extra)


def read_flo(path):
    with open(path, "rb") as f:
        magic, w, h = struct.unpack("<fii", f.read(12))
        assert abs(magic - 202021.25) < 1e-3
        return np.frombuffer(f.read(), dtype="<f4").reshape(h, w, 2).copy()


def main():
    g = torch.Generator().manual_seed(20261018)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    ru = lambda *s: torch.rand(*s, generator=g)  # noqa: E731

    # ---------------------------------------------------------------- warp
    # random frames, sigma=3 px flow plus a band of huge displacements that leave the frame on every side
    B, C, H, W = 2, 3, 19, 23
    flow = 3.0 * rn(B, 2, H, W)
    flow[0, 0, :, :3] -= 6.0
    flow[0, 0, :, -3:] += 6.0
    flow[1, 1, :3, :] -= 6.0
    flow[1, 1, -3:, :] += 6.0
    flow[1, :, 9, 11] = torch.tensor([1e6, -1e6])
    warp_case("warp_rand", rn(B, C, H, W), flow, with_grad_src=True)

    # exactly-integer displacements, including samples landing exactly on -1, 0, W-1 and W
    H, W = 8, 12
    flow = torch.zeros(1, 2, H, W)
    flow[0, 0] = torch.tensor([-1.0, 0.0, 1.0, 2.0, -2.0, 0.0, 3.0, -3.0, 0.5, -0.5, 1.0, 0.0]).expand(H, W)
    flow[0, 1] = torch.tensor([-1.0, 0.0, 1.0, 2.0, -2.0, 0.0, 1.0, 0.25]).unsqueeze(1).expand(H, W)
    warp_case("warp_integer", rn(1, 3, H, W), flow)

    # model-like flow (|f| < 0.04 px at random init) on a wide, odd-sized frame
    warp_case("warp_tiny_flow", rn(1, 3, 5, 131), 0.03 * rn(1, 2, 5, 131))

    # degenerate sizes: a single column / single row (max(W-1, 1) in ema_vfi.py:165-166)
    warp_case("warp_w1", rn(1, 3, 6, 1), 0.7 * rn(1, 2, 6, 1))
    warp_case("warp_h1", rn(2, 2, 1, 9), 0.7 * rn(2, 2, 1, 9))

    # real image + real flow: Middlebury Urban2 crop, frame11 warped by the ground-truth flow10
    try:
        from PIL import Image

        fr = np.asarray(Image.open(REF / "data/processed/train/Urban2/frame11.png").convert("RGB"), np.float32) / 255
        fl = read_flo(REF / "data/processed/other-gt-flow/Urban2/flow10.flo")
        y0, x0, hh, ww = 200, 300, 40, 56
        src = torch.from_numpy(fr[y0:y0 + hh, x0:x0 + ww].transpose(2, 0, 1).copy())[None]
        src = (src - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        flow = torch.from_numpy(fl[y0:y0 + hh, x0:x0 + ww].transpose(2, 0, 1).copy())[None]
        flow[flow.abs() > 1e8] = 0.0
        warp_case("warp_urban2", src, flow)
    except Exception as e:  # pragma: no cover - the data blobs may be absent
        print("skip warp_urban2:", e)

    # ---------------------------------------------------------------- DCN (op level)
    def dcn_case(name, B, C, O, H, W, sigma, zero_mask_half=False):
        x = rn(B, C, H, W).requires_grad_(True)
        offset = (sigma * rn(B, 18, H, W)).requires_grad_(True)
        mask = (torch.full((B, 9, H, W), 0.5) if zero_mask_half else torch.sigmoid(rn(B, 9, H, W))).requires_grad_(True)
        bound = 1.0 / (C * 9) ** 0.5
        weight = ((ru(O, C, 3, 3) * 2 - 1) * bound).requires_grad_(True)
        bias = ((ru(O) * 2 - 1) * bound).requires_grad_(True)
        out = torchvision.ops.deform_conv2d(x, offset, weight, bias, stride=(1, 1), padding=(1, 1),
                                            dilation=(1, 1), mask=mask)
        go = rn(*out.shape)
        out.backward(go)
        save(name, x=x, offset=offset, mask=mask, weight=weight, bias=bias, out=out, grad_out=go,
             grad_x=x.grad, grad_offset=offset.grad, grad_mask=mask.grad, grad_weight=weight.grad,
             grad_bias=bias.grad)

    dcn_case("dcn_c67_sigma3", 2, 67, 67, 9, 11, 3.0)       # model geometry, samples leave the image on all sides
    dcn_case("dcn_c67_zero", 1, 67, 67, 6, 7, 0.0, True)      # random-init degenerate case: offsets 0, mask 0.5
    dcn_case("dcn_c5_o7_sigma1", 3, 5, 7, 7, 5, 1.0)          # general channel counts
    dcn_case("dcn_c67_sigma16", 1, 67, 67, 5, 33, 16.0)       # mostly out-of-bounds

    # ---------------------------------------------------------------- the pack (ema_vfi.py:23-60)
    torch.manual_seed(7)
    pack = ModulatedDeformConvPack(67, 67, kernel_size=3, padding=1, groups=1)
    with torch.no_grad():  # F4: offset_conv is zero-initialised; randomise it so the gather is exercised
        pack.offset_conv.weight.normal_(0, 0.02, generator=g)
        pack.offset_conv.bias.normal_(0, 0.5, generator=g)
    x = rn(1, 67, 8, 10)
    conv27 = pack.offset_conv(x)
    o1, m, o2 = torch.chunk(conv27, 3, dim=1)
    save("pack_c67", x=x, conv27=conv27, offset=torch.cat((o1, o2), 1), mask=torch.sigmoid(m),
         weight=pack.dcn_v2.weight, bias=pack.dcn_v2.bias, out=pack(x))

    # ---------------------------------------------------------------- whole model, hot-path tensors recorded
    torch.manual_seed(1234)
    model = EMA_VFI().eval()
    with torch.no_grad():
        for blk in model.attention_blocks:
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
        # random init gives |flow| < 0.04 px; scale the last motion conv so the warp actually moves pixels
        model.motion_estimation[-1].weight.mul_(40.0)
    rec = {}
    orig_warp = EMA_VFI.warp

    def spy_warp(self, frame2, feature, flow):
        out = orig_warp(self, frame2, feature, flow)
        rec.update(frame2=frame2, feat=feature, flow=flow, warped=out)
        return out

    EMA_VFI.warp = spy_warp
    hooks = []
    for i, blk in enumerate(model.attention_blocks):
        hooks.append(blk.offset_conv.register_forward_hook(lambda m, a, o, i=i: rec.__setitem__(f"conv27_{i}", o)))
        hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: rec.__setitem__(f"block_out_{i}", o)))
    f1, f2 = rn(1, 3, 24, 32), rn(1, 3, 24, 32)
    with torch.no_grad():
        out = model(f1, f2)
    EMA_VFI.warp = orig_warp
    for h in hooks:
        h.remove()
    wb = {}
    for i, blk in enumerate(model.attention_blocks):
        wb[f"dcn_weight_{i}"] = blk.dcn_v2.weight
        wb[f"dcn_bias_{i}"] = blk.dcn_v2.bias
    save("model_24x32", frame1=f1, frame2_in=f2, model_out=out, **rec, **wb)


def state_digest(sd):
    """sha256 over the float32 bytes of a state_dict in key order (pins seeded construction across machines)."""
    import hashlib

    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().float().numpy().tobytes())
    return h.hexdigest()


def psnr_case():
    """BASELINE config 1 size (256 x 256, batch 1): the unmodified reference interpolates Middlebury Beanbags frame10 / frame12;
    frame11 is the ground truth the PSNR gate of the north star is measured against (interpolated-frame PSNR delta of a
    low-precision hot path <= 0.01 dB).  Weights: seeded default init (digest recorded) with offset_conv randomised (F4) and
    the last motion convolution scaled so the warp moves pixels -- both stored, they are what the GPU test re-creates."""
    from PIL import Image

    SEED = 2026
    torch.manual_seed(SEED)
    model = EMA_VFI().eval()
    digest = state_digest(model.state_dict())
    g = torch.Generator().manual_seed(77)
    extra = {}
    with torch.no_grad():
        for i, blk in enumerate(model.attention_blocks):
            blk.offset_conv.weight.normal_(0, 0.02, generator=g)
            blk.offset_conv.bias.normal_(0, 0.5, generator=g)
            extra[f"offset_conv_weight_{i}"] = blk.offset_conv.weight.clone()
            extra[f"offset_conv_bias_{i}"] = blk.offset_conv.bias.clone()
        model.motion_estimation[-1].weight.mul_(40.0)
    y0, x0, n = 120, 200, 256
    frames = [np.asarray(Image.open(REF / f"data/processed/train/Beanbags/frame{i}.png").convert("RGB"))[y0:y0 + n, x0:x0 + n].copy()
              for i in (10, 11, 12)]
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    to_t = lambda a: ((torch.from_numpy(a.transpose(2, 0, 1).copy()).float() / 255.0)[None] - mean) / std  # noqa: E731  (inference.py:38-41)
    with torch.no_grad():
        out = model(to_t(frames[0]), to_t(frames[2]))
    save("model_psnr_256", frame_a=frames[0], frame_gt=frames[1], frame_b=frames[2], model_out=out, seed=np.array(SEED),
         motion_scale=np.array(40.0), state_digest=np.array(digest), **extra)
    gt = torch.from_numpy(frames[1].transpose(2, 0, 1).copy()).float()[None] / 255.0
    print("PSNR of the reference output vs frame11:", float(-10 * torch.log10(((out - gt) ** 2).mean())), "dB")


if __name__ == "__main__":
    if "--only-psnr" in sys.argv:
        psnr_case()
    else:
        main()
        psnr_case()
