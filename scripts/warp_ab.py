"""Times the planar warp of the library selected by VFI_B200_LIB (default: the in-tree build) at the bench's cfg2 shape, with
the bench's smooth flow and the model-like N(0, 0.03^2) flow, bf16 and fp32; prints one JSON line with a checksum of the
results so that variants can be compared for bit-identity.  Every case runs twice: the TMA-staged kernel (default route, with
the count of tiles that fell back to L1) and round 1's L1-gather kernel (VFI_WARP_NO_STAGING).  A third flow class -- incoherent
16 px displacements -- shows the cost of the in-launch fallback.  Usage: [VFI_B200_LIB=...] python scripts/warp_ab.py [tag]"""
import hashlib
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import vfi_b200  # noqa: E402
from vfi_b200.hotpath import synthetic_inputs  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 8, 1080, 1920
frame2, flow, _, _ = synthetic_inputs(B, H, W, dtype=torch.float32, device=dev, seed=1234)
small = 0.03 * torch.randn(B, 2, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(77))
res = {"tag": sys.argv[1] if len(sys.argv) > 1 else "default"}
iid = 16.0 * torch.randn(B, 2, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(78))
for fname, fl32 in (("smooth", flow), ("model_like", small), ("iid16", iid)):
    for dname, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
        f2, fl = frame2.to(dt), fl32.to(dt)
        outs = {}
        for staging in (True, False):
            vfi_b200.ops.warp_tile_counts(reset=True)
            out = vfi_b200.warp(f2, fl, staging=staging, count_tiles=True)
            tiles = vfi_b200.ops.warp_tile_counts(reset=True)
            for _ in range(3):
                out = vfi_b200.warp(f2, fl, staging=staging)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                out = vfi_b200.warp(f2, fl, staging=staging)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 20
            px_bytes = 16 if dt == torch.bfloat16 else 32
            outs[staging] = out.clone()
            res[f"{fname}_{dname}_{'staged' if staging else 'l1'}"] = {
                "us": round(us, 1), "GBps": round(B * H * W * px_bytes / us / 1e3, 0), "tiles_staged_l1": tiles,
                "sha": hashlib.sha1(out.view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:10]}
        # torch.equal: -0.0 == +0.0 (the L1 kernel can return -0 where a zero-weight corner is negative; the staged one returns +0)
        res[f"{fname}_{dname}_staged_equals_l1"] = bool(torch.equal(outs[True], outs[False]))
print(json.dumps(res), flush=True)
