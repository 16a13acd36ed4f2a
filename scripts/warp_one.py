"""A few launches of the planar bf16 warp at cfg2 size (argv[1]: smooth | model_like | iid16; argv[2]: staged | l1) -- the
command ncu captures for the warp kernels."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import vfi_b200  # noqa: E402
from vfi_b200.hotpath import synthetic_inputs  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 8, 1080, 1920
kind = sys.argv[1] if len(sys.argv) > 1 else "model_like"
staging = (sys.argv[2] if len(sys.argv) > 2 else "staged") == "staged"
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.bfloat16
frame2, flow, _, _ = synthetic_inputs(B, H, W, dtype=torch.float32, device=dev, seed=1234)
if kind == "model_like":
    flow = 0.03 * torch.randn(B, 2, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(77))
elif kind == "iid16":
    flow = 16.0 * torch.randn(B, 2, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(78))
f2, fl = frame2.to(dt), flow.to(dt)
for _ in range(6):
    out = vfi_b200.warp(f2, fl, staging=staging)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
