"""Timing of the warp backward (grad_flow only, the model path; and with grad_frame) against its HBM byte model
P * (3 + 3 + 2 + 2) * sizeof."""
import importlib, json, sys
import torch
sys.path.insert(0, ".")
pkg = importlib.import_module("video-frame-interpolation_b200")


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {}
for name, (B, H, W) in (("cfg3", (16, 256, 256)), ("cfg2", (8, 1080, 1920))):
    src = torch.randn(B, 3, H, W, device="cuda")
    coarse = torch.randn(B, 2, max(H // 32, 2), max(W // 32, 2), device="cuda")
    flow = (torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True) * 8.0).requires_grad_(True)
    go = torch.randn(B, 3, H, W, device="cuda")
    out = pkg.warp(src, flow)
    ms = timed(lambda: torch.autograd.grad(out, flow, go, retain_graph=True))
    bytes_ = B * H * W * (3 + 3 + 2 + 2) * 4
    res[name] = {"ms_grad_flow": ms, "GBps": bytes_ / ms / 1e6, "frac_of_6527": bytes_ / ms / 1e6 / 6527.1}
    s2 = src.clone().requires_grad_(True)
    out2 = pkg.warp(s2, flow)
    res[name]["ms_grad_flow_and_frame"] = timed(lambda: torch.autograd.grad(out2, (s2, flow), go, retain_graph=True))
print(json.dumps(res))
