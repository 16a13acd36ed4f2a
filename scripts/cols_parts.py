import torch, json, sys, importlib
sys.path.insert(0, ".")
pkg = importlib.import_module("video-frame-interpolation_b200"); ops = pkg.ops
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
dev = "cuda:0"; res = {}
n = 6 * 65536
for dt in (torch.float32, torch.bfloat16):
    g = torch.randn(n, 72, device=dev, dtype=dt); w = torch.randn(72, 648, device=dev, dtype=dt)
    torch.backends.cuda.matmul.allow_tf32 = False
    res[f"gemm_{dt}_ms_per_Mpx"] = timed(lambda: torch.matmul(g, w)) * (1048576 / n)
torch.backends.cuda.matmul.allow_tf32 = True
g = torch.randn(n, 72, device=dev); w = torch.randn(72, 648, device=dev)
res["gemm_tf32_ms_per_Mpx"] = timed(lambda: torch.matmul(g, w)) * (1048576 / n)
torch.backends.cuda.matmul.allow_tf32 = False
B, C, H, W = 16, 67, 256, 256
import os
sig = float(os.environ.get("COLS_SIGMA", "1.5"))
res["sigma"] = sig
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(B, C, H, W, device=dev).to(dt); off = (sig * torch.randn(B, 18, H, W, device=dev)).to(dt)
    m = torch.rand(B, 9, H, W, device=dev).to(dt); wgt = (0.04 * torch.randn(C, C, 3, 3, device=dev)).to(dt)
    go = torch.randn(B, C, H, W, device=dev).to(dt)
    res[f"cols_total_{dt}_ms"] = timed(lambda: ops._dcn_bwd_data_cols(go, x, off, m, wgt, True, True, True, f32_math=dt == torch.float32), n=3)
print(json.dumps(res))
