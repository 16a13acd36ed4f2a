"""Per-kernel timing of the DCNv2 training step at BASELINE config 3 (16 x 67 x 256 x 256): forward, backward-data and
backward-weight, fp32 CUDA-core kernels and bf16 tensor-core kernels, CUDA events on the current stream."""
import json, sys, importlib
import torch
sys.path.insert(0, ".")
pkg = importlib.import_module("video-frame-interpolation_b200")
ops = pkg.ops


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    B, C, H, W = 16, 67, 256, 256
    res = {"shape": [B, C, H, W]}
    for name, dt, math in (("fp32", torch.float32, "fp32"), ("bf16_tc", torch.bfloat16, "bf16_tc")):
        x = torch.randn(B, C, H, W, device=dev, generator=g).to(dt)
        off = (1.5 * torch.randn(B, 18, H, W, device=dev, generator=g)).to(dt)
        msk = torch.sigmoid(torch.randn(B, 9, H, W, device=dev, generator=g)).to(dt)
        w = (0.04 * torch.randn(C, C, 3, 3, device=dev, generator=g)).to(dt)
        bias = torch.zeros(C, device=dev, dtype=dt)
        go = torch.randn(B, C, H, W, device=dev, generator=g).to(dt)

        def run(req_data, req_w):
            xs = x.clone().requires_grad_(req_data); os_ = off.clone().requires_grad_(req_data)
            ms = msk.clone().requires_grad_(req_data); ws = w.clone().requires_grad_(req_w); bs = bias.clone().requires_grad_(req_w)
            out = ops.deform_conv2d(xs, os_, ws, bs, stride=1, padding=1, dilation=1, mask=ms, math=math)
            return out, go.to(memory_format=torch.channels_last) if out.is_contiguous(memory_format=torch.channels_last) else go

        with torch.no_grad():
            res[name + "_fwd_ms"] = timed(lambda: ops.deform_conv2d(x, off, w, bias, stride=1, padding=1, dilation=1, mask=msk, math=math))
        for tag, rd, rw in (("bwd_data", True, False), ("bwd_weight", False, True)):
            out, gg = run(rd, rw)
            res[f"{name}_{tag}_ms"] = timed(lambda: out.backward(gg, retain_graph=True), n=3)
    # the tensor-core weight gradient alone, grad_out NCHW-contiguous vs channels_last
    xb, ob, mb = x, off, msk
    for tag, gg in (("nchw", go.contiguous()), ("channels_last", go.contiguous(memory_format=torch.channels_last))):
        res[f"wgrad_tc_{tag}_ms"] = timed(lambda: ops.dcn_weight_grad_tc(gg, xb, ob, mb, C))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
