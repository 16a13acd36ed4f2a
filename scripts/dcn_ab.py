"""A/B of the tcgen05 DCN forward variants in fresh processes (VFI_DCN_KERNEL / VFI_DCN_RAW / VFI_DCN_NO_TMA_STORE are read once
per process): checks that v7 reproduces v6 bit for bit on several shapes / layouts and times both at cfg2 size.
usage: dcn_ab.py [--time] [--out FILE]"""
import json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, torch
sys.path.insert(0, %r)
from vfi_b200 import ops
mode, out_path = sys.argv[1], sys.argv[2]
res = {}
def case(B, H, W, sigma, cl, seed):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    src = ops.Planes(B, H, W, 'cuda', zero_tail=True)
    src.set_tail(torch.randn(B, 3, H, W, generator=g))
    c27 = torch.randn(B, 27, H, W, generator=g)
    c27[:, :9] *= sigma
    c27[:, 18:] *= sigma
    c27 = c27.to(torch.bfloat16).cuda()
    if cl:
        c27 = c27.contiguous(memory_format=torch.channels_last)
    w = ((torch.rand(67, 67, 3, 3, generator=g) * 2 - 1) / 603 ** 0.5).to(torch.bfloat16).cuda()
    b = (torch.randn(67, generator=g) * 0.01).to(torch.bfloat16).cuda()
    y = ops.deform_conv2d_fused(feat, src.tail_nchw(3), c27, w, b)
    y2 = ops.deform_conv2d_fused(y.main_nchw, y.tail_nchw(), c27, w, b)      # planes in (second layer)
    off = torch.cat((c27[:, :9], c27[:, 18:]), 1).contiguous()
    m = torch.sigmoid(c27[:, 9:18]).contiguous()
    x67 = torch.cat((feat, src.tail_nchw(3)), 1).contiguous(memory_format=torch.channels_last)
    y3 = ops.deform_conv2d(x67, off, w, b, stride=1, padding=1, dilation=1, mask=m, math="bf16_tc")   # drop-in form
    torch.cuda.synchronize()
    return [y.to_nchw().cpu(), y2.to_nchw().cpu(), y3.cpu()]
if mode == "check":
    outs = {}
    for name, args in {"a_8x16": (1, 8, 16, 1.5, False, 1), "b_44x88_s6": (2, 44, 88, 6.0, False, 2), "c_64x96_cl": (1, 64, 96, 1.5, True, 3),
                       "d_37x56": (2, 37, 56, 3.0, False, 4), "e_270x480": (1, 270, 480, 1.5, False, 5),
                       "f_40x72_s0": (1, 40, 72, 0.0, False, 6)}.items():
        outs[name] = case(*args)
    torch.save(outs, out_path)
else:
    from vfi_b200.hotpath import HotPath, synthetic_inputs, synthetic_weights
    B, H, W = 8, 1080, 1920
    f2, fl, ft, convs = synthetic_inputs(B, H, W, device='cuda')
    ft = ft.contiguous(memory_format=torch.channels_last)
    ws, bs = synthetic_weights(device='cuda')
    path = HotPath(ws, bs, math='bf16_tc')
    for _ in range(3):
        y = path.run(f2, fl, ft, convs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = path.run(f2, fl, ft, convs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    json.dump({"ms_per_step": ms, "checksum": float(y.main.float().sum()) + float(y.tail.float().sum())}, open(out_path, "w"))
""" % ROOT


def run(env, mode, path):
    r = subprocess.run([sys.executable, "-c", CHILD, mode, path], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print("FAILED", env, r.stderr[-3000:])
    return r.returncode == 0


def main():
    import torch
    variants = {"v6": {"VFI_DCN_KERNEL": "v6"}, "v7": {}, "v7_ldg": {"VFI_DCN_RAW": "2"}, "v7_nostore": {"VFI_DCN_NO_TMA_STORE": "1"}}
    report = {"check": {}, "time": {}}
    with tempfile.TemporaryDirectory() as d:
        ok = {k: run(env, "check", os.path.join(d, k + ".pt")) for k, env in variants.items()}
        if ok.get("v6"):
            ref = torch.load(os.path.join(d, "v6.pt"))
            for k in variants:
                if k == "v6" or not ok[k]:
                    continue
                got = torch.load(os.path.join(d, k + ".pt"))
                for name in ref:
                    for i, (a, b) in enumerate(zip(ref[name], got[name])):
                        a, b = a.float(), b.float()
                        report["check"][f"{k}/{name}/{i}"] = {"equal": bool(torch.equal(a, b)), "maxdiff": float((a - b).abs().max()),
                                                             "frac_diff": float((a != b).float().mean()), "finite": bool(torch.isfinite(b).all())}
        report["ran"] = ok
        if "--time" in sys.argv:
            for k, env in variants.items():
                p = os.path.join(d, k + ".json")
                if run(env, "time", p):
                    report["time"][k] = json.load(open(p))
    bad = [k for k, v in report["check"].items() if not v["equal"]]
    report["mismatches"] = bad
    txt = json.dumps(report, indent=1)
    print(txt)
    if "--out" in sys.argv:
        open(sys.argv[sys.argv.index("--out") + 1], "w").write(txt)
    sys.exit(1 if bad or not all(ok.values()) else 0)


if __name__ == "__main__":
    main()
