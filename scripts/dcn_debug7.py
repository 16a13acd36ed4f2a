"""Wait-cycle breakdown of the v7 DCN kernel (VFI_DCN_DEBUG=1): where each warp role spends its time."""
import os, sys
os.environ["VFI_DCN_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
import vfi_b200
from vfi_b200 import ops, _lib
from vfi_b200.hotpath import synthetic_inputs, synthetic_weights
B, H, W = 8, 1080, 1920
dev = "cuda"
frame2, flow, feat, convs = synthetic_inputs(B, H, W, device=dev)
ws, bs = synthetic_weights(device=dev)
src = ops.Planes(B, H, W, dev, zero_tail=True)
src.main.copy_(feat.permute(0, 2, 3, 1))
for _ in range(3):
    y = ops.deform_conv2d_fused(src.main_nchw, src.tail_nchw(3), convs[0], ws[0], bs[0])
torch.cuda.synchronize()
buf = np.zeros(256 * 32 * 8, dtype=np.uint64)
_lib.check(_lib.load().vfi_debug_read(buf.ctypes.data_as(ctypes.c_void_p), buf.size))
d = buf.reshape(256, 32, 8)[:148].astype(np.float64)
def show(name, warps, labels):
    x = d[:, warps, :]
    tot = x[..., 0].mean()
    idx = {0: 1, 1: 2, 2: 3, 3: 4, 4: 6, 5: 7}
    print(f"{name:10s} total {tot/1e3:8.1f} kcyc  " + "  ".join(f"{l} {x[..., idx[i]].mean()/tot*100:5.1f}%" for i, l in enumerate(labels) if l != "-"))
P = int(os.environ.get("V7_PARTS", "3"))
GEO, EPI, MMA, COPY, BLOAD, PROD = 0, 4 * P, 4 * P + 4, 4 * P + 5, 4 * P + 6, 4 * P + 8
show("producers", list(range(PROD, 32)), ["geo", "box_full", "a_empty"])
show("mma", [MMA], ["acc_empty", "tail_full", "a_full", "issue", "commit"])
show("copy", [COPY], ["box_empty", "raw_empty"])
show("bload", [BLOAD], ["b_empty"])
show("epilogue", list(range(EPI, EPI + 4)), ["-", "acc_full", "-", "epilogue"])
show("geometry", list(range(GEO, GEO + 4 * P)), ["geo_empty", "raw_full", "compute", "box+acc_wait", "raw_read", "tail"])
print("tiles per CTA", d[:, PROD, 5].mean(), " cycles per tile", d[:, PROD, 0].mean() / d[:, PROD, 5].mean())
