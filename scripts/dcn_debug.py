"""Wait-cycle breakdown of the staged DCN kernel (VFI_DCN_DEBUG=1): where each warp role spends its time."""
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
src.main.copy_(feat.permute(0, 2, 3, 1)); 
for _ in range(3):
    y = ops.deform_conv2d_fused(src.main_nchw, src.tail_nchw(3), convs[0], ws[0], bs[0])
torch.cuda.synchronize()
buf = np.zeros(256 * 32 * 8, dtype=np.uint64)
_lib.check(_lib.load().vfi_debug_read(buf.ctypes.data_as(ctypes.c_void_p), buf.size))
d = buf.reshape(256, 32, 8)[:148].astype(np.float64)
def show(name, warps, labels):
    x = d[:, warps, :]
    tot = x[..., 0].mean()
    print(f"{name:10s} total {tot/1e3:8.1f} kcyc  " + "  ".join(f"{l} {x[..., i+1].mean()/tot*100:5.1f}%" for i, l in enumerate(labels)))
show("producers", list(range(20)), ["geo_full", "box_full", "a_empty"])
show("mma", [20], ["acc_empty", "b_full", "a_full", "issue"])
print("   mma commit %5.1f%%" % (d[:, 20, 6].mean() / d[:, 20, 0].mean() * 100))
show("box", [21], ["box_empty"])
show("bload", [22], ["b_empty"])
show("epilogue", list(range(24, 28)), ["-", "acc_full", "-", "epilogue"])
show("geometry", list(range(28, 32)), ["geo_empty", "-", "geometry"])
print("tiles per CTA", d[:, 0, 5].mean(), " cycles per tile", d[:, 0, 0].mean() / d[:, 0, 5].mean())
