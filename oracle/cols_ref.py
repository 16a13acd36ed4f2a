"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the column-gradient split of
torchvision::_deform_conv2d_backward that vfi_dcn_bwd_data_cols implements on the GPU (include/vfi_b200.h;
reference call site /root/reference/src/models/ema_vfi.py:60 through autograd; arithmetic: SURVEY.md Appendix B).

    gcol[b, y, x, k, c] = sum_o grad_out[b, o, y, x] * weight[o, c, k]                      (dense half, a GEMM)
    D_j                 = <gcol[b, y, x, k, :], x[b, :, corner_j]>            (zero for corners outside the image)
    grad_mask[b, k]     = live ? sum_j w_j D_j : 0
    grad_offset[b, 2k]  = mask * (lw (D11 - D01) + hw (D10 - D00)),     [b, 2k + 1] = mask * (lh (D11 - D10) + hh (D01 - D00))
    grad_x[b, :, corner_j] += gcol * mask * w_j                                 (live samples, corners inside the image)

It is pinned against the C oracle (oracle.dcn_bwd, itself pinned on the reference's outputs) in tests/test_oracle.py."""
import numpy as np


def dcn_bwd_data_from_cols(gcol, x, offset, mask):
    """gcol [B,H,W,9,C]; x [B,C,H,W]; offset [B,18,H,W]; mask [B,9,H,W] -> (grad_x, grad_offset, grad_mask), float64."""
    gcol, x, offset, mask = (np.asarray(a, np.float64) for a in (gcol, x, offset, mask))
    B, C, H, W = x.shape
    gx = np.zeros((B, H, W, C))
    goff = np.zeros((B, 18, H, W))
    gmask = np.zeros((B, 9, H, W))
    xl = np.transpose(x, (0, 2, 3, 1))                              # channels-last view, as the kernel reads it
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    bi = np.arange(B)[:, None, None] * np.ones((1, H, W), np.int64)
    for k in range(9):
        i, j = divmod(k, 3)
        # positions in fp32, as every implementation computes them
        py = ((ys - 1 + i).astype(np.float32)[None] + offset[:, 2 * k].astype(np.float32)).astype(np.float64)
        px = ((xs - 1 + j).astype(np.float32)[None] + offset[:, 2 * k + 1].astype(np.float32)).astype(np.float64)
        live = (py > -1) & (py < H) & (px > -1) & (px < W)
        y0, x0 = np.floor(py), np.floor(px)
        lh, lw = py - y0, px - x0
        hh, hw = 1 - lh, 1 - lw
        y0, x0 = y0.astype(np.int64), x0.astype(np.int64)
        g = gcol[:, :, :, k, :]
        m = mask[:, k]
        D = []
        for yy, xx, w in ((y0, x0, hh * hw), (y0, x0 + 1, hh * lw), (y0 + 1, x0, lh * hw), (y0 + 1, x0 + 1, lh * lw)):
            valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            yc, xc = np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)
            v = xl[bi, yc, xc] * valid[..., None]
            D.append((g * v).sum(-1))
            contrib = g * (m * w * (live & valid))[..., None]
            np.add.at(gx, (bi, yc, xc), contrib)
        d00, d01, d10, d11 = D
        gmask[:, k] = live * (hh * hw * d00 + hh * lw * d01 + lh * hw * d10 + lh * lw * d11)
        goff[:, 2 * k] = m * (lw * (d11 - d01) + hw * (d10 - d00))
        goff[:, 2 * k + 1] = m * (lh * (d11 - d10) + hh * (d01 - d00))
    return np.transpose(gx, (0, 3, 1, 2)), goff, gmask
