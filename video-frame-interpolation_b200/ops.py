"""Autograd operators of the hot path, implemented by libvfi_b200.so (hand-written sm_100a CUDA).

* :func:`warp`            -- ``EMA_VFI.warp`` (/root/reference/src/models/ema_vfi.py:149-171)
* :func:`warp_blend`      -- north-star extension (no reference counterpart)
* :func:`deform_conv2d`   -- ``torchvision.ops.deform_conv2d`` for the reference geometry (ema_vfi.py:45-60)

Every op raises when its inputs are not on a B200: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import collections
import ctypes
import warnings
from typing import Optional

import torch

from . import _lib
from ._lib import check, desc, dtype_code, ref, require_cuda, stream_handle

__all__ = ["warp", "warp_blend", "deform_conv2d", "deform_conv2d_fused", "Planes", "dcn_workspace_bytes", "launch_count", "reset_launch_count",
           "fallback_counts", "reset_fallback_counts"]


# Slower-but-correct routes the operators may take silently are counted here and announced once each (warnings.warn), so
# that a shape or layout change that costs performance is visible: `fallback_counts()` after a run, or -W error in tests.
_FALLBACKS: "collections.Counter[str]" = collections.Counter()
_WARNED = set()


def _note(kind: str, detail: str) -> None:
    _FALLBACKS[kind] += 1
    if kind not in _WARNED:
        _WARNED.add(kind)
        warnings.warn(f"vfi_b200: {kind}: {detail} (counted in vfi_b200.ops.fallback_counts(); reported once)", RuntimeWarning,
                      stacklevel=3)


def fallback_counts() -> dict:
    """How often each slower route was taken since import / the last :func:`reset_fallback_counts`."""
    return dict(_FALLBACKS)


def reset_fallback_counts() -> None:
    _FALLBACKS.clear()
    _WARNED.clear()


def launch_count() -> int:
    return int(_lib.load().vfi_launch_count())


def reset_launch_count() -> None:
    _lib.load().vfi_reset_launch_count()


# ------------------------------------------------------------------------------------------------------ warp
class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src: torch.Tensor, flow: torch.Tensor, flags: int = 0, out=None) -> torch.Tensor:
        dev = require_cuda(src, flow, out)
        ctx.flags = flags
        if src.dim() != 4 or flow.dim() != 4:
            raise ValueError("warp expects src [B,C,H,W] and flow [B,2,H,W]")
        if flow.dtype != torch.float32 and flow.dtype != src.dtype:
            flow = flow.float()
        if out is None:
            out = torch.empty_like(src)   # keeps src's memory format (NCHW or channels_last)
        elif out.shape != src.shape or out.dtype != src.dtype:
            raise ValueError("warp: out must have src's shape and dtype")
        else:
            ctx.mark_dirty(out)
        with torch.cuda.device(dev):
            check(_lib.load().vfi_warp_fwd(ref(desc(src)), ref(desc(flow)), ref(desc(out)), flags, stream_handle(dev)),
                  "vfi_warp_fwd")
        ctx.save_for_backward(src, flow)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        src, flow = ctx.saved_tensors
        dev = src.device
        need_src, need_flow = ctx.needs_input_grad[:2]
        if grad_out.dtype != src.dtype:
            grad_out = grad_out.to(src.dtype)
        gflow = torch.empty(flow.shape, dtype=torch.float32, device=dev)
        gsrc = torch.zeros(src.shape, dtype=torch.float32, device=dev) if need_src else None
        with torch.cuda.device(dev):
            check(_lib.load().vfi_warp_bwd(ref(desc(grad_out)), ref(desc(src)), ref(desc(flow)), ref(desc(gflow)),
                                           ref(desc(gsrc)) if gsrc is not None else None, ctx.flags,
                                           stream_handle(dev)), "vfi_warp_bwd")
        return (gsrc.to(src.dtype) if need_src else None), (gflow.to(flow.dtype) if need_flow else None), None, None


_DIVISION = {"ieee": _lib.WARP_DIV_IEEE, "reciprocal": _lib.WARP_DIV_RECIPROCAL}


def warp(src: torch.Tensor, flow: torch.Tensor, division: str = "ieee", out: Optional[torch.Tensor] = None,
         tail_record: bool = False, staging: bool = True, count_tiles: bool = False) -> torch.Tensor:
    """Backward-warp ``src`` [B,C,H,W] by ``flow`` [B,2,H,W] (pixels; channel 0 = x, 1 = y).

    Same result as the reference's grid build + normalise + ``F.grid_sample(bilinear, zeros,
    align_corners=True)`` (fp32: max-abs 1e-5), in one kernel and without materialising the grid.

    ``division`` selects which bit-level meaning of the reference's ``2.0 * v / (size-1)`` is replayed: ``"ieee"``
    (what the reference computes on CPU -- BASELINE config 1 and the golden vectors) or ``"reciprocal"`` (what aten
    computes for ``tensor / python_scalar`` on CUDA: a multiplication by the fp32 reciprocal).
    ``out`` (optional, any strides) receives the result in place: exactly the C channels it describes are written.
    ``tail_record=True`` (VFI_WARP_OUT_TAIL_RECORD): ``out`` is the [B,3,H,W] view of a DCN tail plane
    (:meth:`Planes.tail_nchw`) and the kernel writes whole 16-byte records ``[c0 c1 c2 0 | c0 c1 c2 0]`` -- elements 3..7
    of every pixel are overwritten.  This is what removes the ``torch.cat`` of ema_vfi.py:134.
    ``staging=False`` (VFI_WARP_NO_STAGING) forces round 1's L1-gather kernel instead of the TMA-staged one (same results;
    A/B runs and parity tests); ``count_tiles=True`` makes the launch count its staged / L1 tiles (:func:`warp_tile_counts`).
    """
    flags = _DIVISION[division] | (_lib.WARP_OUT_TAIL_RECORD if tail_record else 0)
    flags |= (0 if staging else _lib.WARP_NO_STAGING) | (_lib.WARP_COUNT_TILES if count_tiles else 0)
    return _WarpFn.apply(src, flow, flags, out)


def warp_tile_counts(reset: bool = True):
    """(tiles served from the staged shared-memory window, tiles gathered through L1) over the ``count_tiles=True`` launches on
    the current device since the last reset.  Synchronises."""
    import ctypes

    a, b = ctypes.c_uint64(), ctypes.c_uint64()
    check(_lib.load().vfi_warp_tile_counts(ctypes.byref(a), ctypes.byref(b), 1 if reset else 0), "vfi_warp_tile_counts")
    return int(a.value), int(b.value)


def warp_blend(src_a, flow_a, src_b, flow_b, m, division: str = "ieee") -> torch.Tensor:
    """``m * warp(src_a, flow_a) + (1 - m) * warp(src_b, flow_b)`` in one pass (forward only; m is [B,1,H,W])."""
    dev = require_cuda(src_a, flow_a, src_b, flow_b, m)
    if flow_a.dtype != torch.float32 and flow_a.dtype != src_a.dtype:
        flow_a = flow_a.float()
    flow_b = flow_b.to(flow_a.dtype)
    m = m.to(src_a.dtype)
    out = torch.empty_like(src_a, memory_format=torch.contiguous_format)
    with torch.cuda.device(dev):
        check(_lib.load().vfi_warp_blend_fwd(ref(desc(src_a)), ref(desc(flow_a)), ref(desc(src_b)), ref(desc(flow_b)),
                                             ref(desc(m)), ref(desc(out)), _DIVISION[division], stream_handle(dev)),
              "vfi_warp_blend_fwd")
    return out


# ------------------------------------------------------------------------------------------------------ DCN
_MATH = {"auto": _lib.MATH_AUTO, "fp32": _lib.MATH_FP32, "bf16_tc": _lib.MATH_BF16_TC}
MAIN_C, TAIL_C = 64, 8   # channels per pixel in the two activation planes of the tensor-core path


class Planes:
    """An activation of up to 72 channels in the layout the tcgen05 DCN kernel gathers from and writes: two dense
    channels-last bf16 buffers, ``main`` [B,H,W,64] (128 B per pixel) and ``tail`` [B,H,W,8] (16 B per pixel, channels
    64.. and zero padding; with at most four tail channels the upper half of the record mirrors the lower half, which
    the kernels that write planes do themselves -- use :meth:`set_tail` when filling one by hand).  ``main_nchw`` /
    ``tail_nchw(c)`` are the logical [B,C,H,W] views handed to the C ABI."""

    def __init__(self, B: int, H: int, W: int, device, channels: int = 67, zero_tail: bool = False):
        self.channels = channels
        self.main = torch.empty((B, H, W, MAIN_C), dtype=torch.bfloat16, device=device)
        alloc = torch.zeros if zero_tail else torch.empty
        self.tail = alloc((B, H, W, TAIL_C), dtype=torch.bfloat16, device=device)

    @classmethod
    def from_buffer72(cls, buf: torch.Tensor, channels: int = 67) -> "Planes":
        """View one dense channels-last activation buffer ``[B,H,W,72]`` (bf16) as planes: main = channels 0..63, tail record
        = channels 64..71 of every pixel (pixel stride 144 bytes for both).  The logical [B,channels,H,W] tensor is then the
        plain strided view ``buf.permute(0, 3, 1, 2)[:, :channels]`` -- what the fused drop-in hands between the reference's
        blocks and on to its stock convolutions, with no layout pass in between."""
        if buf.dim() != 4 or buf.shape[-1] != MAIN_C + TAIL_C or buf.dtype != torch.bfloat16 or not buf.is_contiguous():
            raise ValueError("from_buffer72 expects a contiguous bf16 [B,H,W,72] tensor")
        self = cls.__new__(cls)
        self.channels = channels
        self.main = buf[..., :MAIN_C]
        self.tail = buf[..., MAIN_C:]
        self.buffer = buf
        return self

    def nchw_view(self) -> torch.Tensor:
        """Zero-copy logical tensor of planes made by :meth:`from_buffer72`."""
        return self.buffer.permute(0, 3, 1, 2)[:, : self.channels]

    def set_tail(self, t: torch.Tensor) -> None:
        """Fill the tail plane from a [B,c,H,W] tensor (c <= 8): channels, zero padding and, for c <= 4, the mirrored half."""
        c = t.shape[1]
        v = t.permute(0, 2, 3, 1).to(device=self.tail.device, dtype=torch.bfloat16)
        self.tail.zero_()
        self.tail[..., :c] = v
        if c <= 4:
            self.tail[..., 4:4 + c] = v

    @property
    def main_nchw(self) -> torch.Tensor:
        return self.main.permute(0, 3, 1, 2)

    def tail_nchw(self, c: Optional[int] = None) -> torch.Tensor:
        return self.tail.permute(0, 3, 1, 2)[:, : (self.channels - MAIN_C if c is None else c)]

    def to_nchw(self) -> torch.Tensor:
        """Logical [B,channels,H,W] tensor (a copy; for checks and for handing the result to stock PyTorch ops)."""
        return torch.cat((self.main_nchw, self.tail_nchw()), dim=1)


def selftest_umma(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Diagnostic: D = a @ b.T for bf16 a [128,K], b [80,K] through the library's tcgen05 plumbing (one CTA)."""
    dev = require_cuda(a, b)
    assert a.dtype == b.dtype == torch.bfloat16 and a.shape[0] == 128 and b.shape[0] == 80 and a.shape[1] == b.shape[1]
    a, b = a.contiguous(), b.contiguous()
    d = torch.empty((128, 80), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().vfi_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), a.shape[1], stream_handle(dev)),
              "vfi_selftest_umma")
    return d


def selftest_umma_ts(a: torch.Tensor, b: torch.Tensor):
    """Diagnostic: (2 * a @ b.T, raw TMEM image of a) for bf16 a [128,64], b [80,64] with the A operand staged in tensor
    memory by tcgen05.st and consumed by the A-from-TMEM MMA form (the plumbing of the v6 DCN kernel)."""
    dev = require_cuda(a, b)
    assert a.dtype == b.dtype == torch.bfloat16 and tuple(a.shape) == (128, 64) and tuple(b.shape) == (80, 64)
    a, b = a.contiguous(), b.contiguous()
    d = torch.empty((128, 80), dtype=torch.float32, device=dev)
    raw = torch.zeros((128, 32), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().vfi_selftest_umma_ts(a.data_ptr(), b.data_ptr(), d.data_ptr(), raw.data_ptr(), stream_handle(dev)),
              "vfi_selftest_umma_ts")
    return d, raw


def dcn_weight_grad_tc(grad_out: torch.Tensor, x: torch.Tensor, offset: torch.Tensor, mask: torch.Tensor, O: int):
    """grad_weight [O,C,3,3] and grad_bias [O] (fp32) of the DCNv2 layer on the tensor cores (vfi_dcn_bwd_weight_tc):
    bf16 operands, fp32 accumulation in tensor memory.  Raises NotImplementedError for shapes / dtypes it does not take."""
    dev = require_cuda(grad_out, x, offset, mask)
    B, C, H, W = x.shape
    lib = _lib.load()
    gw = torch.zeros((O, C, 3, 3), dtype=torch.float32, device=dev)
    gb = torch.zeros((O,), dtype=torch.float32, device=dev)
    ws = _workspace(dev, int(lib.vfi_dcn_workspace_bytes(B, C, O, H, W, _lib.MATH_BF16_TC)))
    with torch.cuda.device(dev):
        check(lib.vfi_dcn_bwd_weight_tc(ref(desc(grad_out)), ref(desc(x)), ref(desc(offset)), ref(desc(mask)), O, gw.data_ptr(),
                                        gb.data_ptr(), ws.data_ptr(), ws.numel(), stream_handle(dev)), "vfi_dcn_bwd_weight_tc")
    return gw, gb


def dcn_workspace_bytes(B: int, C: int, O: int, H: int, W: int, math: str = "auto") -> int:
    return int(_lib.load().vfi_dcn_workspace_bytes(B, C, O, H, W, _MATH[math]))


def _workspace(dev, nbytes: int) -> torch.Tensor:
    # Caller-owned workspace (the library never allocates): one uint8 tensor from the caching allocator.
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)


_COLS_LD = 9 * (MAIN_C + TAIL_C)          # gcol columns: 9 taps x 72 channels
_GX_LD = MAIN_C + 4                       # fp32 grad_x accumulator row
_COLS_CHUNK_BYTES = 1 << 30               # target size of one materialised column-gradient block; blocks are whole images, so one
                                          # image is the floor: 2.7 GB (bf16) / 5.4 GB (fp32) at 1080p, 4x that at 4K


def cols_weight_matrix(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[72, 648] matrix M with ``M[o, k*72 + c] = weight[o, c, k // 3, k % 3]`` (zero rows / columns beyond O / C): the
    right-hand side of the column-gradient GEMM ``gcol = grad_out_rows @ M`` in the column order
    ``vfi_dcn_bwd_data_cols`` reads (include/vfi_b200.h)."""
    O, C = weight.shape[:2]
    TAP = MAIN_C + TAIL_C
    wt = torch.zeros((TAP, 9, TAP), dtype=dtype, device=weight.device)
    wt[:O, :, :C] = weight.detach().reshape(O, C, 9).permute(0, 2, 1)
    return wt.view(TAP, _COLS_LD)


def _dcn_bwd_data_cols(grad_out, x, offset, mask, weight, need_x, need_off, need_mask, f32_math: bool):
    """grad_x / grad_offset / grad_mask in column-gradient form (torchvision::_deform_conv2d_backward, reference call
    site src/models/ema_vfi.py:60): the dense half ``gcol = grad_out x W`` is a plain GEMM (cuBLAS through torch.matmul),
    materialised one batch chunk at a time; everything that depends on the sampling positions is
    ``vfi_dcn_bwd_data_cols``.  ``f32_math``: fp32 operands and true fp32 GEMM arithmetic (TF32 off) for the parity path,
    else bf16 operands with fp32 accumulation (1.3 KB per pixel).  Returns fp32 tensors (grad_x as a channels-last view)."""
    dev = x.device
    B, C, H, W = x.shape
    O = weight.shape[0]
    lib = _lib.load()
    TAP = MAIN_C + TAIL_C
    cdt = torch.float32 if f32_math else torch.bfloat16
    f32 = dict(dtype=torch.float32, device=dev)
    wt = cols_weight_matrix(weight, cdt) if (f32_math or O > 68) else None
    hw = H * W
    gx_rows = torch.zeros((B * hw, _GX_LD), **f32) if need_x else None
    goff = torch.empty(offset.shape, **f32) if need_off else None
    gmask = torch.empty(mask.shape, **f32) if need_mask else None
    step = max(1, _COLS_CHUNK_BYTES // max(1, hw * _COLS_LD * (4 if f32_math else 2)))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    try:
        if f32_math:
            torch.backends.cuda.matmul.allow_tf32 = False
        with torch.cuda.device(dev):
            for b0 in range(0, B, step):
                b1 = min(B, b0 + step)
                n = (b1 - b0) * hw
                if n == 0:
                    continue
                if not f32_math and O <= 68:
                    # tensor-core training path: the column gradient on tcgen05, grad_out read where it lies (vfi_dcn_gcol)
                    gcol = torch.empty((n, _COLS_LD), dtype=cdt, device=dev)
                    wsg = _workspace(dev, int(lib.vfi_dcn_gcol_workspace_bytes()))
                    check(lib.vfi_dcn_gcol(ref(desc(grad_out[b0:b1])), weight.data_ptr(), dtype_code(weight.dtype), C, gcol.data_ptr(),
                                           _COLS_LD, wsg.data_ptr(), wsg.numel(), stream_handle(dev)), "vfi_dcn_gcol")
                else:
                    # fp32 parity path: true fp32 GEMM arithmetic through the BLAS (TF32 off)
                    g72 = torch.zeros((n, TAP), dtype=cdt, device=dev)
                    g72[:, :O] = grad_out[b0:b1].permute(0, 2, 3, 1).reshape(n, O)
                    gcol = torch.matmul(g72, wt)
                ws = _workspace(dev, int(lib.vfi_dcn_bwd_data_cols_workspace_bytes(b1 - b0, H, W, dtype_code(cdt))))
                check(lib.vfi_dcn_bwd_data_cols(gcol.data_ptr(), dtype_code(cdt), _COLS_LD, ref(desc(x[b0:b1])),
                                                ref(desc(offset[b0:b1])), ref(desc(mask[b0:b1])),
                                                gx_rows[b0 * hw:].data_ptr() if need_x else None, _GX_LD,
                                                ref(desc(goff[b0:b1])) if need_off else None,
                                                ref(desc(gmask[b0:b1])) if need_mask else None,
                                                ws.data_ptr(), ws.numel(), stream_handle(dev)), "vfi_dcn_bwd_data_cols")
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    gx = gx_rows.view(B, H, W, _GX_LD)[..., :C].permute(0, 3, 1, 2) if need_x else None
    return gx, goff, gmask


class _DcnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, math: int):
        dev = require_cuda(x, offset, mask, weight, bias)
        B, C, H, W = x.shape
        O = weight.shape[0]
        ctx.offset_dtype, ctx.mask_dtype = offset.dtype, mask.dtype       # gradients go back in the callers' dtypes
        if offset.dtype != mask.dtype:
            mask = mask.to(offset.dtype)
        weight_c = weight.contiguous()
        bias_c = None if bias is None else bias.contiguous()
        lib = _lib.load()
        tc = math == _lib.MATH_BF16_TC or (math == _lib.MATH_AUTO and x.dtype != torch.float32)
        if tc and offset.dtype == torch.float32:
            # The staged-box kernels (forward v6, weight gradient) read 16-bit offsets / masks.  fp16 keeps a sampling position
            # to 2^-11 of the offset (0.004 px at 8 px) -- far inside what rounding the activations to bf16 costs -- and is what
            # the reference's own autocast path hands over.
            _note("offset_fp16_rounding", "fp32 offsets / mask rounded to fp16 for the tensor-core kernels (math='bf16_tc' on fp32 "
                  "offset tensors; 2^-11 relative on the sampling position)")
            offset = offset.clamp(-30000.0, 30000.0).half()
            mask = mask.half()
        if tc and offset.dtype != torch.float32:
            # the staged-box kernels (forward v6, weight gradient) stream offset / mask rows with bulk copies: unit pixel
            # stride.  A channels_last offset_conv output (layers 2 and 3 under channels_last training) costs one 54 B/px copy
            # here instead of the slower generic kernels.
            if offset.stride(3) != 1 or mask.stride(3) != 1:
                _note("offset_layout_copy", "offset / mask without unit pixel stride (channels_last) copied to NCHW for the "
                      "tensor-core kernels: one 54 B/px pass")
            if offset.stride(3) != 1:
                offset = offset.contiguous()
            if mask.stride(3) != 1:
                mask = mask.contiguous()
        # tensor-core results of low-precision inputs come back channels_last (what cuDNN wants next); fp32 stays NCHW
        fmt = torch.channels_last if (tc and x.dtype != torch.float32) else torch.contiguous_format
        out = torch.empty((B, O, H, W), dtype=x.dtype, device=dev, memory_format=fmt)
        nbytes = int(lib.vfi_dcn_workspace_bytes(B, C, O, H, W, math))
        ws = _workspace(dev, nbytes)
        with torch.cuda.device(dev):
            check(lib.vfi_dcn_fwd(ref(desc(x)), ref(desc(offset)), ref(desc(mask)), weight_c.data_ptr(),
                                  dtype_code(weight_c.dtype), None if bias_c is None else bias_c.data_ptr(),
                                  dtype_code(bias_c.dtype) if bias_c is not None else 0, ref(desc(out)), O, math,
                                  ws.data_ptr(), ws.numel(), stream_handle(dev)), "vfi_dcn_fwd")
        ctx.save_for_backward(x, offset, mask, weight_c)
        ctx.tc = tc
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, offset, mask, weight = ctx.saved_tensors
        dev = x.device
        need_x, need_off, need_mask, need_w, need_b, _ = ctx.needs_input_grad
        need_b = need_b and ctx.has_bias
        B, C, H, W = x.shape
        O = weight.shape[0]
        if grad_out.dtype != x.dtype:
            grad_out = grad_out.to(x.dtype)
        lib = _lib.load()
        f32 = dict(dtype=torch.float32, device=dev)
        # column-gradient form for every channel count the planes hold; fp32 arithmetic unless the forward ran on the tensor cores
        cols = C <= MAIN_C + 4 and O <= MAIN_C + TAIL_C and B > 0 and H > 0 and W > 0
        gx = torch.zeros(x.shape, **f32) if need_x and not cols else None
        goff = torch.empty(offset.shape, **f32) if need_off and not cols else None
        gmask = torch.empty(mask.shape, **f32) if need_mask and not cols else None
        if cols and (need_x or need_off or need_mask):
            gx, goff, gmask = _dcn_bwd_data_cols(grad_out, x, offset, mask, weight, need_x, need_off, need_mask,
                                                  f32_math=not ctx.tc)
        with torch.cuda.device(dev):
            if (need_x or need_off or need_mask) and not cols:
                ws = _workspace(dev, int(lib.vfi_dcn_workspace_bytes(B, C, O, H, W, _lib.MATH_FP32)))
                check(lib.vfi_dcn_bwd_data(ref(desc(grad_out)), ref(desc(x)), ref(desc(offset)), ref(desc(mask)),
                                           weight.data_ptr(), dtype_code(weight.dtype), O,
                                           ref(desc(gx)) if need_x else None, ref(desc(goff)) if need_off else None,
                                           ref(desc(gmask)) if need_mask else None, ws.data_ptr(), ws.numel(),
                                           stream_handle(dev)), "vfi_dcn_bwd_data")
            gw = torch.zeros(weight.shape, **f32) if need_w else None
            gb = torch.zeros((O,), **f32) if need_b else None
            done = False
            if (need_w or need_b) and ctx.tc:
                # tensor-core forward -> tensor-core weight gradient (bf16 operands, fp32 accumulation); shapes it does not
                # take (VFI_ERR_UNSUPPORTED) use the fp32 CUDA-core kernel below
                ws = _workspace(dev, int(lib.vfi_dcn_workspace_bytes(B, C, O, H, W, _lib.MATH_BF16_TC)))
                # the kernel's grad_out^T loaders take any strides, but a channels_last grad_out costs them 128 two-byte
                # loads per thread and tile (measured 2.1 vs 0.7 ms per layer at config 3): one 134 B/px copy is cheaper
                if grad_out.stride(3) != 1:
                    _note("grad_out_layout_copy", "channels_last grad_out copied to NCHW for the tcgen05 weight-gradient kernel: "
                          "one 134 B/px pass")
                g_w = grad_out if grad_out.stride(3) == 1 else grad_out.contiguous()
                rc = lib.vfi_dcn_bwd_weight_tc(ref(desc(g_w)), ref(desc(x)), ref(desc(offset)), ref(desc(mask)), O,
                                               gw.data_ptr() if need_w else None, gb.data_ptr() if need_b else None,
                                               ws.data_ptr(), ws.numel(), stream_handle(dev))
                if rc == 0:
                    done = True
                elif rc != 2:
                    check(rc, "vfi_dcn_bwd_weight_tc")
                else:
                    _note("wgrad_cuda_core_fallback", "the tcgen05 weight-gradient kernel does not take this call ("
                          + _lib.last_error() + "); using the fp32 CUDA-core kernel (~13x slower at config 3)")
            if (need_w or need_b) and not done:
                check(lib.vfi_dcn_bwd_weight(ref(desc(grad_out)), ref(desc(x)), ref(desc(offset)), ref(desc(mask)), O,
                                             gw.data_ptr() if need_w else None, gb.data_ptr() if need_b else None,
                                             stream_handle(dev)), "vfi_dcn_bwd_weight")
        return (gx.to(x.dtype) if need_x else None, goff.to(ctx.offset_dtype) if need_off else None,
                gmask.to(ctx.mask_dtype) if need_mask else None, gw.to(weight.dtype) if need_w else None,
                gb.to(ctx.bias_dtype) if need_b else None, None)


def deform_conv2d_fused(x_main: torch.Tensor, x_tail: Optional[torch.Tensor], conv27: torch.Tensor, weight: torch.Tensor,
                        bias: Optional[torch.Tensor] = None, math: str = "bf16_tc", out: Optional[Planes] = None) -> Planes:
    """Hot-path form of the DCNv2 forward (inference only, tensor-core math): consumes the raw 27-channel
    ``offset_conv`` output (the chunk/cat/sigmoid of ema_vfi.py:57-59 happens inside the kernel) and the input as
    planes -- ``x_main`` [B,64,H,W] + ``x_tail`` [B,<=8,H,W], channels-last bf16, i.e. feat and the warped frame before
    the torch.cat of ema_vfi.py:134, or the :class:`Planes` a previous layer produced.  ``x_tail=None``: ``x_main`` is
    any [B,C,H,W] tensor.  Returns :class:`Planes` (what the next layer reads without a layout pass)."""
    dev = require_cuda(x_main, x_tail, conv27, weight, bias)
    if math not in ("auto", "bf16_tc"):
        raise NotImplementedError("deform_conv2d_fused implements the tensor-core math modes only")
    B, _, H, W = x_main.shape
    O = weight.shape[0]
    if not (MAIN_C < O <= MAIN_C + TAIL_C):
        raise NotImplementedError("deform_conv2d_fused writes planes: 64 < O <= 72")
    lib = _lib.load()
    weight_c = weight.contiguous()
    bias_c = None if bias is None else bias.contiguous()
    if out is None:
        out = Planes(B, H, W, dev, channels=O)
    ws = _workspace(dev, int(lib.vfi_dcn_workspace_bytes(B, x_main.shape[1], O, H, W, _MATH[math]))
                    if x_tail is None else int(lib.vfi_dcn_packed_weight_bytes()) + 512)
    with torch.cuda.device(dev):
        check(lib.vfi_dcn_fwd_fused(ref(desc(x_main)), ref(desc(x_tail)) if x_tail is not None else None,
                                    ref(desc(conv27)), weight_c.data_ptr(), dtype_code(weight_c.dtype),
                                    None if bias_c is None else bias_c.data_ptr(),
                                    dtype_code(bias_c.dtype) if bias_c is not None else 0, ref(desc(out.main_nchw)),
                                    ref(desc(out.tail_nchw(O - MAIN_C))), O, _MATH[math], ws.data_ptr(), ws.numel(),
                                    stream_handle(dev)), "vfi_dcn_fwd_fused")
    return out


# ------------------------------------------------------------------------------------------------------ fused training block
def records_buffer(B: int, H: int, W: int, device, zero: bool = False) -> torch.Tensor:
    """A 72-channel *record* activation: a [B,72,H,W] bf16 tensor whose memory is dense channels-last ([B,H,W,72], 144 bytes
    per pixel).  Channels 0..66 are data, 67 is zero and 68..71 mirror 64..67 (the tail record of :class:`Planes`).  Stock
    cuDNN convolutions take it as it lies (zero weights for channels 67..71); the DCN kernels read and write it in place."""
    alloc = torch.zeros if zero else torch.empty
    return alloc((B, H, W, MAIN_C + TAIL_C), dtype=torch.bfloat16, device=device).permute(0, 3, 1, 2)


def _is_records(t: torch.Tensor) -> bool:
    return (t.dim() == 4 and t.shape[1] == MAIN_C + TAIL_C and t.dtype == torch.bfloat16 and t.is_cuda
            and t.permute(0, 2, 3, 1).is_contiguous())


class _DcnBlockFn(torch.autograd.Function):
    """y72 = DCNv2(x72[:, :67]; split / sigmoid of conv27) as records -- forward ``vfi_dcn_fwd_fused``, backward
    ``vfi_dcn_gcol`` + ``vfi_dcn_bwd_data_cols_fused`` + ``vfi_dcn_bwd_weight_tc_fused``."""

    @staticmethod
    def forward(ctx, x72, conv27, weight, bias):
        dev = require_cuda(x72, conv27, weight, bias)
        B, _, H, W = x72.shape
        O, C = weight.shape[:2]
        src = Planes.from_buffer72(x72.permute(0, 2, 3, 1), channels=C)
        out72 = records_buffer(B, H, W, dev)
        out = Planes.from_buffer72(out72.permute(0, 2, 3, 1), channels=O)
        weight = weight.contiguous()
        deform_conv2d_fused(src.main_nchw, src.tail_nchw(), conv27, weight, bias, math="bf16_tc", out=out)
        ctx.save_for_backward(x72, conv27, weight)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return out72

    @staticmethod
    def backward(ctx, grad_y72):
        x72, conv27, weight = ctx.saved_tensors
        dev = x72.device
        need_x, need_c27, need_w, need_b = ctx.needs_input_grad
        need_b = need_b and ctx.has_bias
        B, _, H, W = x72.shape
        O, C = weight.shape[:2]
        lib = _lib.load()
        f32 = dict(dtype=torch.float32, device=dev)
        # channels 67..71 of a record are derived (zero pad + mirror): what reaches them from downstream is not a gradient of
        # anything (stock convolutions carry zero weights there, the next block returns zeros) -- only [:, :O] counts
        g_out = grad_y72[:, :O]
        if g_out.dtype != torch.bfloat16:
            g_out = g_out.to(torch.bfloat16)
        src = Planes.from_buffer72(x72.permute(0, 2, 3, 1), channels=C)
        gx72 = g27 = gw = gb = None
        hw = H * W
        with torch.cuda.device(dev):
            if need_x or need_c27:
                gx_rows = torch.zeros((B * hw, _GX_LD), **f32) if need_x else None
                g27 = torch.empty(conv27.shape, **f32).contiguous(memory_format=torch.channels_last) if need_c27 else None
                step = max(1, _COLS_CHUNK_BYTES // max(1, hw * _COLS_LD * 2))
                wsg = _workspace(dev, int(lib.vfi_dcn_gcol_workspace_bytes()))
                for b0 in range(0, B, step):
                    b1 = min(B, b0 + step)
                    n = (b1 - b0) * hw
                    gcol = torch.empty((n, _COLS_LD), dtype=torch.bfloat16, device=dev)
                    check(lib.vfi_dcn_gcol(ref(desc(g_out[b0:b1])), weight.data_ptr(), dtype_code(weight.dtype), C, gcol.data_ptr(),
                                           _COLS_LD, wsg.data_ptr(), wsg.numel(), stream_handle(dev)), "vfi_dcn_gcol")
                    check(lib.vfi_dcn_bwd_data_cols_fused(gcol.data_ptr(), _COLS_LD, ref(desc(src.main_nchw[b0:b1])),
                                                          ref(desc(src.tail_nchw()[b0:b1])), ref(desc(conv27[b0:b1])),
                                                          gx_rows[b0 * hw:].data_ptr() if need_x else None, _GX_LD,
                                                          ref(desc(g27[b0:b1])) if need_c27 else None, stream_handle(dev)),
                          "vfi_dcn_bwd_data_cols_fused")
                if need_x:
                    gx72 = records_buffer(B, H, W, dev, zero=True)
                    gx72[:, :C] = gx_rows.view(B, H, W, _GX_LD)[..., :C].permute(0, 3, 1, 2)      # one cast pass; 67.. stay zero
                if need_c27:
                    g27 = g27.to(conv27.dtype)
            if need_w or need_b:
                gw = torch.zeros(weight.shape, **f32) if need_w else None
                gb = torch.zeros((O,), **f32) if need_b else None
                ws = _workspace(dev, int(lib.vfi_dcn_workspace_bytes(B, C, O, H, W, _lib.MATH_BF16_TC)))
                # the kernel's grad_out^T loaders want unit pixel stride (see _DcnFn.backward): one 134 B/px copy
                g_w = g_out.contiguous()
                check(lib.vfi_dcn_bwd_weight_tc_fused(ref(desc(g_w)), ref(desc(src.nchw_view())), ref(desc(conv27)), O,
                                                      gw.data_ptr() if need_w else None, gb.data_ptr() if need_b else None,
                                                      ws.data_ptr(), ws.numel(), stream_handle(dev)), "vfi_dcn_bwd_weight_tc_fused")
                if need_w:
                    gw = gw.to(weight.dtype)                 # fp32 master weights may be passed as they are: no rounding then
                if need_b:
                    gb = gb.to(ctx.bias_dtype)
        return gx72, g27, gw, gb


def deform_conv2d_block(x72: torch.Tensor, conv27: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The fusion block's DCNv2 (ema_vfi.py:57-60) with its glue folded in, forward AND backward, on 72-channel record
    activations (:func:`records_buffer`): ``x72`` in, the raw 27-channel ``offset_conv`` output in (16-bit, NCHW or
    channels_last; chunk / cat / sigmoid happen in the kernels, and the gradient returns to this tensor), records out.
    bf16 tensor-core math; ``weight`` [O,C,3,3] with 64 < C, O <= 68, W % 8 == 0.  Differentiable in all four arguments."""
    if not _is_records(x72):
        raise ValueError("deform_conv2d_block: x72 must be a bf16 [B,72,H,W] record tensor (ops.records_buffer)")
    if conv27.dim() != 4 or conv27.shape[1] != 27 or conv27.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("deform_conv2d_block: conv27 must be a 16-bit [B,27,H,W] tensor")
    O, C = weight.shape[:2]
    if not (MAIN_C < C <= MAIN_C + 4 and MAIN_C < O <= MAIN_C + 4) or tuple(weight.shape[2:]) != (3, 3):
        raise NotImplementedError("deform_conv2d_block implements 64 < C, O <= 68 and 3x3 kernels")
    if x72.shape[3] % 8:
        raise NotImplementedError("deform_conv2d_block needs W % 8 == 0")
    return _DcnBlockFn.apply(x72, conv27, weight, bias)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def deform_conv2d(input: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                  stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask: Optional[torch.Tensor] = None,
                  math: str = "auto") -> torch.Tensor:
    """Drop-in for ``torchvision.ops.deform_conv2d`` (same positional/keyword arguments, same error for bad shapes)
    restricted to the geometry the reference uses: 3x3 kernel, stride 1, padding 1, dilation 1, one weight group,
    one offset group, modulation mask present.  Anything else raises ``NotImplementedError`` -- by design there is
    no fallback to torchvision.
    """
    if input.dim() != 4 or weight.dim() != 4 or offset.dim() != 4:
        raise ValueError("deform_conv2d expects 4-D input, offset and weight")
    kh, kw = weight.shape[-2:]
    geometry = (_pair(stride), _pair(padding), _pair(dilation), (kh, kw))
    if geometry != ((1, 1), (1, 1), (1, 1), (3, 3)):
        raise NotImplementedError(
            f"vfi_b200.deform_conv2d implements stride=1, padding=1, dilation=1, 3x3 only (got stride={stride}, "
            f"padding={padding}, dilation={dilation}, kernel={kh}x{kw})")
    if mask is None:
        raise NotImplementedError("vfi_b200.deform_conv2d implements the modulated (mask) form only (DCNv2)")
    if weight.shape[1] != input.shape[1]:
        raise NotImplementedError("vfi_b200.deform_conv2d implements groups=1 only")
    if offset.shape[1] != 2 * kh * kw or mask.shape[1] != kh * kw:
        raise NotImplementedError("vfi_b200.deform_conv2d implements one offset group only "
                                  f"(offset has {offset.shape[1]} channels, mask {mask.shape[1]})")
    if math not in _MATH:
        raise ValueError(f"math must be one of {sorted(_MATH)}")
    return _DcnFn.apply(input, offset, mask, weight, bias, _MATH[math])
