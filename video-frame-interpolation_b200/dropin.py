"""Drop-in installation behind the reference's own seams.

Seam 1: the bound method ``EMA_VFI.warp(self, frame2, feature, flow)``
        (/root/reference/src/models/ema_vfi.py:149, sole call site :130) -- replaced on the class.
Seam 2: the module-level function ``torchvision.ops.deform_conv.deform_conv2d`` which ``DeformConv2d.forward``
        resolves as a global at call time (torchvision/ops/deform_conv.py:170) -- replaced in that module, so the
        ``DeformConv2d`` class, its Parameters, ``reset_parameters`` and the state-dict keys
        ``attention_blocks.N.dcn_v2.{weight,bias}`` stay untouched (checkpoints load, seeded init is identical).

``fuse=True`` (inference under CUDA autocast or with 16-bit activations) additionally patches,
again on classes and module globals only:

Seam 3: ``ModulatedDeformConvPack.forward`` (ema_vfi.py:53-60) -- the 27-channel ``offset_conv`` output goes straight to
        ``vfi_dcn_fwd_fused`` (no chunk / cat / sigmoid passes), activations travel between the three blocks as ONE
        ``[B,H,W,72]`` bf16 buffer (64 main channels + the 16-byte tail record) that the tcgen05 kernel reads and writes
        through TMA tensor maps, and ``offset_conv`` itself runs as a stock cuDNN convolution on that buffer as it lies.
Seam 4: the ``torch.cat`` of ema_vfi.py:134 -- the model module's global name ``torch`` is replaced by a proxy whose ``cat``
        recognises ``[feat, warped]`` (the warp has already written its three channels into the tail records of a fresh
        buffer) and fills in the 64 feature channels instead of concatenating.

The reference's source files are not edited; ``inference.py`` / ``train.py`` run unmodified through
``python -m vfi_b200.run <script> ...``.
"""
from __future__ import annotations

import sys
import types
from typing import Optional

import torch

from . import ops

_state = {"installed": False, "warp_classes": [], "tv_orig": None, "math": "auto", "division": "reciprocal", "fuse": False,
          "pack_classes": [], "torch_proxies": [], "calls": {"warp": 0, "dcn": 0, "fused_block": 0, "fused_cat": 0, "fused_warp": 0}}
_DEFAULT_PACK_NAMES = ("ModulatedDeformConvPack", "FusionPack")


def _find_model_class():
    for name in ("src.models.ema_vfi", "models.ema_vfi", "ema_vfi"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "EMA_VFI"):
            return mod.EMA_VFI
    return None


# ------------------------------------------------------------------------------------------------------ fused inference
def _fusable(t: torch.Tensor) -> bool:
    """Fused (tensor-core, bf16) route: CUDA inference where the reference itself computes in 16 bits -- under an active CUDA
    autocast (inference.py:158-159) or on 16-bit activations -- and nothing needs a gradient."""
    return (_state["fuse"] and t.is_cuda and not torch.is_grad_enabled()
            and (torch.is_autocast_enabled() or t.dtype in (torch.bfloat16, torch.float16)))


def _new_buffer(B: int, H: int, W: int, device) -> torch.Tensor:
    return torch.empty((B, H, W, ops.MAIN_C + ops.TAIL_C), dtype=torch.bfloat16, device=device)


def _tag(view: torch.Tensor, buf: torch.Tensor) -> torch.Tensor:
    view._vfi_buf = buf          # plain attribute on the tensor object the model passes along
    return view


def _buffer_of(x: torch.Tensor) -> Optional[torch.Tensor]:
    buf = getattr(x, "_vfi_buf", None)
    if buf is not None and x.data_ptr() == buf.data_ptr() and x.shape[1] == 67 and x.dtype == torch.bfloat16:
        return buf
    return None


def _to_buffer(x: torch.Tensor) -> torch.Tensor:
    """Any [B,67,H,W] tensor -> a fresh [B,H,W,72] buffer (one pass; only when the activation did not come from this module)."""
    B, C, H, W = x.shape
    buf = _new_buffer(B, H, W, x.device)
    buf[..., :C] = x.permute(0, 2, 3, 1)
    buf[..., C:68] = 0
    buf[..., 68:68 + (C - 64)] = x[:, 64:].permute(0, 2, 3, 1)      # the tail record's mirrored half
    buf[..., 68 + (C - 64):] = 0
    return buf


def _padded_offset_conv(pack):
    """offset_conv's parameters for the 72-channel buffer: zero weights for channels 67..71 (pad + mirrored tail), bf16,
    channels-last.  Cached on the module, refreshed when the parameters change."""
    w, b = pack.offset_conv.weight, pack.offset_conv.bias
    key = (w.data_ptr(), w._version, None if b is None else b._version, w.device)
    cached = getattr(pack, "_vfi_w72", None)
    if cached is None or cached[0] != key:
        w72 = torch.zeros((w.shape[0], ops.MAIN_C + ops.TAIL_C, 3, 3), dtype=torch.bfloat16, device=w.device)
        w72[:, : w.shape[1]] = w.detach()
        cached = (key, w72.contiguous(memory_format=torch.channels_last), None if b is None else b.detach().to(torch.bfloat16))
        pack._vfi_w72 = cached
    return cached[1], cached[2]


def _fused_pack_forward(orig):
    def forward(self, x):
        dcn = getattr(self, "dcn_v2", None)
        conv = getattr(self, "offset_conv", None)
        if (dcn is None or conv is None or not _fusable(x) or x.dim() != 4 or x.shape[1] != 67 or tuple(dcn.weight.shape) != (67, 67, 3, 3)
                or conv.kernel_size != (3, 3) or conv.stride != (1, 1) or conv.padding != (1, 1) or conv.dilation != (1, 1)
                or x.shape[3] % 8 != 0):
            return orig(self, x)
        _state["calls"]["fused_block"] += 1
        buf = _buffer_of(x)
        if buf is None:
            buf = _to_buffer(x)
        B, H, W, _ = buf.shape
        x72 = buf.permute(0, 3, 1, 2)                                  # dense channels_last [B,72,H,W]: cuDNN takes it as it lies
        w72, b27 = _padded_offset_conv(self)
        with torch.autocast("cuda", enabled=False):
            conv27 = torch.nn.functional.conv2d(x72, w72, b27, stride=1, padding=1)   # [B,27,H,W] channels_last bf16
        out = ops.Planes.from_buffer72(_new_buffer(B, H, W, buf.device))
        ops.deform_conv2d_fused(x72[:, :ops.MAIN_C], x72[:, ops.MAIN_C:67], conv27, dcn.weight, dcn.bias, math="bf16_tc", out=out)
        return _tag(out.nchw_view(), out.buffer)

    forward._vfi_orig = orig
    return forward


class _TorchProxy(types.ModuleType):
    """Stands in for the name ``torch`` inside the model's module: everything resolves to the real package except ``cat``,
    which fills the feature channels of the buffer the warp has already written into (ema_vfi.py:134)."""

    def __init__(self, real):
        super().__init__("torch")
        self.__dict__["_real"] = real

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    def cat(self, tensors, dim=0, **kw):
        real = self.__dict__["_real"]
        if (not kw and dim == 1 and isinstance(tensors, (list, tuple)) and len(tensors) == 2
                and getattr(tensors[1], "_vfi_buf", None) is not None and tensors[1].shape[1] == 3
                and tensors[0].dim() == 4 and tensors[0].shape[1] == ops.MAIN_C and _fusable(tensors[0])
                and tensors[0].shape[0] == tensors[1].shape[0] and tensors[0].shape[2:] == tensors[1].shape[2:]):
            _state["calls"]["fused_cat"] += 1
            buf = tensors[1]._vfi_buf
            buf[..., : ops.MAIN_C].copy_(tensors[0].permute(0, 2, 3, 1))     # the cat's only traffic: 64 channels in, bf16 out
            return _tag(buf.permute(0, 3, 1, 2)[:, :67], buf)
        return real.cat(tensors, dim, **kw)


def _warp_method(self, frame2, feature, flow):
    """Replacement for EMA_VFI.warp: same signature; ``feature`` is unused by the reference beyond ``.is_cuda``."""
    _state["calls"]["warp"] += 1
    if _fusable(frame2) and frame2.dim() == 4 and frame2.shape[1] == 3 and frame2.shape[3] % 8 == 0:
        # the three warped channels land in the tail records of the buffer block 1 gathers from
        _state["calls"]["fused_warp"] += 1
        B, _, H, W = frame2.shape
        buf = _new_buffer(B, H, W, frame2.device)
        view = buf.permute(0, 3, 1, 2)[:, ops.MAIN_C:67]
        ops.warp(frame2.to(torch.bfloat16), flow, division=_state["division"], out=view, tail_record=True)
        return _tag(view, buf)
    return ops.warp(frame2, flow, division=_state["division"])


def _deform_conv2d(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None):
    """Replacement for torchvision.ops.deform_conv.deform_conv2d (same signature and argument meaning)."""
    _state["calls"]["dcn"] += 1
    math = _state["math"]
    if math == "auto" and input.is_cuda and torch.is_autocast_enabled() and not torch.is_grad_enabled():
        # Under CUDA autocast the activations reach this call in fp32 (the cat promotes) while the reference's own arithmetic
        # around it is 16-bit (inference.py:158-159): take the tensor cores instead of the fp32 parity kernel (108 ms / layer).
        math = "bf16_tc"
    return ops.deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask, math=math)


def install(model_cls=None, *, math: str = "auto", division: str = "reciprocal", fuse: bool = False, pack_cls=None,
            patch_torchvision: bool = True) -> None:
    """Route the reference model's warp and DeformConv2d calls to libvfi_b200.

    ``model_cls``: the reference's ``EMA_VFI`` class (or any class exposing the same ``warp`` seam).  When omitted,
    an already-imported ``src.models.ema_vfi`` is looked up in ``sys.modules``.
    ``math``: ``"auto"`` (fp32 tensors -> fp32 parity kernels, bf16/fp16 tensors or inference under CUDA autocast -> tcgen05),
    ``"fp32"`` or ``"bf16_tc"``.
    ``division``: which bit-level meaning of ``2.0 * v / (size - 1)`` the warp replays -- ``"reciprocal"`` (default: what the
    unmodified model computes on a GPU, where aten multiplies by the fp32 reciprocal) or ``"ieee"`` (what it computes on CPU:
    BASELINE config 1 and the golden vectors).
    ``fuse``: also patch seams 3 and 4 (module docstring).  ``pack_cls``: the block class when it cannot be found by name
    (``ModulatedDeformConvPack`` / ``FusionPack``) in ``model_cls``'s module.
    """
    from . import _lib

    _lib.load()  # fail now, loudly, if the CUDA library has not been built
    if division not in ("ieee", "reciprocal"):
        raise ValueError("division must be 'ieee' or 'reciprocal'")
    _state["math"], _state["division"], _state["fuse"] = math, division, bool(fuse)
    cls = model_cls or _find_model_class()
    if cls is not None and not any(c is cls for c, _ in _state["warp_classes"]):
        _state["warp_classes"].append((cls, cls.__dict__.get("warp")))
        cls.warp = _warp_method
    if fuse and cls is not None:
        # the module that defines the model's forward (the class handed in, or the base class it inherits everything from)
        mods = [m for m in (sys.modules.get(k.__module__) for k in cls.__mro__) if m is not None]
        packs = [pack_cls] if pack_cls is not None else [getattr(m, n) for m in mods for n in _DEFAULT_PACK_NAMES if hasattr(m, n)]
        if not packs:
            raise RuntimeError("install(fuse=True): no ModulatedDeformConvPack-like class found; pass pack_cls=")
        for pk in packs:
            if not any(c is pk for c, _ in _state["pack_classes"]):
                _state["pack_classes"].append((pk, pk.__dict__.get("forward")))
                pk.forward = _fused_pack_forward(pk.forward)
        for mod in mods:
            if (any(hasattr(mod, n) for n in _DEFAULT_PACK_NAMES) and isinstance(getattr(mod, "torch", None), types.ModuleType)
                    and not isinstance(mod.torch, _TorchProxy)):
                _state["torch_proxies"].append((mod, mod.torch))
                mod.torch = _TorchProxy(mod.torch)
    if patch_torchvision and _state["tv_orig"] is None:
        import torchvision.ops
        import torchvision.ops.deform_conv as tv

        _state["tv_orig"] = (tv.deform_conv2d, torchvision.ops.deform_conv2d)
        tv.deform_conv2d = _deform_conv2d
        torchvision.ops.deform_conv2d = _deform_conv2d
    _state["installed"] = True


def uninstall() -> None:
    for cls, orig in _state["warp_classes"]:
        if orig is not None:
            cls.warp = orig
        elif "warp" in cls.__dict__:
            delattr(cls, "warp")             # the class inherited its warp: remove our override
    _state["warp_classes"].clear()
    for pk, orig in _state["pack_classes"]:
        if orig is not None:
            pk.forward = orig
        elif "forward" in pk.__dict__:
            delattr(pk, "forward")
    _state["pack_classes"].clear()
    for mod, real in _state["torch_proxies"]:
        mod.torch = real
    _state["torch_proxies"].clear()
    if _state["tv_orig"] is not None:
        import torchvision.ops
        import torchvision.ops.deform_conv as tv

        tv.deform_conv2d, torchvision.ops.deform_conv2d = _state["tv_orig"]
        _state["tv_orig"] = None
    _state["installed"] = False
    _state["fuse"] = False


def installed() -> bool:
    return _state["installed"]


def call_counts() -> dict:
    """How many times each seam has been hit since import (3 DCN + 1 warp per EMA_VFI.forward; in fused mode 3 fused_block,
    1 fused_cat and 1 fused_warp instead of the 3 DCN)."""
    return dict(_state["calls"])
