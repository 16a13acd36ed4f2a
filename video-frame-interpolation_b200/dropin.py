"""Drop-in installation behind the reference's own seams.

Seam 1: the bound method ``EMA_VFI.warp(self, frame2, feature, flow)``
        (/root/reference/src/models/ema_vfi.py:149, sole call site :130) -- replaced on the class.
Seam 2: the module-level function ``torchvision.ops.deform_conv.deform_conv2d`` which ``DeformConv2d.forward``
        resolves as a global at call time (torchvision/ops/deform_conv.py:170) -- replaced in that module, so the
        ``DeformConv2d`` class, its Parameters, ``reset_parameters`` and the state-dict keys
        ``attention_blocks.N.dcn_v2.{weight,bias}`` stay untouched (checkpoints load, seeded init is identical).

The reference's source files are not edited; ``inference.py`` / ``train.py`` run unmodified through
``python -m vfi_b200.run <script> ...``.
"""
from __future__ import annotations

import sys
from typing import Optional

from . import ops

_state = {"installed": False, "warp_classes": [], "tv_orig": None, "math": "auto", "calls": {"warp": 0, "dcn": 0}}


def _find_model_class():
    for name in ("src.models.ema_vfi", "models.ema_vfi", "ema_vfi"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "EMA_VFI"):
            return mod.EMA_VFI
    return None


def _warp_method(self, frame2, feature, flow):
    """Replacement for EMA_VFI.warp: same signature; ``feature`` is unused by the reference beyond ``.is_cuda``."""
    _state["calls"]["warp"] += 1
    return ops.warp(frame2, flow)


def _deform_conv2d(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None):
    """Replacement for torchvision.ops.deform_conv.deform_conv2d (same signature and argument meaning)."""
    _state["calls"]["dcn"] += 1
    return ops.deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask,
                             math=_state["math"])


def install(model_cls=None, *, math: str = "auto", patch_torchvision: bool = True) -> None:
    """Route the reference model's warp and DeformConv2d calls to libvfi_b200.

    ``model_cls``: the reference's ``EMA_VFI`` class (or any class exposing the same ``warp`` seam).  When omitted,
    an already-imported ``src.models.ema_vfi`` is looked up in ``sys.modules``.
    ``math``: ``"auto"`` (fp32 tensors -> fp32 parity kernels, bf16/fp16 -> tcgen05), ``"fp32"`` or ``"bf16_tc"``.
    """
    from . import _lib

    _lib.load()  # fail now, loudly, if the CUDA library has not been built
    _state["math"] = math
    cls = model_cls or _find_model_class()
    if cls is not None and not any(c is cls for c, _ in _state["warp_classes"]):
        _state["warp_classes"].append((cls, cls.__dict__.get("warp")))
        cls.warp = _warp_method
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
    _state["warp_classes"].clear()
    if _state["tv_orig"] is not None:
        import torchvision.ops
        import torchvision.ops.deform_conv as tv

        tv.deform_conv2d, torchvision.ops.deform_conv2d = _state["tv_orig"]
        _state["tv_orig"] = None
    _state["installed"] = False


def installed() -> bool:
    return _state["installed"]


def call_counts() -> dict:
    """How many times each seam has been hit since import (3 DCN + 1 warp per EMA_VFI.forward)."""
    return dict(_state["calls"])
