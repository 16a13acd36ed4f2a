"""Launcher that runs the reference's scripts byte-for-byte unmodified on the B200 path.

    python -m vfi_b200.run [--math auto|fp32|bf16_tc] /path/to/reference/inference.py --input_video ...
    python -m vfi_b200.run /path/to/reference/train.py

It (1) puts the script's directory on ``sys.path`` and imports ``src.models.ema_vfi`` so the seams exist,
(2) installs the drop-in (dropin.install), (3) applies the compatibility shims the reference needs on the pinned
torch and (4) hands control to the script with ``runpy`` under ``__main__``.

Shims (none touches the reference's files):
* ``ReduceLROnPlateau(..., verbose=True)`` (train.py:84) -- ``verbose`` was removed in torch 2.7+; accepted and dropped.
* ``models.vgg16(pretrained=True)`` (src/utils/loss_functions.py:31-34) -- no network in the sandbox; when the
  weights cannot be fetched the constructor falls back to random weights and says so.
"""
from __future__ import annotations

import argparse
import importlib
import os
import runpy
import sys
import warnings


def apply_compat_shims() -> None:
    import torch

    sched = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(sched, "_vfi_shim", False):
        orig_init = sched.__init__

        def init(self, *a, verbose=None, **kw):  # noqa: ANN001
            return orig_init(self, *a, **kw)

        sched.__init__ = init
        sched._vfi_shim = True

    try:
        import torchvision.models as tvm
    except Exception:  # pragma: no cover
        return
    if not getattr(tvm, "_vfi_shim", False):
        orig_vgg16 = tvm.vgg16

        def vgg16(*a, pretrained=None, weights=None, **kw):  # noqa: ANN001
            try:
                if pretrained is not None:
                    return orig_vgg16(*a, pretrained=pretrained, **kw)
                return orig_vgg16(*a, weights=weights, **kw)
            except Exception as e:  # offline: URLError / OSError
                warnings.warn(f"vgg16 weights unavailable offline ({type(e).__name__}); using random weights")
                return orig_vgg16(*a, weights=None, **kw)

        tvm.vgg16 = vgg16
        tvm._vfi_shim = True


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(prog="python -m vfi_b200.run", description=__doc__.split("\n")[0])
    ap.add_argument("--math", default="auto", choices=["auto", "fp32", "bf16_tc"])
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)

    script = os.path.abspath(ns.script)
    sys.path.insert(0, os.path.dirname(script))
    sys.dont_write_bytecode = True
    apply_compat_shims()
    from . import dropin

    model_cls = None
    try:
        model_cls = importlib.import_module("src.models.ema_vfi").EMA_VFI
    except Exception as e:
        warnings.warn(f"could not import src.models.ema_vfi next to {script}: {e}; only the torchvision seam is patched")
    dropin.install(model_cls, math=ns.math)
    sys.argv = [script, *ns.args]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
