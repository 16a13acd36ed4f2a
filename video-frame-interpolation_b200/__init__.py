"""vfi_b200 -- B200-native (sm_100a) warp + DeformConv2d hot path of the EMA-VFI-style interpolator in
424635328/video-frame-interpolation.

    import vfi_b200
    vfi_b200.install()                        # patch EMA_VFI.warp and torchvision's deform_conv2d seam
    out = model(frame1, frame2)               # the reference model, unmodified, now runs the CUDA path

The directory is called ``video-frame-interpolation_b200`` (not importable as a Python name); the repo-level
``vfi_b200`` shim loads it under this name.
"""
from . import _lib  # noqa: F401
from ._build import build  # noqa: F401
from .dropin import install, installed, uninstall  # noqa: F401
from .hotpath import HotPath  # noqa: F401
from .ops import deform_conv2d, launch_count, reset_launch_count, warp, warp_blend  # noqa: F401

__version__ = "0.1.0"
