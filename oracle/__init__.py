"""CPU oracle for the warp + DeformConv2d hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  The product package never does; it fails loudly without its CUDA library.

Two oracles live here:

* :mod:`oracle.c_oracle` -- ctypes binding of ``vfi_oracle.c``, the plain-C restatement of the reference
  algorithm (``/root/reference/src/models/ema_vfi.py:149-171`` and ``:53-60`` plus the torch / torchvision op
  bodies restated from SURVEY.md Appendix A/B).  Pinned against ``tests/golden/*.npz``.
* :mod:`oracle.torch_ref` -- the same call sequence the reference makes, expressed with the stock
  ``torch.nn.functional.grid_sample`` / ``torchvision.ops.deform_conv2d`` CPU kernels (the third-party ops the
  reference itself runs).  Used as the timed CPU baseline and as a large-size cross-check on the GPU box.
"""
from .c_oracle import (  # noqa: F401
    build, lib_path, warp_fwd, warp_bwd, warp_blend_fwd, dcn_fwd, dcn_bwd, pack_split, threads, set_threads,
)
