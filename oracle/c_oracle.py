"""ctypes binding of oracle/vfi_oracle.c (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "vfi_oracle.c"
_OUT = _HERE / "_build" / "libvfi_oracle.so"
_lib = None


def lib_path() -> Path:
    return _OUT


def build(force: bool = False) -> Path:
    """Compile vfi_oracle.c with gcc (same flags as oracle/Makefile)."""
    if force or not _OUT.exists() or _OUT.stat().st_mtime < _SRC.stat().st_mtime:
        _OUT.parent.mkdir(exist_ok=True)
        cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
               str(_SRC), "-o", str(_OUT), "-lm"]
        subprocess.run(cmd, check=True)
    return _OUT


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_OUT))
        _lib.vfi_oracle_threads.restype = ctypes.c_int
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def threads() -> int:
    return int(_load().vfi_oracle_threads())


def set_threads(n: int) -> None:
    _load().vfi_oracle_set_threads(ctypes.c_int(n))


def warp_fwd(src, flow) -> np.ndarray:
    src, flow = _f32(src), _f32(flow)
    B, C, H, W = src.shape
    assert flow.shape == (B, 2, H, W)
    out = np.empty_like(src)
    _load().vfi_oracle_warp_fwd(_p(src), _p(flow), _p(out), B, C, H, W)
    return out


def warp_bwd(grad_out, src, flow, need_grad_src: bool = False):
    grad_out, src, flow = _f32(grad_out), _f32(src), _f32(flow)
    B, C, H, W = src.shape
    gflow = np.empty_like(flow)
    gsrc = np.zeros_like(src) if need_grad_src else None
    _load().vfi_oracle_warp_bwd(_p(grad_out), _p(src), _p(flow), _p(gflow), _p(gsrc), B, C, H, W)
    return (gflow, gsrc) if need_grad_src else gflow


def warp_blend_fwd(src_a, flow_a, src_b, flow_b, m) -> np.ndarray:
    src_a, flow_a, src_b, flow_b, m = map(_f32, (src_a, flow_a, src_b, flow_b, m))
    B, C, H, W = src_a.shape
    assert m.shape == (B, 1, H, W)
    out = np.empty_like(src_a)
    _load().vfi_oracle_warp_blend_fwd(_p(src_a), _p(flow_a), _p(src_b), _p(flow_b), _p(m), _p(out), B, C, H, W)
    return out


def dcn_fwd(x, offset, mask, weight, bias=None) -> np.ndarray:
    x, offset, mask, weight = map(_f32, (x, offset, mask, weight))
    bias = None if bias is None else _f32(bias)
    B, C, H, W = x.shape
    O = weight.shape[0]
    assert weight.shape == (O, C, 3, 3) and offset.shape == (B, 18, H, W) and mask.shape == (B, 9, H, W)
    out = np.empty((B, O, H, W), np.float32)
    _load().vfi_oracle_dcn_fwd(_p(x), _p(offset), _p(mask), _p(weight), _p(bias), _p(out), B, C, O, H, W)
    return out


def dcn_bwd(grad_out, x, offset, mask, weight):
    """Returns (grad_x, grad_offset, grad_mask, grad_weight, grad_bias)."""
    grad_out, x, offset, mask, weight = map(_f32, (grad_out, x, offset, mask, weight))
    B, C, H, W = x.shape
    O = weight.shape[0]
    gx = np.zeros_like(x)
    goff = np.empty_like(offset)
    gmask = np.empty_like(mask)
    gw = np.empty_like(weight)
    gb = np.empty((O,), np.float32)
    _load().vfi_oracle_dcn_bwd(_p(grad_out), _p(x), _p(offset), _p(mask), _p(weight), _p(gx), _p(goff), _p(gmask),
                               _p(gw), _p(gb), B, C, O, H, W)
    return gx, goff, gmask, gw, gb


def pack_split(conv27):
    conv27 = _f32(conv27)
    B, _, H, W = conv27.shape
    off = np.empty((B, 18, H, W), np.float32)
    msk = np.empty((B, 9, H, W), np.float32)
    _load().vfi_oracle_pack_split(_p(conv27), _p(off), _p(msk), B, H, W)
    return off, msk
