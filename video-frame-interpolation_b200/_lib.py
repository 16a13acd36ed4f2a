"""ctypes binding of libvfi_b200.so -- the stub a maintainer of the reference would add to call the C ABI
declared in include/vfi_b200.h.  No torch types cross this boundary: raw device pointers,
extents, strides, a dtype enum and the CUDA stream handle.

There is no fallback: if the shared object is missing, or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

import torch

import os

# VFI_B200_LIB selects another build of the same library (kernel tuning experiments); the default is the in-tree one
LIB_PATH = Path(os.environ.get("VFI_B200_LIB") or Path(__file__).resolve().parent / "libvfi_b200.so")

VFI_OK = 0
ERR_NAMES = {1: "VFI_ERR_INVALID", 2: "VFI_ERR_UNSUPPORTED", 3: "VFI_ERR_CUDA", 4: "VFI_ERR_WORKSPACE", 5: "VFI_ERR_DEVICE"}
F32, BF16, F16 = 0, 1, 2
MATH_AUTO, MATH_FP32, MATH_BF16_TC, MATH_BF16_TC_HQ = 0, 1, 2, 3
WARP_DIV_IEEE, WARP_DIV_RECIPROCAL = 0, 1
WARP_OUT_TAIL_RECORD, WARP_NO_STAGING, WARP_COUNT_TILES = 2, 4, 8
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class VfiTensor(ctypes.Structure):
    _fields_ = [("data", c_void_p), ("dtype", c_int32), ("reserved", c_int32),
                ("n", c_int64), ("c", c_int64), ("h", c_int64), ("w", c_int64),
                ("sn", c_int64), ("sc", c_int64), ("sh", c_int64), ("sw", c_int64)]


_T = POINTER(VfiTensor)
# name -> (restype, argtypes): exactly the declarations of include/vfi_b200.h (tests/test_abi.py cross-checks them)
SIGNATURES = {
    "vfi_abi_version": (c_int, []),
    "vfi_version_string": (c_char_p, []),
    "vfi_last_error": (c_char_p, []),
    "vfi_check_device": (c_int, []),
    "vfi_launch_count": (c_int64, []),
    "vfi_reset_launch_count": (None, []),
    "vfi_warp_fwd": (c_int, [_T, _T, _T, c_int32, c_void_p]),
    "vfi_warp_tile_counts": (c_int, [POINTER(c_uint64), POINTER(c_uint64), c_int32]),
    "vfi_warp_bwd": (c_int, [_T, _T, _T, _T, _T, c_int32, c_void_p]),
    "vfi_warp_blend_fwd": (c_int, [_T, _T, _T, _T, _T, _T, c_int32, c_void_p]),
    "vfi_dcn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int64, c_int32]),
    "vfi_dcn_packed_weight_bytes": (c_size_t, []),
    "vfi_dcn_pack_weight": (c_int, [c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "vfi_dcn_k_order": (c_int, [c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "vfi_dcn_pack_input": (c_int, [_T, c_void_p, c_void_p, c_void_p]),
    "vfi_dcn_fwd": (c_int, [_T, _T, _T, c_void_p, c_int32, c_void_p, c_int32, _T, c_int64, c_int32, c_void_p, c_size_t, c_void_p]),
    "vfi_dcn_fwd_fused": (c_int, [_T, _T, _T, c_void_p, c_int32, c_void_p, c_int32, _T, _T, c_int64, c_int32, c_void_p, c_size_t, c_void_p]),
    "vfi_dcn_bwd_data": (c_int, [_T, _T, _T, _T, c_void_p, c_int32, c_int64, _T, _T, _T, c_void_p, c_size_t, c_void_p]),
    "vfi_dcn_bwd_weight": (c_int, [_T, _T, _T, _T, c_int64, c_void_p, c_void_p, c_void_p]),
    "vfi_dcn_gcol_workspace_bytes": (c_size_t, []),
    "vfi_dcn_gcol": (c_int, [_T, c_void_p, c_int32, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "vfi_dcn_bwd_data_cols_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int32]),
    "vfi_dcn_bwd_data_cols": (c_int, [c_void_p, c_int32, c_int64, _T, _T, _T, c_void_p, c_int64, _T, _T, c_void_p, c_size_t,
                              c_void_p]),
    "vfi_dcn_bwd_weight_tc_fused": (c_int, [_T, _T, _T, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vfi_dcn_bwd_data_cols_fused": (c_int, [c_void_p, c_int64, _T, _T, _T, c_void_p, c_int64, _T, c_void_p]),
    "vfi_dcn_bwd_weight_tc": (c_int, [_T, _T, _T, _T, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vfi_selftest_umma": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "vfi_selftest_umma_ts": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vfi_debug_read": (c_int, [c_void_p, c_size_t]),
    "vfi_debug_abort_info": (c_int, [c_void_p]),
}

_lib = None
_device_ok = set()


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built: there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "vfi_b200 has no CPU or PyTorch fallback for the warp / DeformConv2d path.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.vfi_abi_version() != 1:
            raise RuntimeError(f"libvfi_b200.so ABI {lib.vfi_abi_version()} != 1 expected by the Python host")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().vfi_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "libvfi_b200") -> None:
    if rc != VFI_OK:
        msg = last_error()
        if rc == 2:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {ERR_NAMES.get(rc, rc)}: {msg}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on one CUDA device, and that device must be sm_100.  No fallback."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("vfi_b200 ops need CUDA tensors on a B200 (sm_100a); there is no CPU fallback "
                               f"(got a tensor on {t.device})")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"vfi_b200: tensors on different devices ({dev} vs {t.device})")
    if dev is None:
        raise RuntimeError("vfi_b200: no tensor given")
    if dev.index not in _device_ok:
        with torch.cuda.device(dev):
            check(load().vfi_check_device(), "vfi_check_device")
        _device_ok.add(dev.index)
    return dev


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise NotImplementedError(f"vfi_b200: dtype {dt} is not supported (float32, bfloat16, float16)") from None


def desc(t: torch.Tensor) -> VfiTensor:
    """Describe a 4-D torch tensor (any strides) as a vfi_tensor."""
    if t.dim() != 4:
        raise ValueError(f"expected a 4-D tensor, got {tuple(t.shape)}")
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    return VfiTensor(t.data_ptr(), dtype_code(t.dtype), 0, n, c, h, w, sn, sc, sh, sw)


def ref(d):
    return None if d is None else ctypes.byref(d)


def stream_handle(dev: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)
