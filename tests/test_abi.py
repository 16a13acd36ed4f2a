"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/vfi_b200.h declares,
the Python binding declares the same set, the seams patch and restore, and nothing silently falls back on CPU."""
import ctypes
import os
import re
import sys
from pathlib import Path

import pytest
import torch

import vfi_b200
from vfi_b200 import _lib, dropin

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "vfi_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(vfi_[a-z0-9_]+)\s*\(", body)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vfi_b200.h but not exported"


def test_python_binding_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.vfi_abi_version() == int(re.search(r"VFI_B200_ABI_VERSION (\d+)", HEADER).group(1))
    assert b"sm_100a" in lib.vfi_version_string()
    assert lib.vfi_dcn_packed_weight_bytes() == 11 * 80 * 128
    assert lib.vfi_dcn_workspace_bytes(1, 67, 67, 8, 8, _lib.MATH_FP32) >= 2 * 9 * 72 * 72 * 4


def test_struct_layout_matches_header():
    # 8 (ptr) + 4 + 4 + 8 * 8
    assert ctypes.sizeof(_lib.VfiTensor) == 80
    t = torch.zeros(2, 3, 4, 5).to(memory_format=torch.channels_last)
    d = _lib.desc(t)
    assert (d.n, d.c, d.h, d.w) == (2, 3, 4, 5) and (d.sn, d.sc, d.sh, d.sw) == (60, 1, 15, 3)


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vfi_b200.warp(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vfi_b200.deform_conv2d(torch.zeros(1, 67, 4, 4), torch.zeros(1, 18, 4, 4), torch.zeros(67, 67, 3, 3),
                               torch.zeros(67), stride=1, padding=1, dilation=1, mask=torch.zeros(1, 9, 4, 4))


def test_the_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke()/build() and bench.py's CPU reference legs may touch it.
    Static check over the package sources (imports, importlib calls, ctypes loads of the oracle library) plus a dynamic one:
    importing every product module leaves no `oracle` module behind."""
    import re
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|import_module\(\s*[\"']oracle|libvfi_oracle|oracle/_build|oracle/_ref", re.M)
    for f in sorted((root / "video-frame-interpolation_b200").rglob("*")):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and "_obj" not in f.parts:
            assert not pat.search(f.read_text(errors="ignore")), f
    hits = [ln for ln in (root / "bench.py").read_text().splitlines() if re.search(r"\bfrom oracle\b|\bimport oracle\b", ln)]
    assert len(hits) == 1                                   # cpu_reference_step_factory: the cpu_baseline / --impl reference legs
    code = ("import sys; sys.path.insert(0, %r); import vfi_b200; "
            "from vfi_b200 import ops, dropin, hotpath, shard, stream, trainstep, refmodel, run, _lib, _build; "
            "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')], 'oracle imported by the product'" % str(root))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]


def test_geometry_outside_the_path_is_refused():
    x, w = torch.zeros(1, 67, 4, 4), torch.zeros(67, 67, 3, 3)
    off, m = torch.zeros(1, 18, 4, 4), torch.zeros(1, 9, 4, 4)
    with pytest.raises(NotImplementedError):
        vfi_b200.deform_conv2d(x, off, w, None, stride=2, padding=1, dilation=1, mask=m)
    with pytest.raises(NotImplementedError):
        vfi_b200.deform_conv2d(x, off, w, None, stride=1, padding=0, dilation=1, mask=m)
    with pytest.raises(NotImplementedError):
        vfi_b200.deform_conv2d(x, off, w, None, stride=1, padding=1, dilation=1, mask=None)
    with pytest.raises(NotImplementedError):
        vfi_b200.deform_conv2d(x, torch.zeros(1, 36, 4, 4), w, None, stride=1, padding=1, dilation=1, mask=m)
    with pytest.raises(NotImplementedError):
        vfi_b200.deform_conv2d(x, off, torch.zeros(67, 67, 5, 5), None, stride=1, padding=1, dilation=1, mask=m)


def test_abi_argument_validation_without_a_gpu():
    """Bad descriptors are rejected before any CUDA call, with a message."""
    lib = _lib.load()
    src = _lib.VfiTensor(1, 0, 0, 1, 3, 4, 4, 48, 16, 4, 1)      # fake non-null pointer, never dereferenced
    flow_bad = _lib.VfiTensor(1, 0, 0, 1, 3, 4, 4, 48, 16, 4, 1)  # 3 channels instead of 2
    rc = lib.vfi_warp_fwd(ctypes.byref(src), ctypes.byref(flow_bad), ctypes.byref(src), 0, None)
    assert rc == 1 and b"flow must be [B,2,H,W]" in lib.vfi_last_error()
    rc = lib.vfi_warp_fwd(None, None, None, 0, None)
    assert rc == 1 and b"null" in lib.vfi_last_error()
    x = _lib.VfiTensor(1, 0, 0, 1, 100, 4, 4, 1600, 16, 4, 1)
    off = _lib.VfiTensor(1, 0, 0, 1, 18, 4, 4, 288, 16, 4, 1)
    m = _lib.VfiTensor(1, 0, 0, 1, 9, 4, 4, 144, 16, 4, 1)
    out = _lib.VfiTensor(1, 0, 0, 1, 100, 4, 4, 1600, 16, 4, 1)
    rc = lib.vfi_dcn_fwd(ctypes.byref(x), ctypes.byref(off), ctypes.byref(m), 1, 0, None, 0, ctypes.byref(out), 100,
                         _lib.MATH_FP32, 1, 1 << 20, None)
    assert rc == 2 and b"at most 72" in lib.vfi_last_error()
    # the training entry points reject what they do not take before touching the device
    x67 = _lib.VfiTensor(256, 2, 0, 1, 67, 8, 16, 67 * 128, 128, 16, 1)           # bf16 [1,67,8,16]
    off16 = _lib.VfiTensor(256, 2, 0, 1, 18, 8, 16, 18 * 128, 128, 16, 1)
    m16 = _lib.VfiTensor(256, 2, 0, 1, 9, 8, 16, 9 * 128, 128, 16, 1)
    rc = lib.vfi_dcn_bwd_data_cols(256, 1, 640, ctypes.byref(x67), ctypes.byref(off16), ctypes.byref(m16), None, 68, None, None,
                                   None, 0, None)
    assert rc == 1 and b"gcol rows must hold 9 x 72" in lib.vfi_last_error()
    rc = lib.vfi_dcn_bwd_data_cols(256, 2, 648, ctypes.byref(x67), ctypes.byref(off16), ctypes.byref(m16), None, 68, None, None,
                                   None, 0, None)                            # f16 columns
    assert rc == 2 and b"gcol must be bf16 or f32" in lib.vfi_last_error()
    rc = lib.vfi_dcn_bwd_data_cols(256, 1, 648, ctypes.byref(x67), ctypes.byref(off16), ctypes.byref(m16), 256, 68,
                                   None, None, None, 0, None)
    assert rc != 0 and b"workspace" in lib.vfi_last_error()
    assert lib.vfi_dcn_bwd_data_cols_workspace_bytes(1, 8, 16, 0) >= 128 * 72 * 4 > lib.vfi_dcn_bwd_data_cols_workspace_bytes(1, 8, 16, 1) >= 128 * 72 * 2
    xf = _lib.VfiTensor(256, 2, 0, 1, 100, 8, 16, 100 * 128, 128, 16, 1)
    rc = lib.vfi_dcn_bwd_weight_tc(ctypes.byref(xf), ctypes.byref(xf), ctypes.byref(off16), ctypes.byref(m16), 100, None, None,
                                   None, 0, None)
    assert rc == 2 and b"supports C <=" in lib.vfi_last_error()


def test_fused_training_entry_points_validate_without_a_gpu():
    """vfi_dcn_bwd_data_cols_fused / vfi_dcn_bwd_weight_tc_fused (backward of vfi_dcn_fwd_fused) reject bad descriptors before any
    CUDA call; the record helpers of the Python layer refuse tensors that are not records."""
    import pytest

    from vfi_b200 import ops

    lib = _lib.load()
    # one [1,8,16,72] bf16 record buffer at a fake address: main = channels 0..63, tail = 64.., pixel stride 72 elements
    main = _lib.VfiTensor(256, 1, 0, 1, 64, 8, 16, 72 * 128, 1, 72 * 16, 72)
    tail = _lib.VfiTensor(256 + 128, 1, 0, 1, 3, 8, 16, 72 * 128, 1, 72 * 16, 72)
    c27 = _lib.VfiTensor(256, 1, 0, 1, 27, 8, 16, 27 * 128, 128, 16, 1)
    rc = lib.vfi_dcn_bwd_data_cols_fused(256, 640, ctypes.byref(main), ctypes.byref(tail), ctypes.byref(c27), None, 68, None, None)
    assert rc == 1 and b"gcol rows must hold 9 x 72" in lib.vfi_last_error()
    nchw_main = _lib.VfiTensor(256, 1, 0, 1, 64, 8, 16, 64 * 128, 128, 16, 1)                    # planar: not a plane view
    rc = lib.vfi_dcn_bwd_data_cols_fused(256, 648, ctypes.byref(nchw_main), ctypes.byref(tail), ctypes.byref(c27), None, 68, None, None)
    assert rc == 2 and b"x must be bf16 planes" in lib.vfi_last_error()
    c26 = _lib.VfiTensor(256, 1, 0, 1, 26, 8, 16, 26 * 128, 128, 16, 1)
    rc = lib.vfi_dcn_bwd_data_cols_fused(256, 648, ctypes.byref(main), ctypes.byref(tail), ctypes.byref(c26), None, 68, None, None)
    assert rc == 1 and b"conv27 must be a 16-bit [B,27,H,W]" in lib.vfi_last_error()
    assert lib.vfi_dcn_bwd_data_cols_fused(256, 648, ctypes.byref(main), ctypes.byref(tail), ctypes.byref(c27), None, 68, None, None) == 0  # nothing asked for
    x67 = _lib.VfiTensor(256, 1, 0, 1, 67, 8, 16, 72 * 128, 1, 72 * 16, 72)
    g67 = _lib.VfiTensor(256, 1, 0, 1, 67, 8, 16, 67 * 128, 128, 16, 1)
    c27f = _lib.VfiTensor(256, 0, 0, 1, 27, 8, 16, 27 * 128, 128, 16, 1)                          # f32 conv27
    rc = lib.vfi_dcn_bwd_weight_tc_fused(ctypes.byref(g67), ctypes.byref(x67), ctypes.byref(c27f), 67, 256, None, None, 0, None)
    assert rc == 2 and b"conv27 must be a 16-bit tensor" in lib.vfi_last_error()
    rc = lib.vfi_dcn_bwd_weight_tc_fused(ctypes.byref(g67), ctypes.byref(x67), ctypes.byref(c27), 67, 256, None, None, 0, None)
    assert rc != 0 and b"workspace" in lib.vfi_last_error()
    # Python layer
    r = ops.records_buffer(2, 4, 8, "cpu", zero=True)
    assert tuple(r.shape) == (2, 72, 4, 8) and r.stride() == (72 * 32, 1, 72 * 8, 72) and r.dtype == torch.bfloat16
    assert not ops._is_records(r)                                                  # records live on the GPU
    with pytest.raises(ValueError, match="record tensor"):
        ops.deform_conv2d_block(r, torch.zeros(2, 27, 4, 8, dtype=torch.bfloat16), torch.zeros(67, 67, 3, 3))


def test_dropin_patches_and_restores_both_seams():
    import torchvision.ops
    import torchvision.ops.deform_conv as tv

    sys.path.insert(0, str(ROOT))
    from oracle.torch_ref import WarpHost

    orig_tv, orig_warp = tv.deform_conv2d, WarpHost.warp
    dropin.install(WarpHost)
    try:
        assert dropin.installed()
        assert tv.deform_conv2d is not orig_tv and torchvision.ops.deform_conv2d is tv.deform_conv2d
        assert WarpHost.warp is not orig_warp
        # the patched seams reach the CUDA ops -- which refuse CPU tensors instead of falling back
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            WarpHost().warp(torch.zeros(1, 3, 4, 4), None, torch.zeros(1, 2, 4, 4))
        blk = torchvision.ops.DeformConv2d(67, 67, 3, padding=1)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            blk(torch.zeros(1, 67, 4, 4), torch.zeros(1, 18, 4, 4), torch.zeros(1, 9, 4, 4))
        assert sorted(blk.state_dict()) == ["bias", "weight"] and blk.weight.shape == (67, 67, 3, 3)
    finally:
        dropin.uninstall()
    assert tv.deform_conv2d is orig_tv and WarpHost.warp is orig_warp and not dropin.installed()


def test_fused_dropin_patches_and_restores_seams_3_and_4():
    """install(fuse=True): the block class's forward and the module-level name `torch` (the cat of ema_vfi.py:134) are patched
    on the class / module of the model handed in and restored by uninstall(); a class that only inherits `warp` gets its
    override removed again.  With CPU tensors (no autocast) the fused routes do not apply and the stock code runs."""
    import vfi_b200.refmodel as rm
    from vfi_b200.refmodel import FusionPack, StockInterpolator

    class Child(StockInterpolator):        # inherits warp
        pass

    orig_fwd, real_torch = FusionPack.forward, rm.torch
    dropin.install(Child, fuse=True, patch_torchvision=False)
    try:
        assert FusionPack.forward is not orig_fwd and FusionPack.forward._vfi_orig is orig_fwd
        assert rm.torch is not real_torch and rm.torch.nn is real_torch.nn and rm.torch.float32 is real_torch.float32
        assert "warp" in Child.__dict__
        a, b = torch.rand(1, 64, 4, 4), torch.rand(1, 3, 4, 4)
        assert torch.equal(rm.torch.cat((a, b), dim=1), real_torch.cat((a, b), dim=1))       # ordinary cats pass through
        pack = FusionPack(67).eval()
        x = torch.randn(1, 67, 8, 8)
        with torch.no_grad():
            assert torch.equal(pack(x), orig_fwd(pack, x))                                   # CPU / fp32: not fusable -> stock forward
    finally:
        dropin.uninstall()
    assert FusionPack.forward is orig_fwd and rm.torch is real_torch and "warp" not in Child.__dict__
    assert Child.warp is StockInterpolator.warp


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_fused_dropin_finds_the_reference_block_class():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        import src.models.ema_vfi as ref
    finally:
        sys.path.remove("/root/reference")
    orig = ref.ModulatedDeformConvPack.forward
    dropin.install(ref.EMA_VFI, fuse=True, patch_torchvision=False)
    try:
        assert ref.ModulatedDeformConvPack.forward is not orig and type(ref.torch).__name__ == "_TorchProxy"
    finally:
        dropin.uninstall()
    assert ref.ModulatedDeformConvPack.forward is orig and type(ref.torch).__name__ == "module"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_dropin_intercepts_the_unmodified_reference_model():
    """EMA_VFI.forward must hit the warp seam first (call site ema_vfi.py:130); on a CPU box the CUDA op then refuses."""
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.ema_vfi import EMA_VFI
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(0)
    model = EMA_VFI().eval()
    keys = [k for k in model.state_dict() if "dcn_v2" in k]
    assert keys == [f"attention_blocks.{i}.dcn_v2.{n}" for i in range(3) for n in ("weight", "bias")]
    before = dropin.call_counts()
    dropin.install(EMA_VFI)
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
            model(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))
        assert dropin.call_counts()["warp"] == before["warp"] + 1
    finally:
        dropin.uninstall()
    with torch.no_grad():
        assert model(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16)).shape == (1, 3, 16, 16)


def test_launcher_shims():
    from vfi_b200 import run

    run.apply_compat_shims()
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
    torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=5, verbose=True)  # train.py:84


def test_weight_image_k_order_is_a_bijection():
    """Host logic of the tensor-core path (no GPU): every (tap, channel) of the 3x3x67 contraction appears exactly once in
    the K order of the v6 / v7 weight image, and its main blocks follow the tcgen05.st.16x256b thread mapping
    (thread u of a pixel holds the 16-byte chunks u and u + 4)."""
    import ctypes

    from vfi_b200 import _lib

    lib = _lib.load()
    tap, c = ctypes.c_int32(), ctypes.c_int32()
    assert lib.vfi_dcn_k_order(4, 0, 0, ctypes.byref(tap), ctypes.byref(c)) != 0   # the round-1 v4 kernel and its K order are gone
    for variant, blocks in ((6, 10),):
        seen = {}
        for kb in range(blocks):
            for kk in range(64):
                tap, c = ctypes.c_int32(), ctypes.c_int32()
                assert lib.vfi_dcn_k_order(variant, kb, kk, ctypes.byref(tap), ctypes.byref(c)) == 0
                if c.value >= 0:
                    assert (tap.value, c.value) not in seen, (variant, kb, kk)
                    seen[(tap.value, c.value)] = (kb, kk)
        channels = 68
        assert set(seen) == {(t, c) for t in range(9) for c in range(channels)}
    # v6 main block: K elements 16 i + 4 u + j  <->  channel 32 (i // 2) + 8 u + 4 (i % 2) + j
    for kk in range(64):
        tap, c = ctypes.c_int32(), ctypes.c_int32()
        lib.vfi_dcn_k_order(6, 3, kk, ctypes.byref(tap), ctypes.byref(c))
        i, u, j = kk // 16, (kk // 4) % 4, kk % 4
        assert tap.value == 3 and c.value == 32 * (i // 2) + 8 * u + 4 * (i % 2) + j
    # bias slots of the v6 tail block are not weight elements
    for kk in (36, 37, 40, 63):
        tap, c = ctypes.c_int32(), ctypes.c_int32()
        lib.vfi_dcn_k_order(6, 9, kk, ctypes.byref(tap), ctypes.byref(c))
        assert c.value == -1
    assert lib.vfi_dcn_k_order(5, 0, 0, ctypes.byref(tap), ctypes.byref(c)) != 0
