"""Host-side logic of bench.py that needs no GPU: the traffic stamp (a number from an ncu capture counts only for the kernel
sources it was captured from), the measured-peaks loader, the gather-floor reader, and the CPU reference arm's JSON contract."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_traffic_counts_only_for_the_sources_it_was_captured_from(tmp_path, monkeypatch):
    real = bench.csrc_digest()
    assert len(real) == 16 and real == bench.csrc_digest()                     # deterministic
    monkeypatch.setattr(bench, "ROOT", tmp_path)                               # an empty tree: its own (different) digest
    now = bench.csrc_digest()
    assert now != real
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "traffic.json").write_text(json.dumps({
        "dcn_fwd": {"dram_bytes_per_launch": 5.0e9, "csrc_digest": now, "commit": "abc", "report": "r"},
        "warp_fwd_planar_bf16": {"dram_bytes_per_launch": 3.0e8, "csrc_digest": "0" * 16, "commit": "old", "report": "r"},
        "legacy_float_entry": 1.0, "_note": "free text"}))
    t = bench.measured_traffic()
    assert t["dcn_fwd"] == 5.0e9 and "abc" in t["dcn_fwd_source"]
    assert "warp_fwd_planar_bf16" not in t and t["warp_fwd_planar_bf16_source"].startswith("stale")
    assert "legacy_float_entry" not in t                                       # round-1 format: never trusted


def test_committed_traffic_file_is_well_formed():
    d = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    e = d["dcn_fwd"]
    assert e["dram_bytes_per_launch"] > 5.34e9                                  # never below the algorithmic 322 B/px x 16.6 M px
    assert len(e["csrc_digest"]) == 16 and e["commit"]


def test_peaks_and_gather_floor():
    pk = bench.peaks()
    assert pk["hbm"] > 1000 and pk["tensor_burst"] >= pk["tensor_sustained"] > 100 and pk["source"] in ("measured", "fallback")
    m = bench.measured_gather_floor()
    assert m is None or (30.0 < m["clk_per_px"] < 80.0 and m["file"].startswith("r02_"))


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference's CPU path on a bounded stripe, one JSON line with the base contract's keys."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
