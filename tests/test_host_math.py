"""Host-side unit test of the kernels' coordinate arithmetic (csrc/warp_math.h compiled with g++)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_markstein_division_and_coordinate_replay_are_bit_exact(tmp_path):
    exe = tmp_path / "warp_math_check"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tests/host/warp_math_check.cpp")],
                   check=True)
    r = subprocess.run([str(exe), "3000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout
