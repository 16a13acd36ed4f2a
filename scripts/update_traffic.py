#!/usr/bin/env python
"""Write profiles/traffic.json (DRAM bytes per launch of the hot kernels) from `ncu --set full` reports, stamped with the
digest of the kernel sources they were captured from: bench.py reports `roofline.traffic` only while that digest matches the
tree (a stale capture reads as null, not as a number).
usage: update_traffic.py NAME=REPORT:KERNEL_REGEX ...     e.g. dcn_fwd=gpurun_out/prof.ncu-rep:dcn_tc7_fwd"""
import csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import csrc_digest  # noqa: E402


def traffic(rep, regex):
    raw = subprocess.run(["ncu", "-i", rep, "-k", f"regex:{regex}", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    idx, units = {h: i for i, h in enumerate(rows[0])}, rows[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    vals = []
    for r in rows[2:]:
        t = sum(float(r[idx[k]].replace(",", "")) * mult.get(units[idx[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        vals.append(t)
    return sum(vals) / len(vals), len(vals)


def main():
    f = ROOT / "profiles" / "traffic.json"
    d = json.loads(f.read_text()) if f.exists() else {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=", 1)
        rep, regex = rest.rsplit(":", 1)
        t, n = traffic(rep, regex)
        d[name] = {"dram_bytes_per_launch": t, "launches_averaged": n, "report": rep, "kernel": regex, "csrc_digest": csrc_digest(),
                   "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()}
        print(name, d[name])
    f.write_text(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
