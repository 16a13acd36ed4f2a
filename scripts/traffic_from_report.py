"""On the GPU box: DRAM bytes per launch of the bench's hot kernels from one `ncu --set full` report, split by kernel
instantiation (the staged warp's tail-record and planar forms share a function name), written in the format bench.measured_traffic()
reads.  usage: traffic_from_report.py REPORT OUT.json [COMMIT]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import csrc_digest  # noqa: E402

rep, out = sys.argv[1], sys.argv[2]
commit = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
idx, units = {h: i for i, h in enumerate(rows[0])}, rows[1]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
groups = {"dcn_fwd": [], "warp_fwd_record_bf16": [], "warp_fwd_planar_bf16": []}
times = {k: [] for k in groups}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    t = sum(float(r[idx[k]].replace(",", "")) * mult.get(units[idx[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    dur = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)
    if "dcn_tc7_fwd" in name:
        key = "dcn_fwd"
    elif "warp_fwd_staged" in name and "__nv_bfloat16" in name:
        key = "warp_fwd_record_bf16" if "(bool)1, " in name.split("warp_fwd_staged_kernel")[1][:60] else "warp_fwd_planar_bf16"
    else:
        continue
    groups[key].append(t)
    times[key].append(dur)
d = {}
for k, v in groups.items():
    if v:
        d[k] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_averaged": len(v), "ms_under_ncu": sum(times[k]) / len(times[k]),
                "report": "ncu --set full, bench.py cfg2 (" + Path(rep).name + ")", "kernel": k, "csrc_digest": csrc_digest(), "commit": commit}
Path(out).write_text(json.dumps(d, indent=1))
print(json.dumps(d))
