"""Validates the stripe extrapolation of bench.py's CPU reference arm: times ONE full 1080p frame (batch 1, fp32) of the
reference's CPU path (stock grid_sample + torchvision deform_conv2d CPU kernels, oracle/torch_ref.hot_path) and the 32-row
stripe the arm times, on the same host cores, and prints both frames/s figures and their ratio.  ~2 minutes of CPU."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

H, W = 1080, 1920
res = {"cores": os.cpu_count()}
stripe = bench.cpu_reference_step_factory(32, W)
ts = bench.time_cpu(stripe, 5, 1)
res["stripe_rows"] = 32
res["stripe_s_mean"] = sum(ts) / len(ts)
res["stripe_frames_per_s_extrapolated"] = (32 * W / float(H * W)) / res["stripe_s_mean"]
full = bench.cpu_reference_step_factory(H, W)
t0 = time.perf_counter()
full()
res["full_frame_s"] = time.perf_counter() - t0
res["full_frame_frames_per_s"] = 1.0 / res["full_frame_s"]
res["extrapolated_over_measured"] = res["stripe_frames_per_s_extrapolated"] / res["full_frame_frames_per_s"]
print(json.dumps(res), flush=True)
