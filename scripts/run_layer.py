"""Small driver for profiling: builds one bench workload and runs a few fwd(+bwd) steps.
    python scripts/run_layer.py [--workload c3] [--compute-dtype bf16] [--steps 3] [--fwd-only]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--compute-dtype", default="f32")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--points", type=int, default=0, help="override the number of points (radius rescaled to keep the neighbour count)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
w = bench.workload(a.workload)
if a.points:
    if w["r"] is not None:
        w["r"] = w["r"] * (w["N"] / a.points) ** (1.0 / w["d"])
    w["N"] = a.points
mesh, layer, x_host = bench.build_problem(w, dev, compute_dtype=torch.bfloat16 if a.compute_dtype == "bf16" else torch.float32)
x = x_host.to(dev).requires_grad_(not a.fwd_only)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
for i in range(a.steps):
    ev[i].record()
    if a.fwd_only:
        with torch.no_grad():
            y = layer(mesh, x)
    else:
        x.grad = None
        y = layer(mesh, x)
        y.sum().backward()
ev[a.steps].record()
torch.cuda.synchronize()
print("P", int(layer._tables[0].numel()), "ms per step:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(a.steps)], "tc_active", layer.tc_active)
