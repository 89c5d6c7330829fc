"""GPU-time table per kernel of an autoencoder training step (C2/C5 model) at a given batch: torch.profiler CUDA
activity, aggregated by kernel name.  Usage: python scripts/kernel_table.py <batch> [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    w = bench.workload("c5")
    dev = torch.device("cuda:0")
    mesh, model, x_host = bench.build_autoencoder(w, dev, B)
    x = x_host.to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)

    def step():
        for p in params:
            p.grad = None
        y = model(x)
        loss = torch.nn.functional.mse_loss(y, x)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    rows = [(e.key, e.self_device_time_total / steps, e.count // steps) for e in prof.key_averages() if e.self_device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"batch {B}: GPU time per step {tot/1e3:.3f} ms over {sum(r[2] for r in rows)} launches")
    for k, t, n in rows[:40]:
        print(f"{t/1e3:9.3f} ms  {100*t/tot:5.1f}%  x{n:3d}  {k[:150]}")


if __name__ == "__main__" and not (len(sys.argv) > 3 and sys.argv[3] == "launches"):
    main()


def per_launch():
    """Per-launch durations of the contraction kernels in launch order (one step)."""
    B = int(sys.argv[1])
    w = bench.workload("c5")
    dev = torch.device("cuda:0")
    mesh, model, x_host = bench.build_autoencoder(w, dev, B)
    x = x_host.to(dev)
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        y = model(x)
        torch.nn.functional.mse_loss(y, x).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("k_contract" in e.name or "k_dphi" in e.name)]
    evs.sort(key=lambda e: e.time_range.start)
    lv = mesh.point_seq
    print("levels", lv, "pairs", [int(m.eval_indices.shape[0]) for m in list(model.enc) + list(model.dec)])
    for e in evs:
        print(f"{e.device_time:9.1f} us  {e.name[:80]}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "launches":
    per_launch()
