"""Host-side profile of an eager (no CUDA graph) C2 autoencoder training step: where the CPU time of the ~40 launches
per layer goes (torch.library dispatch, ctypes calls, allocations).  Usage: python scripts/host_profile.py [c2|c1]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    w = bench.workload(name)
    dev = torch.device("cuda:0")
    if w.get("model") == "autoencoder":
        mesh, model, x_host = bench.build_autoencoder(w, dev, w["B"])
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
    else:
        mesh, layer, x_host = bench.build_problem(w, dev)
        x = x_host.to(dev).requires_grad_()

        def step():
            x.grad = None
            y = layer(mesh, x)
            y.sum().backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    print(f"{name}: host enqueue {t_host*1e3:.3f} ms/step, wall {t_all*1e3:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(35)
    st.sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
