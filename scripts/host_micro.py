"""Where does the host time of a small-mesh layer step go?  Times, for the C1 layer: the raw C calls through ctypes,
the same behind the torch.library ops, and the whole module step (enqueue time and wall time per iteration)."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from pytorch_quadconv_b200 import _cabi as C, ops  # noqa: E402


def timeit(name, fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:55s} enqueue {1e6*(t1-t0)/n:8.1f} us   wall {1e6*(t2-t0)/n:8.1f} us", flush=True)


def main():
    w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c1")
    dev = torch.device("cuda:0")
    mesh, layer, x_host = bench.build_problem(w, dev)
    x = x_host.to(dev).requires_grad_()
    y = layer(mesh, x)
    y.sum().backward()
    t = layer._tables
    pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in = t[:8]
    pts = mesh.input_points().detach()
    lin = [m for m in layer.filter if isinstance(m, torch.nn.Linear)]
    hidden = [m.weight.detach() for m in lin[:-1]]
    w_last = lin[-1].weight.detach()
    qw = mesh.weights().detach()
    xd = x.detach()
    B, Ci, N = xd.shape
    Co = layer.out_channels
    desc = ops._layer_desc(pts, pts, qw, hidden, w_last, None, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row,
                           order_out, order_in, Ci, Co)
    L = C.lib()
    ref = ctypes.byref(desc)
    wsb = L.qc_layer_workspace_bytes(ref, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty(B, Co, N, device=dev)
    g = torch.ones_like(out)
    dx = torch.empty_like(xd)
    dqw = torch.empty_like(qw)
    dwl = torch.empty_like(w_last)
    dh = [torch.empty_like(h) for h in hidden]
    dptrs = (ctypes.c_void_p * len(dh))(*[h.data_ptr() for h in dh])
    st = C.stream()
    timeit("empty python loop body", lambda: None)
    timeit("torch.empty x1", lambda: torch.empty(16, device=dev))
    timeit("raw C qc_layer_forward", lambda: L.qc_layer_forward(ref, xd.data_ptr(), B, out.data_ptr(), ws.data_ptr(), wsb, st))
    timeit("raw C qc_layer_backward", lambda: L.qc_layer_backward(ref, g.data_ptr(), B, dx.data_ptr(), dqw.data_ptr(), dwl.data_ptr(), dptrs, None, ws.data_ptr(), wsb, st))
    args_f = (xd, pts, pts, qw, hidden, w_last, None, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in, Ci, Co)
    timeit("ops.layer_forward (torch.library op)", lambda: ops.layer_forward(*args_f))
    timeit("ops.layer_forward._init_fn (python body only)", lambda: ops.layer_forward._init_fn(*args_f))
    o, wsx = ops.layer_forward(*args_f)
    args_b = (g, wsx, pts, pts, qw, hidden, w_last, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in, Ci, Co, False)
    timeit("ops.layer_backward (torch.library op)", lambda: ops.layer_backward(*args_b))
    timeit("ops.layer_backward._init_fn (python body only)", lambda: ops.layer_backward._init_fn(*args_b))
    timeit("mesh.weights()", lambda: mesh.weights())

    def fwd_only():
        with torch.no_grad():
            layer(mesh, xd)
    timeit("module forward, no_grad", fwd_only)

    def step():
        x.grad = None
        layer(mesh, x).sum().backward()
    timeit("module step (fwd + sum + backward)", step)


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] == "steps"):
    main()


def main2():
    """Break the module step down: forward with grad, sum, backward; then a torch.profiler CPU table of 20 steps."""
    w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c1")
    dev = torch.device("cuda:0")
    mesh, layer, x_host = bench.build_problem(w, dev)
    x = x_host.to(dev).requires_grad_()
    n = 200
    tf = ts = tb = 0.0
    for i in range(n + 20):
        x.grad = None
        t0 = time.perf_counter()
        y = layer(mesh, x)
        t1 = time.perf_counter()
        loss = y.sum()
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        if i >= 20:
            tf += t1 - t0
            ts += t2 - t1
            tb += t3 - t2
    torch.cuda.synchronize()
    print(f"forward(grad) {1e6*tf/n:.1f} us  sum {1e6*ts/n:.1f} us  backward {1e6*tb/n:.1f} us", flush=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            x.grad = None
            layer(mesh, x).sum().backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "steps":
    main2()
