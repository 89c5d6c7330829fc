#!/usr/bin/env python
"""
bench.py -- QuadConv layer fwd+bwd throughput (BASELINE.json metric: "QuadConv layer fwd+bwd Mpairs/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c4|c2|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward + backward of one QuadConv layer over one batch of synthetic features:
`y = layer(mesh, x); y.sum().backward()`, parameter-gradient all-reduce included
when N > 1.  One "pair" is one row of eval_indices (reference quadconv.py:183).

Default workload (N=1) is C3, the largest single-GPU layer configuration in BASELINE.json's configs:
100 000 uniform points in the unit cube, 32->32 channels, batch 32, filter_seq [8,8],
decay_param (150/(4 pi N))^(1/3) -> P = 4 812 390 pairs.  With N GPUs every rank runs the same
per-GPU batch on its own replica of the mesh (batch sharding, weak scaling) and one flat NCCL
all-reduce combines the parameter gradients.

Printed keys follow the driver contract; `roofline` and `cpu_baseline` are described in DESIGN.md.
`--impl reference` times the oracle's CPU restatement of the reference path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QuadConv layer fwd+bwd Mpairs/s"
UNIT = "Mpairs/s"


def workload(name: str):
    if name == "c1":
        return dict(name="C1: 2-D 2500->2500 pts, 16->16 ch, batch 8, filter_seq [4,4], r=4/sqrt(N)",
                    N=2500, d=2, ci=16, co=16, B=8, fs=[4, 4], r=None)
    if name == "c3":
        N = 100_000
        return dict(name="C3: 3-D 100k pts, 32->32 ch, batch 32, filter_seq [8,8], r=(150/(4 pi N))^(1/3)",
                    N=N, d=3, ci=32, co=32, B=32, fs=[8, 8], r=(150 / (4 * np.pi * N)) ** (1 / 3))
    if name == "c4":
        N = 10_000
        return dict(name="C4: 2-D 10k pts, 64->128 ch, batch 8, filter_seq [32,32], r=sqrt(200/(pi N)) (fp32 path)",
                    N=N, d=2, ci=64, co=128, B=8, fs=[32, 32], r=float(np.sqrt(200 / (np.pi * N))))
    if name in ("c2", "c5"):
        # QuadConv + Mesh_MaxPool autoencoder on a 10k-point 2-D mesh, 4 pooling stages
        return dict(name=("C2" if name == "c2" else "C5") + ": autoencoder 10k-pt 2-D mesh, levels via mfnus, "
                    "channels [4,8,16,32,32], filter_seq [8,8], training step (MSE + Adam)",
                    N=10_000, d=2, ci=4, co=4, B=8 if name == "c2" else 256, fs=[8, 8], r=None, model="autoencoder",
                    strong=(name == "c5"))
    raise SystemExit(f"unknown workload {name}")


def weight_args(d):
    return {"const": False} if d == 2 else {"element_size": d + 1, "dimension": d, "const": False}


# --------------------------------------------------------------------------------------------
# algorithmic work
# --------------------------------------------------------------------------------------------

def flops_per_pair(w, k_avg):
    d, ci, co, B, fs = w["d"], w["ci"], w["co"], w["B"], w["fs"]
    dims = [d] + list(fs)
    S = sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))
    hL = dims[-1]
    direct_fwd = 2 * (S + hL * ci * co) + 2 * B * ci * co
    direct_bwd = 4 * B * ci * co + 4 * hL * ci * co + 2 * ci * co + 2 * S + 2 * (S - d * dims[1] if fs else 0)
    fact_fwd = 2 * S + 2 * hL * B * ci + 2 * B * ci * co * hL / k_avg
    # backward of the factorised form: d phi (dense dT + dots), d features (stage B + C) and d W_last
    fact_bwd = (2 * hL * B * ci + 2 * B * ci * co * hL / k_avg) + (2 * hL * B * co + 2 * B * ci * co * hL / k_avg) \
        + 2 * B * ci * co * hL / k_avg + 4 * S
    return dict(direct=direct_fwd + direct_bwd, factorised=fact_fwd + fact_bwd, fact_fwd=fact_fwd)


def bytes_per_step(w, P):
    N, d, ci, co, B = w["N"], w["d"], w["ci"], w["co"], w["B"]
    fwd = 4 * (B * ci * N + B * co * N + 2 * N * d + N) + 4 * P
    bwd = 4 * (B * co * N + 2 * B * ci * N + 2 * N + 2 * N * d) + 4 * P
    return fwd + bwd


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of the reference on host cores
# --------------------------------------------------------------------------------------------

def cpu_sample_config(w):
    """Bounded sample of the workload: same channels / batch / filter / neighbours-per-point,
    fewer points (the reference's materialised [P,Ci,Co] tensors do not fit at full size)."""
    s = dict(w)
    if w["N"] > 5000:
        s["N"] = 5000 if w["d"] == 3 else 2500
        if w["r"] is not None:   # keep the expected neighbour count: r ~ N^(-1/d)
            s["r"] = w["r"] * (w["N"] / s["N"]) ** (1.0 / w["d"])
    return s


def run_cpu_reference(w, steps, warmup, gpu_twin=False):
    """`y = layer(mesh, x); y.sum().backward()` with the oracle's torch-CPU restatement of
    quadconv.py:238-265 (same ATen ops and association order as the reference), all host threads.
    gpu_twin=True also times the SAME ATen composition on the SAME sample with the tensors on cuda:0 -- what a user
    of the reference gets by calling `.cuda()` on it -- and reports it inside the record (`same_sample_aten_on_gpu`):
    the like-for-like number next to the CPU one (the reference's formulation cannot run the full C3/C4 sizes on
    either device: its [P,Ci,Co] filter tensor and [B,C,P] gathers are tens of GB)."""
    from oracle import quadconv_oracle as O
    s = cpu_sample_config(w)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    pts = torch.rand(s["N"], s["d"])
    x = torch.randn(s["B"], s["ci"], s["N"], requires_grad=True)
    r = s["r"] if s["r"] is not None else 4 / np.sqrt(s["N"])
    dims = [s["d"]] + list(s["fs"]) + [s["ci"] * s["co"]]
    ws = [torch.nn.Linear(dims[l], dims[l + 1], bias=False).weight.detach().requires_grad_()
          for l in range(len(dims) - 1)]
    qw = torch.rand(s["N"], requires_grad=True)
    pairs = torch.from_numpy(O.eval_indices_c(pts.numpy(), pts.numpy(), r))
    P = pairs.shape[0]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y = O.quadconv_forward(pts, pts, qw, ws, x, pairs, s["ci"], s["co"], None, block_pairs=1 << 16)
        y.sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        x.grad = None
    t = sum(times) / len(times)
    sample = (f"oracle port of quadconv.py:238-265 (+autograd) on N={s['N']} pts (same channels/batch/filter, "
              f"radius rescaled to keep ~{P / s['N']:.0f} neighbours/pt), P={P} pairs, {t:.2f} s/step")
    rec = dict(value=P / t / 1e6, unit=UNIT, cores=cores, kind="port", sample=sample)
    if gpu_twin and torch.cuda.is_available():
        dev = torch.device("cuda:0")
        g = lambda v: v.detach().to(dev).requires_grad_(v.requires_grad)
        pts_g, x_g, qw_g, pairs_g = pts.to(dev), g(x), g(qw), pairs.to(dev)
        ws_g = [g(v) for v in ws]

        def gstep():
            x_g.grad = None
            O.quadconv_forward(pts_g, pts_g, qw_g, ws_g, x_g, pairs_g, s["ci"], s["co"], None, block_pairs=1 << 16).sum().backward()
        for _ in range(3):
            gstep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gstep()
        e1.record()
        torch.cuda.synchronize()
        tg = e0.elapsed_time(e1) / 10 * 1e-3
        rec["same_sample_aten_on_gpu"] = dict(value=P / tg / 1e6, unit=UNIT, ms_per_step=tg * 1e3,
                                              note="the oracle's ATen composition (the reference's formulation: materialised "
                                                   "[P,Ci,Co] filters, gather, bmm, scatter_add_) with its tensors on the B200, "
                                                   "same sample; cuBLAS/ATen kernels, TF32 off")
    return rec, t, P


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = self.samples[0][1]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(mx) if mx.replace(".", "").isdigit() else None, "reasons": reasons,
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(dev):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch then places them on that node): at 8 ranks the end-to-end arm moves 8 x 410 MB of pinned
    fp32 per step through one host, and remote-node pages halve the achievable H2D rate.  Returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def build_problem(w, dev, cache=True, filter_dtype=torch.float32, compute_dtype=torch.float32):
    import pytorch_quadconv_b200 as q
    torch.manual_seed(0)
    pts = torch.rand(w["N"], w["d"])
    x_host = torch.randn(w["B"], w["ci"], w["N"]).pin_memory()
    mesh = q.MeshHandler(pts).construct([w["N"]], quad_args={"point_args": {}, "weight_args": weight_args(w["d"])})
    layer = q.QuadConv(spatial_dim=w["d"], in_points=w["N"], out_points=w["N"], in_channels=w["ci"],
                       out_channels=w["co"], filter_seq=w["fs"], output_same=True, decay_param=w["r"], cache=cache,
                       filter_dtype=filter_dtype, compute_dtype=compute_dtype)
    return mesh.to(dev), layer.to(dev), x_host


class QuadConvAutoencoder(torch.nn.Module):
    """C2: encoder 4 x (QuadConv(output_same) -> sin -> Mesh_MaxPool), decoder mirrored with the adjoint
    pool; built only from the drop-in modules, as a user of the reference would.  fuse_activation=True (default) writes
    every `pool(sin(x))` as Mesh_MaxPool(activation='sin') (the fused kernels of csrc/pool.cu: same values, a third of
    the launches and [B,C,N] passes); False composes torch.sin and the pool like the reference's scripts."""

    def __init__(self, mesh, channels, filter_seq, fuse_activation=True):
        super().__init__()
        import pytorch_quadconv_b200 as q
        self.mesh = mesh
        lv = mesh.point_seq
        nl = len(lv) - 1
        mk = lambda n, ci, co: q.QuadConv(spatial_dim=2, in_points=n, out_points=n, in_channels=ci, out_channels=co,
                                          filter_seq=filter_seq, output_same=True, bias=True)
        self.enc = torch.nn.ModuleList([mk(lv[l], channels[l], channels[l + 1]) for l in range(nl)])
        self.dec = torch.nn.ModuleList([mk(lv[l], channels[l + 1], channels[l]) for l in reversed(range(nl))])
        self.fused = bool(fuse_activation)
        act = 'sin' if self.fused else None
        self.pools = torch.nn.ModuleList([q.Mesh_MaxPool(activation=act) for _ in range(nl)])
        # the first adjoint pool follows the encoder's last pool directly; the others follow sin(decoder layer)
        self.unpools = torch.nn.ModuleList([q.Mesh_MaxPool(adjoint=True, activation=act if i > 0 else None)
                                            for i in range(nl)])

    def forward(self, x):
        self.mesh.reset()
        if self.fused:
            for conv, pool in zip(self.enc, self.pools):
                x = pool(self.mesh, conv(self.mesh, x))                  # pool(sin(.)) in one kernel
            for up, conv in zip(self.unpools, self.dec):
                x = conv(self.mesh, up(self.mesh, x))                    # up(sin(.)) for every decoder layer but the first
            return x
        for conv, pool in zip(self.enc, self.pools):
            x = pool(self.mesh, torch.sin(conv(self.mesh, x)))
        for i, (up, conv) in enumerate(zip(self.unpools, self.dec)):
            x = conv(self.mesh, up(self.mesh, x))
            if i + 1 < len(self.dec):
                x = torch.sin(x)
        return x

    def pairs(self):
        return sum(int(m.eval_indices.shape[0]) for m in list(self.enc) + list(self.dec))


def build_autoencoder(w, dev, B):
    import pytorch_quadconv_b200 as q
    torch.manual_seed(0)
    pts = torch.rand(w["N"], 2)
    x_host = torch.randn(B, 4, w["N"]).pin_memory()
    mesh = q.MeshHandler(pts).construct([w["N"], 0, 0, 0, 0], mirror=True, quad_map="mfnus")
    model = QuadConvAutoencoder(mesh, [4, 8, 16, 32, 32], w["fs"])
    return mesh.to(dev), model.to(dev), x_host


def time_kernels(w, mesh, layer, x, reps=5):
    """Per-kernel device time of one fwd+bwd, CUDA events on the launching (current) stream around
    each C-ABI call.  Returns {name: ms}."""
    from pytorch_quadconv_b200 import _cabi as C, ops
    L = C.lib()
    names = ["qc_pair_phi_fwd", "qc_pack_features", "qc_repack_wlast", "qc_contract", "qc_unpack_features",
             "qc_batch_sum", "qc_dphi", "qc_pair_phi_bwd", "qc_pair_phi_bwd_ws", "qc_unpack_dwlast", "qc_gather_accumulate", "qc_gather_dot", "qc_contract_mma", "qc_pack_features_bf16", "qc_dphi_mma", "qc_dw_mma", "qc_gemm_nt_tf32x3", "qc_gemm_tn_tf32x3"]
    acc = {}
    orig = {n: getattr(L, n) for n in names}
    phase = {"bwd": False}

    def wrap(n):
        f = orig[n]

        def g(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = f(*a)
            e1.record()
            key = n + (".bwd" if phase["bwd"] else ".fwd")
            acc.setdefault(key, []).append((e0, e1))
            return rc
        return g
    for n in names:
        setattr(L, n, wrap(n))
    single_call_limit, ops.SINGLE_CALL_MAX_BYTES = ops.SINGLE_CALL_MAX_BYTES, 0   # per-kernel table: composed ops (same kernels)
    try:
        for _ in range(reps):
            phase["bwd"] = False
            x.grad = None
            y = layer(mesh, x)
            loss = y.sum()
            phase["bwd"] = True
            loss.backward()
        torch.cuda.synchronize()
    finally:
        ops.SINGLE_CALL_MAX_BYTES = single_call_limit
        for n in names:
            setattr(L, n, orig[n])
    out = {}
    calls = 0
    for k, evs in acc.items():
        per_step = len(evs) // reps
        calls += per_step
        ts = [a.elapsed_time(b) for a, b in evs[per_step:]]      # drop the first repetition
        out[k] = sum(ts) / max(reps - 1, 1)
    return out, calls


def run_ours(args, w, rank, local_rank, world, dev, sub=False, tc_table=False):
    """One workload on this rank's GPU.  Returns the JSON line (rank 0) or None.  sub=True: a secondary record
    riding on the default line (no CPU baseline, no per-kernel table)."""
    import torch.distributed as dist
    from pytorch_quadconv_b200.distributed import FlatGradAllReduce, trainable_parameters
    is_ae = w.get("model") == "autoencoder"
    # A step is three stream-ordered pieces: compute (zero grads, forward, backward, gradients gathered into the
    # flat buffer), the ONE flat all-reduce (eager NCCL call; nothing when world == 1), and the update (Adam for
    # the autoencoder, nothing for the single-layer workloads).  With --graph the first and the last are CUDA
    # graphs replayed per rank, the collective stays between them.
    if is_ae:
        b_local = w["B"] // world if w.get("strong") else w["B"]
        mesh, layer, x_host = build_autoencoder(w, dev, b_local)
        x = x_host.to(dev)
        params = trainable_parameters([layer])      # the model owns the mesh (and its weight network)
        reducer = FlatGradAllReduce(params)
        opt = torch.optim.Adam(params, lr=1e-3, capturable=args.graph)

        def compute(xin):
            for p in params:
                p.grad = None
            y = layer(xin)
            loss = torch.nn.functional.mse_loss(y, xin.detach())
            loss.backward()
            reducer.pack()
            return loss

        update = opt.step
    else:
        mesh, layer, x_host = build_problem(w, dev, cache=not args.no_cache,
                                            filter_dtype=torch.bfloat16 if args.filter_dtype == "bf16" else torch.float32,
                                            compute_dtype=torch.bfloat16 if args.compute_dtype == "bf16" else torch.float32)
        x = x_host.to(dev).requires_grad_()
        params = trainable_parameters([layer, mesh])
        reducer = FlatGradAllReduce(params)

        def compute(xin):
            for p in params:
                p.grad = None
            xin.grad = None
            y = layer(mesh, xin)
            loss = y.sum()
            loss.backward()
            reducer.pack()
            return loss

        update = None

    def step(xin):
        loss = compute(xin)
        if world > 1:
            reducer.reduce()
        if update is not None:
            update()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x)
    P = layer.pairs() if is_ae else int(layer._tables[0].numel())

    if args.graph:
        # compute piece and update piece replayed from CUDA graphs (pytorch_quadconv_b200.graphed), the collective between them
        from pytorch_quadconv_b200 import graphed
        g_compute = graphed(compute, x, warmup=3)
        x_static = g_compute.static_inputs[0]
        g_update = graphed(update, warmup=1) if update is not None else None

        def step(xin):
            loss = g_compute(xin)
            if world > 1:
                reducer.reduce()
            if g_update is not None:
                g_update()
            return loss
        x = x_static

    # ---- device-resident timing --------------------------------------------------------------
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            step(x)
        e1.record()
        sync()
    ms = e0.elapsed_time(e1) / args.steps
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    # weak scaling: every rank walks the pair list for its own batch -> P * world pairs per step; strong scaling
    # (C5): the ranks share ONE global batch, so a step is P pairs however many ranks split it
    units = P * (1 if w.get("strong") else world)
    value = units / (ms * 1e-3) / 1e6

    # ---- end to end: host (pinned) features -> device -> fwd+bwd -> loss + filter grads to host ----
    grads_host = [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in layer.parameters() if p.requires_grad]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    # The H2D copy of step i+1 runs on a copy stream into the other of two device buffers while step i
    # computes (every step's copy is inside the timed region; a user's input pipeline does the same).
    copy_stream = torch.cuda.Stream()
    if args.compute_dtype == "bf16" and not is_ae:
        # the reduced-precision mode rounds the features to bf16 on entry (csrc/contract_mma.cu, k_pack_bf16), so its
        # input pipeline hands them over in bf16: same values on the tensor cores, half the host-to-device bytes
        x_host = x_host.to(torch.bfloat16).pin_memory()
    xbufs = [torch.empty_like(x_host, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def h2d_async(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])          # the step that used this buffer is done
            xbufs[i & 1].copy_(x_host, non_blocking=True)
            ready[i & 1].record(copy_stream)

    def e2e_step(i, prefetch):
        torch.cuda.current_stream().wait_event(ready[i & 1])
        xin = xbufs[i & 1]
        if prefetch:
            h2d_async(i + 1)
        if not is_ae:
            xin = xin.detach().requires_grad_()
        loss = step(xin)
        consumed[i & 1].record(torch.cuda.current_stream())
        loss_host.copy_(loss.detach(), non_blocking=True)
        for h, p in zip(grads_host, [p for p in layer.parameters() if p.requires_grad]):
            h.copy_(p.grad, non_blocking=True)

    for ev in consumed:
        ev.record(torch.cuda.current_stream())
    h2d_async(0)
    for i in range(2):
        e2e_step(i, True)
    sync()
    n_e2e = max(3, args.steps // 2)
    # the copy for the first timed step was issued during warm-up; issue the last step's successor inside
    # the region instead, so exactly n_e2e copies are timed
    e0.record()
    for i in range(2, 2 + n_e2e):
        e2e_step(i, True)
    e1.record()
    sync()
    ems = torch.tensor([e0.elapsed_time(e1) / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = float(ems.item())
    h2d = x_host.numel() * x_host.element_size()
    d2h = 4 + sum(g.numel() * 4 for g in grads_host)

    # ---- per-kernel table + roofline of the dominant kernel (rank 0) ----------------------------
    roof, cb, table = None, None, None
    if rank == 0 and is_ae:
        roof = {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                "note": "multi-layer training step: launch-latency dominated (SURVEY 8d); per-kernel roofline is "
                        "reported on the single-layer workloads (default c3)"}
    if rank == 0 and not is_ae and sub and tc_table:
        # reduced-precision mode: per-kernel table and the roofline of its dominant kernel against the measured bf16 peak
        table, calls_per_step = time_kernels(w, mesh, layer, x)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops", 1590.0)
        B, ci, co, hL = w["B"], w["ci"], w["co"], (w["fs"][-1] if w["fs"] else w["d"])
        stageB, dense = 2.0 * P * B * hL, 2.0 * w["N"] * B * ci * co * hL
        alg = {"qc_contract_mma.fwd": stageB * ci + dense, "qc_contract_mma.bwd": stageB * co + dense,
               "qc_dphi_mma.bwd": stageB * ci + dense, "qc_dw_mma.bwd": stageB * ci + dense}
        tc_k = {k: v for k, v in table.items() if k in alg}
        dom = max(tc_k, key=tc_k.get) if tc_k else None
        if dom:
            ach = alg[dom] / (tc_k[dom] * 1e-3) / 1e12
            gather_bytes = None
            try:
                gather_bytes = json.load(open(os.path.join(ROOT, "profiles", "r02_tc_traffic.json"))).get(args.workload, {}).get(dom)
            except Exception:
                pass
            roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": bf16, "unit": "TFLOP/s", "frac": ach / bf16,
                    "traffic": gather_bytes,
                    "note": f"bf16 tensor peak of {'measured (MEASURED_PEAKS.json)' if peaks else 'fallback'}; numerator = "
                            f"factorised-formulation FLOPs of this launch ({alg[dom] / 1e9:.1f} GFLOP) / CUDA-event time "
                            f"{tc_k[dom]:.3f} ms; traffic = DRAM bytes of this launch (ncu).  The kernel is bound by the "
                            f"L2->SM gather of the union rows (2.6-2.8 GB per launch = 12 bytes/clk/SM through per-lane "
                            f"cp.async; the TMA gather4 variant is 2.4x slower), not by the tensor pipe (16-25 % active, "
                            f"profiles/r02_c3_bf16_tensor_core_kernels.md)"}
    if rank == 0 and not is_ae and not sub:
        table, calls_per_step = time_kernels(w, mesh, layer, x)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops", 1590.0)
        src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        k_avg = P / w["N"]
        fl = flops_per_pair(w, k_avg)
        dom = max(table, key=table.get)
        dom_ms = table[dom]
        # algorithmic flops of the dominant launch, factorised formulation (DESIGN.md "Kernels")
        B, ci, co, hL = w["B"], w["ci"], w["co"], (w["fs"][-1] if w["fs"] else w["d"])
        stageB = 2.0 * P * B * hL
        dense = 2.0 * w["N"] * B * ci * co * hL
        alg = {"qc_contract.fwd": stageB * ci + dense, "qc_contract.bwd": stageB * co + 2 * dense,
               "qc_dphi.bwd": stageB * ci + dense,
               # wide path (csrc/contract_wide.cu): the gather kernels hold only stage B, the dense stage is cuBLAS
               "qc_gather_accumulate.fwd": stageB * ci, "qc_gather_accumulate.bwd": stageB * co,
               "qc_gather_dot.bwd": stageB * ci}.get(dom, 0.0)
        peak = bf16 / 6.0     # fp32 mode on tensor cores = 3xTF32 ~ bf16/6
        ach = alg / (dom_ms * 1e-3) / 1e12
        traffic = None
        try:   # dram__bytes_read+write of this kernel from the committed `ncu --set full` capture
            for f in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
                fp = os.path.join(ROOT, "profiles", f)
                if os.path.exists(fp):
                    traffic = json.load(open(fp)).get(args.workload, {}).get(dom)
                    if traffic is not None:
                        break
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic,
                "note": f"fp32-mode tensor peak = bf16/6 of {src}; numerator = factorised-formulation FLOPs of this "
                        f"launch ({alg / 1e9:.1f} GFLOP) / CUDA-event time {dom_ms:.3f} ms; the dense stage of the kernel "
                        f"runs on tcgen05 (3xTF32), the per-pair gather-accumulate on the FP32 pipe, which bounds it "
                        f"(FFMA peak = 148 SM x 128 lanes x 2 x clock ~ 74 TFLOP/s)",
                "step_direct_formulation_tflops": fl["direct"] * P / (ms * 1e-3) / 1e12,
                "step_factorised_tflops": fl["factorised"] * P / (ms * 1e-3) / 1e12,
                "step_hbm_alg_gbs": bytes_per_step(w, P) / (ms * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs", 6650.0)}
        if args.kernel_table:
            json.dump({"workload": w["name"], "P": P, "ms_per_step": ms, "kernels_ms": table},
                      open(args.kernel_table, "w"), indent=1)
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = run_cpu_reference(w, 1, 1, gpu_twin=True)

    if rank == 0:
        launches = 0
        if is_ae:
            launches = 8 * 17 + 8 * 2 + 4 * 4      # 8 QuadConv layers x 17 kernels (fwd+bwd), 8 pool layers x 2, 4 weight maps x 4
        if table and sub:
            launches = calls_per_step + 4
        elif table:
            # every timed C-ABI call launches one of our kernels (csrc/*.cu); the dW variant of qc_contract adds
            # k_dw_reduce + k_dw_sum_ctas, qc_pair_phi_bwd_ws its finishing kernel, the fused weight map
            # (MeshHandler.weights: k_wm_fwd, k_vertex_sum, k_wm_bwd, k_wm_finish) four more per step; memsets are not counted
            launches = (calls_per_step + (2 if "qc_contract.bwd" in table else 0) +
                        (1 if "qc_pair_phi_bwd_ws.bwd" in table else 0) + 4)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if w.get("strong") else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.compute_dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": w["name"], "pairs_per_gpu": P,
                       "batch_per_gpu": (w["B"] // world if w.get("strong") else w["B"]),
                       "global_batch": w["B"] if w.get("strong") else w["B"] * world,
                       "parallelism": f"batch-sharded dp{world}, one flat NCCL all-reduce (AVG) of all parameter grads "
                                      f"({reducer.numel} floats)",
                       "l2": "inputs larger than L2 (features+grads 1.6 GB/step > 126 MB)" if w["N"] >= 50000
                       else "working set fits in L2 (reference's own CPU-runnable case)",
                       "includes": "MeshHandler.weights() (fused weight-map kernels, mesh_handler.py:101-112) every step"
                                   + ("; sin activations evaluated inside the pool kernels (Mesh_MaxPool(activation='sin'))"
                                      if is_ae and getattr(layer, "fused", False) else ""),
                       "cuda_graph": bool(args.graph), "cache": not args.no_cache, "filter_mlp_operands": args.filter_dtype,
                       "compute_dtype": args.compute_dtype},
            "clocks": clk.summary(),
            "e2e": {"value": units / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    # achieved host-to-device rate of THIS rank's input stream (the copies overlap the compute; when
                    # this approaches the link / host-memory rate the end-to-end step is input bound, as at 8 ranks)
                    "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "numa_node": getattr(args, "numa_node", None),
                    "returns": "loss + every parameter gradient (a training step); y and d_features stay on the device"},
            "gpu_launches": launches * args.steps,
            "roofline": roof, "cpu_baseline": cb, "kernels_ms": table,
        }
        return line
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cache", action="store_true",
                    help="QuadConv(cache=False): the neighbour search runs inside every forward "
                         "(the mode of the reference's py_scripts/quadconv_layer_profile.py:46)")
    ap.add_argument("--filter-dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16: the filter MLP takes bf16 operands with fp32 accumulation (C4's definition in "
                         "BASELINE.json); the rest of the layer stays fp32 (single-layer workloads)")
    ap.add_argument("--compute-dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16: the reduced-precision tensor-core mode (both contraction stages on tcgen05 with bf16 "
                         "operands, fp32 accumulation; stated bound 1e-2 normwise).  f32 is the parity mode and the headline")
    ap.add_argument("--graph", action="store_true",
                    help="capture the step in a CUDA graph and replay it (small meshes are launch-latency bound)")
    ap.add_argument("--no-tc", action="store_true", help="skip the bf16 tensor-core-mode sub-record of the default (c3) line")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 sub-record that rides on the default (c3) line")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch size (profiling aid)")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel timing table (json) here")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    w = workload(args.workload)
    if args.batch:
        w["B"] = args.batch
        w["name"] += f" [batch overridden: {args.batch}]"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        cb, t, P = run_cpu_reference(w, steps, 1)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "note": "CPU host cores; bounded sample, see cpu_baseline.sample"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pytorch_quadconv_b200 import _cabi  # fails loudly if the .so is missing
    _cabi.lib()
    args.numa_node = bind_to_gpu_numa_node(dev) if world > 1 else None

    line = run_ours(args, w, rank, local_rank, world, dev)
    if args.workload == "c3" and args.compute_dtype == "f32" and not args.no_tc:
        # the reduced-precision tensor-core mode of the same workload (north_star: "with a stated bf16 bound"): all four
        # contractions on tcgen05 with bf16 operands, stated bound 1e-2 normwise (tests/test_gpu_mma.py).  Reported next
        # to the fp32 parity mode, which stays the headline `value`.
        tc_args = argparse.Namespace(**vars(args))
        tc_args.compute_dtype, tc_args.steps, tc_args.warmup = "bf16", max(5, args.steps // 2), 3
        tcl = run_ours(tc_args, w, rank, local_rank, world, dev, sub=True, tc_table=True)
        if line is not None:
            line["bf16_tensor_core_mode"] = tcl
    if args.workload == "c3" and not args.no_c5:
        # north_star config 5: the batch-256 autoencoder training step, batch split over the
        # ranks (STRONG scaling), step replayed from per-rank CUDA graphs with the flat NCCL all-reduce between them
        sub_args = argparse.Namespace(**vars(args))
        sub_args.graph, sub_args.steps, sub_args.warmup = True, max(5, args.steps // 2), 3
        c5 = run_ours(sub_args, workload("c5"), rank, local_rank, world, dev, sub=True)
        if line is not None:
            line["c5"] = c5
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
