"""GPU: batch-sharded data parallelism (parity matrix "G-way batch shard + allreduce reproduces the
single-GPU gradients of the full batch to <= 1e-5") on a QuadConv + Mesh_MaxPool autoencoder.

* one GPU: the G ranks run one after the other on the same device, the all-reduce is the sum of their flat
  gradient buffers (what NCCL computes) -- runs on every GPU box;
* >= 2 GPUs: a real 2-rank NCCL job (one process per GPU, FlatGradAllReduce), skipped on single-GPU boxes.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _model(dev, seed=0, n=1500, channels=(4, 8, 8), B=8):
    import pytorch_quadconv_b200 as q
    from bench import QuadConvAutoencoder
    torch.manual_seed(seed)
    pts = torch.rand(n, 2)
    x = torch.randn(B, channels[0], n)
    mesh = q.MeshHandler(pts).construct([n] + [0] * (len(channels) - 1), mirror=True, quad_map="mfnus")
    model = QuadConvAutoencoder(mesh, list(channels), [8, 8]).to(dev)
    return model, x.to(dev)


def _grads(model, x, params):
    for p in params:
        p.grad = None
    y = model(x)
    loss = torch.nn.functional.mse_loss(y, x)
    loss.backward()
    return loss.detach()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gradients_equal_full_batch_single_device(cuda_lib, world):
    from pytorch_quadconv_b200.distributed import FlatGradAllReduce, shard_batch, trainable_parameters
    dev = torch.device("cuda", 0)
    model, x = _model(dev)
    params = trainable_parameters([model])
    assert any("_weight_network" in n for n, _ in model.named_parameters())     # the mesh network is part of the flat buffer
    loss_full = _grads(model, x, params)
    full = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    red = FlatGradAllReduce(params)
    total = torch.zeros_like(red.flat)
    loss_sum = 0.0
    for r in range(world):
        loss_sum += float(_grads(model, shard_batch(x, r, world), params))
        red.pack()                                   # fresh gradients -> flat buffer, p.grad = views
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, red.views))
        total += red.flat
    total /= world                                   # ReduceOp.AVG
    assert abs(loss_sum / world - float(loss_full)) < TOL * abs(float(loss_full))
    assert _rel(total, full) < TOL
    off = 0
    for p in params:                                 # every parameter separately (small ones must not hide in the norm)
        n = p.numel()
        assert _rel(total[off:off + n], full[off:off + n]) < TOL
        off += n


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from pytorch_quadconv_b200.distributed import FlatGradAllReduce, shard_batch, trainable_parameters
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model, x = _model(dev)                           # same seed on every rank: replicated mesh, tables, parameters
    params = trainable_parameters([model])
    _grads(model, x, params)
    full = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    red = FlatGradAllReduce(params)
    _grads(model, shard_batch(x, rank, world), params)
    flat = red()                                     # pack + ncclAllReduce(AVG)
    torch.cuda.synchronize()
    ret[rank] = (_rel(flat, full), float(max(_rel(p.grad.reshape(-1), g) for p, g in
                                             zip(params, torch.split(full, [p.numel() for p in params])))))
    dist.destroy_process_group()


def test_two_rank_nccl_gradients_equal_full_batch(cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); the single-device test above covers the arithmetic")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_nccl_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for r in range(2):
        assert ret[r][0] < TOL and ret[r][1] < TOL, ret[r]


def test_layer_on_non_current_device(cuda_lib):
    """A model on cuda:1 while the current device is cuda:0 launches on cuda:1 (ops are device-guarded)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytorch_quadconv_b200 as q
    torch.cuda.set_device(0)
    torch.manual_seed(3)
    pts = torch.rand(900, 2)
    x = torch.randn(8, 8, 900)
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        torch.manual_seed(4)
        mesh = q.MeshHandler(pts).construct([900]).to(dev)
        layer = q.QuadConv(spatial_dim=2, in_points=900, out_points=900, in_channels=8, out_channels=8,
                           filter_seq=[8, 8], output_same=True).to(dev)
        xx = x.to(dev).requires_grad_()
        y = layer(mesh, xx)
        y.sum().backward()
        assert y.device == dev and xx.grad.device == dev
        outs.append((y.detach().cpu(), xx.grad.cpu()))
    assert torch.cuda.current_device() == 0
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-5 * float(outs[0][0].abs().max())
    assert float((outs[0][1] - outs[1][1]).abs().max()) < 1e-5 * float(outs[0][1].abs().max())
