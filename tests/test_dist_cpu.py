"""CPU, world_size 2, gloo: the batch-shard + flat-gradient all-reduce host logic."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from pytorch_quadconv_b200.distributed import FlatGradAllReduce, shard_batch, trainable_parameters
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    net.register_parameter("frozen", frozen)
    x = torch.randn(8, 6)
    # reference: full batch, mean loss
    full = net(x).pow(2).mean()
    gref = torch.autograd.grad(full, trainable_parameters([net]))
    xs = shard_batch(x, rank, world)
    loss = net(xs).pow(2).mean()      # local mean over the shard
    loss.backward()
    params = trainable_parameters([net])
    assert all(p.requires_grad for p in params) and len(params) == 4
    FlatGradAllReduce(params)()       # sum / world == global mean gradient
    err = max(float((p.grad - g).abs().max()) for p, g in zip(params, gref))
    ret[rank] = err
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_matches_full_batch():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    assert max(ret.values()) < 1e-6
