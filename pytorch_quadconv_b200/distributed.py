"""
Batch-sharded data parallelism for QuadConv models.  The reference has no
distributed code; every term of the layer is independent across the batch except the parameter
gradients, so each rank (one process per GPU) runs the fused forward/backward on its contiguous
batch shard with replicated mesh/pair tables, and ONE all-reduce of a flat fp32 buffer holding every
trainable gradient (filter MLPs, bias, MeshHandler._weight_network) combines them.  The payload is
10^3..10^5 floats, i.e. latency bound, so NCCL's all-reduce over NVLink/NVSwitch is the right tool;
there is no compute step to fuse it with.

The flat buffer OWNS the reduced gradients: after `reducer()` every `p.grad` is a view into it, so a
step costs one multi-tensor copy (all fresh gradients -> flat, one launch), one `all_reduce` (AVG: the
1/world scale is folded into the collective) and nothing on the way back.  All of it is stream-ordered
and capturable: `pack()` and the optimiser step can live in CUDA graphs with the collective between
them (bench.py, C5).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_batch(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Contiguous batch chunk of rank `rank` (global batch must divide evenly)."""
    B = x.shape[0]
    if B % world_size != 0:
        raise ValueError(f"global batch {B} is not divisible by world size {world_size}")
    per = B // world_size
    return x[rank * per:(rank + 1) * per]


def trainable_parameters(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    seen, out = set(), []
    for m in modules:
        for p in m.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
    return out


class FlatGradAllReduce:
    """One flat fp32 buffer for every trainable gradient.

    `__call__(scale=None)`  = `pack()` + `reduce(scale)`; afterwards `p.grad` of every parameter is a
    view into `self.flat` holding the (scaled) sum over ranks.  `scale=None` means 1/world (a global
    mean loss whose local losses are shard means); the default path folds it into the collective
    (ReduceOp.AVG on NCCL), other scales cost one `mul_`.
    """

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = list(params)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view(p.shape))
            off += n

    # -- step pieces (each is stream-ordered and CUDA-graph capturable) ---------------------------
    def pack(self) -> None:
        """Gather the gradients autograd produced into the flat buffer (one multi-tensor copy) and make
        every `p.grad` the view that owns its slice."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g.reshape(v.shape))
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        self.attach()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v

    def reduce(self, scale: Optional[float] = None) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            if scale is None and dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
                return self.flat
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if scale is None:
                scale = 1.0 / world
        if scale is not None and scale != 1.0:
            self.flat.mul_(scale)
        return self.flat

    def __call__(self, scale: Optional[float] = None) -> torch.Tensor:
        self.pack()
        return self.reduce(scale)
