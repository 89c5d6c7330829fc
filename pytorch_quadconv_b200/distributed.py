"""
Batch-sharded data parallelism for QuadConv models.  The reference has no
distributed code; every term of the layer is independent across the batch except the parameter
gradients, so each rank (one process per GPU) runs the fused forward/backward on its contiguous
batch shard with replicated mesh/pair tables, and ONE all-reduce of a flat fp32 buffer holding every
trainable gradient (filter MLPs, bias, MeshHandler._weight_network) combines them.  The payload is
10^3..10^5 floats, i.e. latency bound, so NCCL's all-reduce over NVLink/NVSwitch is the right tool;
there is no compute step to fuse it with.
"""
from __future__ import annotations

from typing import Iterable, List

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
    """Owns one flat fp32 buffer; `__call__` packs .grad of every parameter into it, all-reduces
    (sum) once, scales by `scale` (1/world for a global-mean loss) and writes the result back."""

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = list(params)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def __call__(self, scale: float | None = None) -> torch.Tensor:
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if scale is None:
                scale = 1.0 / dist.get_world_size(self.group)
        if scale is not None and scale != 1.0:
            self.flat.mul_(scale)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[off:off + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(self.flat[off:off + n].reshape(p.shape))
            off += n
        return self.flat
