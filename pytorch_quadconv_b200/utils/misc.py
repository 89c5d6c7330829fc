"""Sin activation module (reference torch_quadconv/utils/misc.py:13-19).  It is kept inside
`QuadConv.filter` so the Sequential indices -- and therefore the state_dict keys
`filter.{0,2,...}.weight` -- match the reference; the fused kernels evaluate sin themselves."""
import torch
import torch.nn as nn


class Sin(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, x):
        return torch.sin(x)
