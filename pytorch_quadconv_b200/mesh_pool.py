"""
Mesh_MaxPool -- drop-in for torch_quadconv.Mesh_MaxPool (reference torch_quadconv/mesh_pool.py:15-98).

forward(handler, data[B,C,N]) pools fine -> coarse with a max over the MFNUS elimination groups
(adjoint=False, mesh_pool.py:88-89) or copies coarse -> fine (adjoint=True, mesh_pool.py:93), then
advances the handler cursor (mesh_pool.py:96).  The group index is built once on the host with the
reference's rule -- keys in dict order, a fine point claimed by several keepers goes to the LAST one
(mesh_pool.py:47-49, :70-72) -- as one int32 table shared by all (batch, channel) rows instead of
the reference's int64 [B,C,N] tensor, and the reduction runs as a segmented-max CUDA kernel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops


def _group_of_fine_points(elim_map: dict, n_fine: int) -> np.ndarray:
    """group[k] = position (dict order) of the last keeper whose member list (+ itself) holds k."""
    group = np.zeros(n_fine, dtype=np.int64)
    for i, (key, members) in enumerate(elim_map.items()):
        m = np.asarray(list(members) + [key], dtype=np.int64)
        group[m] = i
    return group


class Mesh_MaxPool(nn.Module):

    def __init__(self, adjoint=False, activation=None):
        super().__init__()
        self.adjoint = adjoint
        # extension over the reference's signature: activation='sin' computes pool(sin(data)) in the pool
        # kernels -- the composition the reference's autoencoders write as separate modules -- saving the activation's
        # own launch and its [B,C,N] round trip in both directions.  None = the reference's module.
        if activation not in (None, 'sin'):
            raise ValueError(f"activation must be None or 'sin' (got {activation!r})")
        self.activation = activation
        self.cached = False
        self.index = None
        self.output_shape = None
        self._tables = None

    def _build(self, handler, x):
        dev = x.device
        if not self.adjoint:
            n_out = handler.output_points().shape[0]
            elim_map = handler.get_downsample_map(n_out)                  # mesh_pool.py:40
            n_fine = x.shape[2]
        else:
            n_coarse = handler.input_points().shape[0]
            elim_map = handler.get_downsample_map(n_coarse)               # mesh_pool.py:55
            n_fine = handler.output_points().shape[0]
            n_out = n_fine
        group = _group_of_fine_points(elim_map, n_fine)
        n_groups = len(elim_map)
        order = np.argsort(group, kind="stable")
        counts = np.bincount(group, minlength=n_groups)
        grp_ptr = np.zeros(n_groups + 1, dtype=np.int32)
        np.cumsum(counts, out=grp_ptr[1:])
        self.index = torch.from_numpy(group).to(dev)                      # int64 [N_fine]
        self._tables = (torch.from_numpy(group.astype(np.int32)).to(dev),
                        torch.from_numpy(grp_ptr).to(dev),
                        torch.from_numpy(order.astype(np.int32)).to(dev),
                        n_groups)
        self.output_shape = (x.shape[0], x.shape[1], n_groups if not self.adjoint else n_out)

    def forward(self, handler, data):
        x = data
        if not x.is_cuda:
            raise RuntimeError("pytorch_quadconv_b200 has no CPU path: Mesh_MaxPool needs CUDA tensors")
        if self.index is not None and self.index.device != x.device:
            self.index = None
        if self.index is None:
            self._build(handler, x)
        group32, grp_ptr, grp_mem, n_groups = self._tables
        act = 1 if self.activation == 'sin' else 0
        self.output_shape = (x.shape[0], x.shape[1], self.output_shape[2])

        if not self.adjoint:
            output = ops.segmax_apply(x, group32, grp_ptr, grp_mem, n_groups, act)  # mesh_pool.py:88-89
        else:
            output = ops.gather_pool_apply(x, group32, grp_ptr, grp_mem, act)  # mesh_pool.py:93

        handler.step()                                                    # mesh_pool.py:96
        return output


# BASELINE.json spells the class without the underscore
MeshMaxPool = Mesh_MaxPool
