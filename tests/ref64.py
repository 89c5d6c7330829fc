"""Float64 evaluation of the reference formulas (through the oracle) on fp32 inputs: the exact value that the
1e-5 bound of north_star is measured against when the comparison with the reference's own fp32 output is too
noisy (torch's CPU fp32 bmm / scatter_add deviate from the exact value by up to 2.7e-5 on some hosts, more than
the bound; the kernels' own error is ~3e-7).  Test infrastructure only."""
import numpy as np
import torch

from oracle import quadconv_oracle as O


def mesh_params64(mesh_sd):
    """[W0,b0,...,W3,b3] (float64, requires_grad) from a MeshHandler state dict (numpy or tensors)."""
    out = []
    for l in (0, 2, 4, 6):
        for k in ("weight", "bias"):
            v = mesh_sd[f"_weight_network.{l}.{k}"]
            out.append(torch.as_tensor(np.asarray(v)).double().clone().requires_grad_())
    return out


def mesh_param_names():
    return [f"_weight_network.{l}.{k}" for l in (0, 2, 4, 6) for k in ("weight", "bias")]


def layer64(out_pts, in_pts, qw64, ws, x64, pairs, ci, co, bias):
    ws64 = [torch.as_tensor(np.asarray(w)).double().clone().requires_grad_() for w in ws]
    b64 = None if bias is None else torch.as_tensor(np.asarray(bias)).double().clone().requires_grad_()
    y = O.quadconv_forward(out_pts, in_pts, qw64, ws64, x64, pairs, ci, co, b64)
    return y, ws64, b64
