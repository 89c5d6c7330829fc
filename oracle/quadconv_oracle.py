"""
oracle/quadconv_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's QuadConv hot path, used as the parity checker by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`.
Nothing under `pytorch_quadconv_b200/` imports this file; the product path has no CPU fallback.

Each function cites the reference lines (relative to /root/reference) it restates.  Because the path is
floating point, the restatement uses torch fp32 CPU ops in the same association order as the reference
(so it is also a fair *timing* stand-in for the reference on host cores); the neighbour search, which
must be bit exact, additionally exists in plain C (oracle/qc_oracle.c).

Pinning: the reference ships no golden vectors and none of its tests pass, so
the oracle is pinned against outputs of the LIVE reference imported from /root/reference in the build
container: tests/golden/make_golden.py wrote tests/golden/*.npz, and tests/test_oracle.py checks every
function here against them (pairs bit-exact, floats to 1e-6).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "qc_oracle.so")


def build_c_oracle(force: bool = False) -> str:
    """gcc build of oracle/qc_oracle.c into oracle/_build/ (git-ignored)."""
    src = os.path.join(_HERE, "qc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def _clib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
    return _lib


def _fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------
# neighbour search
# ---------------------------------------------------------------------------------------------

def radius_to_f32(decay_param) -> np.float32:
    """quadconv.py:182 compares an fp32 tensor with a python/numpy float: the scalar is demoted
    to fp32 (round to nearest) before the `<=`."""
    return np.float32(decay_param)


def eval_indices_c(out_pts: np.ndarray, in_pts: np.ndarray, decay_param) -> np.ndarray:
    """C restatement of quadconv.py:165-183.  Returns [P,2] int64, row-major (out, in) order."""
    out_pts = np.ascontiguousarray(out_pts, dtype=np.float32)
    in_pts = np.ascontiguousarray(in_pts, dtype=np.float32)
    No, d = out_pts.shape
    Ni = in_pts.shape[0]
    r32 = ctypes.c_float(float(radius_to_f32(decay_param)))
    counts = np.zeros(No, dtype=np.int64)
    lib = _clib()
    lib.qc_oracle_count(_fptr(out_pts), _fptr(in_pts), ctypes.c_int64(No), ctypes.c_int64(Ni),
                        ctypes.c_int(d), r32, _fptr(counts))
    row_ptr = np.zeros(No + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    pairs = np.empty((int(row_ptr[-1]), 2), dtype=np.int64)
    lib.qc_oracle_fill(_fptr(out_pts), _fptr(in_pts), ctypes.c_int64(No), ctypes.c_int64(Ni),
                       ctypes.c_int(d), r32, _fptr(row_ptr), _fptr(pairs))
    return pairs


def eval_indices_torch(out_pts: torch.Tensor, in_pts: torch.Tensor, decay_param,
                       block: int = 2048) -> torch.Tensor:
    """Row-blocked torch restatement of quadconv.py:178-183 (the author's NOTE at :176 says this
    block is what one would loop over).  Same ATen ops as the reference on each block, so it is
    bitwise the reference on CPU while fitting in memory for 100k points."""
    chunks = []
    for s in range(0, out_pts.shape[0], block):
        locs = out_pts[s:s + block].unsqueeze(1) - in_pts.unsqueeze(0)          # :178
        bump_arg = torch.linalg.vector_norm(locs, dim=(2), keepdims=True)       # :153-154, :180
        tf_vec = (bump_arg <= decay_param).squeeze(2)                           # :182
        idx = torch.nonzero(tf_vec, as_tuple=False)                             # :183
        idx[:, 0] += s
        chunks.append(idx)
    return torch.cat(chunks, dim=0)


# ---------------------------------------------------------------------------------------------
# filter MLP + integral
# ---------------------------------------------------------------------------------------------

def _rnd_bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to the nearest bf16 value (ties to even), keeping the dtype; straight-through for autograd (the
    gradient passes unrounded -- routing it through the casts would round the gradient to bf16 as well)."""
    r = t.detach().to(torch.float32).to(torch.bfloat16).to(t.dtype)
    return t + (r - t.detach())


def filter_mlp(weights, z: torch.Tensor, operand_bf16: bool = False) -> torch.Tensor:
    """quadconv.py:130-145: Linear(bias=False) at even positions, Sin (utils/misc.py:18-19)
    between, no activation after the last Linear.  `weights` is the list filter.{0,2,..}.weight.
    operand_bf16: the 'bf16 filter MLP' mode of the product (not in the reference): every Linear takes
    bf16-rounded weights and inputs, accumulates in the working precision."""
    h = z
    for l, W in enumerate(weights):
        if operand_bf16:
            h, W = _rnd_bf16(h), _rnd_bf16(W)
        h = torch.nn.functional.linear(h, W)
        if l + 1 < len(weights):
            h = torch.sin(h)
    return h


def quadconv_forward(out_pts, in_pts, quad_weights, filt_weights, features, eval_indices,
                     in_channels, out_channels, bias=None, block_pairs: int = 1 << 18, operand_bf16: bool = False):
    """Restatement of QuadConv.forward + _integrate (quadconv.py:238-265, :213-227), blocked over
    pairs so that the materialised [pairs, Ci, Co] filter tensor fits in memory at any mesh size.
    Differentiable w.r.t. quad_weights, filt_weights, features, bias through torch autograd, like
    the reference (which has no hand-written backward).

        out[b,j,o] = sum_{p: o(p)=o} w[i(p)] * sum_i G(y_o - x_i(p))[i,j] * f[b,i,i(p)]  (+ bias)
    """
    B = features.shape[0]
    No = out_pts.shape[0]
    integral = features.new_zeros(B, out_channels, No)                          # :222
    P = eval_indices.shape[0]
    for s in range(0, P, block_pairs):
        idx = eval_indices[s:s + block_pairs]
        w = quad_weights[idx[:, 1]]                                             # :247
        eval_locs = (out_pts[idx[:, 0]] - in_pts[idx[:, 1]]).to(features.dtype)  # :250-253 (fp32 subtract, as the reference)
        filters = filter_mlp(filt_weights, eval_locs, operand_bf16).reshape(-1, in_channels, out_channels)  # :257,:91
        values = torch.einsum('n, nij, bin -> bjn', w, filters, features[:, :, idx[:, 1]])    # :216-219
        integral = integral.scatter_add(2, idx[:, 0].expand(B, out_channels, -1), values)    # :225
    if bias is not None:
        integral = integral + bias                                              # :262-263
    return integral


def quadconv_forward_bf16tc(out_pts, in_pts, quad_weights, filt_weights, features, eval_indices,
                            in_channels, out_channels, bias=None, block_pairs: int = 2048):
    """The product's reduced-precision mode (QuadConv(compute_dtype=torch.bfloat16), csrc/contract_mma.cu) restated
    with the SAME operand roundings, everything else in float64: it is quadconv.py:238-265 in the factorised
    association order
        phi[p,h] = w[i(p)] * (last hidden activation of the filter MLP at y_o(p) - x_i(p))      fp32 -> bf16
        T[o,b,i,h] = sum_{p in N(o)} phi[p,h] * bf16(f[b,i,i(p)])                                -> bf16
        out[b,j,o] = sum_{i,h} T[o,b,i,h] * bf16(W_last[i*Co+j, h])   (+ bias)
    Not in the reference (it has one precision); used to bound the mode's own arithmetic."""
    def rb(t):
        return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)
    B = features.shape[0]
    No = out_pts.shape[0]
    H = filt_weights[-1].shape[1]
    fb = rb(features)
    T = torch.zeros(No, B, in_channels, H, dtype=torch.float64)
    P = eval_indices.shape[0]
    for s in range(0, P, block_pairs):
        idx = eval_indices[s:s + block_pairs]
        h = (out_pts[idx[:, 0]] - in_pts[idx[:, 1]]).to(torch.float64)          # :250-253 (fp32 subtract)
        for W in filt_weights[:-1]:
            h = torch.sin(torch.nn.functional.linear(h, W.to(torch.float64)))     # :130-145 hidden layers
        phi = rb(quad_weights[idx[:, 1]].to(torch.float64).unsqueeze(1) * h)      # :247 folded into phi
        contrib = torch.einsum('ph,bip->pbih', phi, fb[:, :, idx[:, 1]])
        T.index_add_(0, idx[:, 0], contrib)
    Wb = rb(filt_weights[-1]).view(in_channels, out_channels, H)
    out = torch.einsum('obih,ijh->bjo', rb(T), Wb)
    if bias is not None:
        out = out + bias.to(torch.float64)
    return out


# ---------------------------------------------------------------------------------------------
# learned quadrature weights
# ---------------------------------------------------------------------------------------------

def weight_map(points: torch.Tensor, adjacency: torch.Tensor, net_params, activation=torch.sigmoid) -> torch.Tensor:
    """mesh_handler.py:101-112: per-simplex MLP  Linear(el*dim,8) s Linear(8,8) s Linear(8,8) s Linear(8,el) s
    (mesh_handler.py:90-99, s = weight_activation) on the simplex's vertex coordinates, scatter_add of the
    el outputs onto the vertices.  net_params = [W0,b0,W1,b1,W2,b2,W3,b3] in the dtype to evaluate in."""
    el = adjacency.shape[1]
    h = points[adjacency].reshape(-1, el * points.shape[1]).to(net_params[0].dtype)          # :104
    for l in range(4):
        h = activation(torch.nn.functional.linear(h, net_params[2 * l], net_params[2 * l + 1]))   # :106
    w = torch.zeros(points.shape[0], dtype=h.dtype)                                          # :108
    return w.scatter_add(0, adjacency.reshape(-1), h.reshape(-1))                            # :110


# ---------------------------------------------------------------------------------------------
# Mesh_MaxPool
# ---------------------------------------------------------------------------------------------

def pool_group_index(elim_map: dict, n_fine: int) -> np.ndarray:
    """mesh_pool.py:40-49 (non-adjoint): group[k] = i for every k in elim_map[key_i] + [key_i],
    iterating keys in dict order; later keys overwrite earlier ones (last keeper wins)."""
    group = np.zeros(n_fine, dtype=np.int64)
    for i, (key, members) in enumerate(elim_map.items()):
        for k in list(members) + [key]:
            group[int(k)] = i
    return group


def pool_adjoint_index(elim_map: dict, n_fine: int) -> np.ndarray:
    """mesh_pool.py:55-78 (adjoint): fine point -> position of the LAST keeper (dict order) whose
    group contains it."""
    new_ind = {key: i for i, key in enumerate(elim_map)}
    index = np.zeros(n_fine, dtype=np.int64)
    inv = {}
    for key, members in elim_map.items():
        for v in list(members) + [key]:
            inv.setdefault(int(v), []).append(key)
    for fine, keepers in inv.items():
        for kk in keepers:
            index[fine] = new_ind[kk]
    return index


def maxpool_forward(x: torch.Tensor, group: torch.Tensor, n_out: int) -> torch.Tensor:
    """mesh_pool.py:88-89."""
    out = torch.zeros(x.shape[0], x.shape[1], n_out, dtype=x.dtype)   # (fp32 in the reference; float64 when used as the exact checker)
    index = group.view(1, 1, -1).expand_as(x)
    return out.scatter_reduce(dim=2, index=index, src=x, reduce='amax', include_self=False)


def maxpool_adjoint(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """mesh_pool.py:93."""
    return torch.gather(x, 2, index.view(1, 1, -1).expand(x.shape[0], x.shape[1], -1))


def segmax_c(x: np.ndarray, group: np.ndarray, n_out: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    group = np.ascontiguousarray(group, dtype=np.int64)
    B, C, Nin = x.shape
    out = np.empty((B, C, n_out), dtype=np.float32)
    _clib().qc_oracle_segmax(_fptr(x), _fptr(group), ctypes.c_int64(B * C), ctypes.c_int64(Nin),
                             ctypes.c_int64(n_out), _fptr(out))
    return out
