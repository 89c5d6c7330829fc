"""
torch.library custom ops over the C ABI (include/quadconv_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and autograd bookkeeping; every
FLOP of the path runs in libquadconv_b200.so.  Ops (namespace `quadconv_b200`):

    build_pairs     neighbour search           <- QuadConv._compute_eval_indices, quadconv.py:165-199
    pairs_to_tables tables from a given list   <- cached eval_indices Parameter, quadconv.py:186-188
    qc_forward      fused layer forward        <- QuadConv.forward/_integrate, quadconv.py:238-265,:213-227
    qc_backward     fused layer backward       <- autograd of the above (the reference has none)
    segmax / gather_pool (+ backward)          <- Mesh_MaxPool.forward, mesh_pool.py:88-93
"""
from __future__ import annotations

import functools
import math
import os

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi as C

Tensor = torch.Tensor


def _on_device(fn):
    """Run an op body with the CUDA device of its first tensor argument current: libquadconv launches on the
    current device's current stream (C.stream()), so a model on cuda:1 must not launch on cuda:0."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if not a.is_cuda:
                    raise RuntimeError("pytorch_quadconv_b200 kernels need CUDA tensors (there is no CPU path)")
                if a.device.index == torch.cuda.current_device():     # the usual case: no context switch needed
                    return fn(*args, **kwargs)
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _call(op, *args):
    """Run a quadconv_b200:: op from the autograd glue.  Eager mode calls the op's Python body directly: the
    torch.library dispatcher costs 20-25 us per call on the host, a tenth of a small-mesh layer step, and adds nothing
    here (autograd is handled by the enclosing Function, the body is CUDA-only).  Under torch.compile the registered op
    (torch.ops.quadconv_b200.*, with its fake kernel) is what gets traced."""
    if torch.compiler.is_compiling():
        return op(*args)
    return op._init_fn(*args)


def _hp(h: int) -> int:
    for v in (4, 8, 16, 32):
        if h <= v:
            return v
    raise RuntimeError(f"filter MLP last hidden width {h} > 32 is not implemented by the sm_100a kernels")


def _bp(B: int) -> int:
    bp = 1
    while bp < B and bp < 32:
        bp <<= 1
    return bp


def _nblk(B: int) -> int:
    return (B + 31) // 32


def _packed_shape(B: int, Cc: int, N: int):
    """Packed feature layout [nblk][N][CP4/4][Bp][4] (csrc/common.cuh)."""
    return (_nblk(B), N, ((Cc + 3) // 4), _bp(B), 4)


def _f32c(t: Tensor, name: str) -> Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (got {t.dtype}); the fp32 path is the only one built")
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# pair tables
# ---------------------------------------------------------------------------------------------

def _scan(counts: Tensor) -> Tensor:
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int32, device=counts.device)
    C.check(C.lib().qc_exclusive_scan_i32(C.ptr(counts), n, C.ptr(out), C.stream()), "qc_exclusive_scan_i32")
    return out


def _csc(pair_row: Tensor, pair_col: Tensor, Ni: int) -> Tuple[Tensor, Tensor, Tensor]:
    P = pair_row.numel()
    dev = pair_row.device
    counts = torch.empty(Ni, dtype=torch.int32, device=dev)
    C.check(C.lib().qc_csc_count(C.ptr(pair_col), P, Ni, C.ptr(counts), C.stream()), "qc_csc_count")
    csc_ptr = _scan(counts)
    cursor = torch.empty(Ni, dtype=torch.int32, device=dev)
    csc_pair = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    csc_row = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    C.check(C.lib().qc_csc_fill(C.ptr(pair_row), C.ptr(pair_col), P, Ni, C.ptr(csc_ptr), C.ptr(cursor),
                                C.ptr(csc_pair), C.ptr(csc_row), C.stream()), "qc_csc_fill")
    return csc_ptr, csc_pair, csc_row


@torch.library.custom_op("quadconv_b200::build_pairs", mutates_args=())
@_on_device
def build_pairs(out_pts: Tensor, in_pts: Tensor, radius: float, same: bool) -> List[Tensor]:
    """Grid-binned neighbour search.  Returns
    [eval_indices int64 [P,2], pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in,
     cell_out, cell_in]: the last two are the plain cell-sorted orders (16 consecutive entries = spatial neighbours,
    the tiles of the tensor-core path), order_out/in the same re-sorted by segment length inside 2048-point windows
    (the processing order of the fp32 kernels)."""
    out_pts = _f32c(out_pts, "output points")
    in_pts = _f32c(in_pts, "input points")
    if same:
        out_pts = in_pts
    No, d = out_pts.shape
    Ni = in_pts.shape[0]
    dev = in_pts.device
    L = C.lib()
    r32 = float(np.float32(radius))  # quadconv.py:182: python scalar demoted to fp32
    ws_bytes = L.qc_search_workspace_bytes(No, Ni, d)
    if ws_bytes == 0:
        raise RuntimeError(f"neighbour search supports spatial_dim 1..3 (got {d})")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    counts = torch.empty(No, dtype=torch.int32, device=dev)
    order_out = torch.empty(No, dtype=torch.int32, device=dev)
    order_in = torch.empty(Ni, dtype=torch.int32, device=dev)
    C.check(L.qc_search_count(C.ptr(out_pts), C.ptr(in_pts), No, Ni, d, r32, C.ptr(ws), ws_bytes,
                              C.ptr(counts), C.ptr(order_out), C.ptr(order_in), C.stream()), "qc_search_count")
    row_ptr = _scan(counts)
    P = int(row_ptr[-1].item())  # data-dependent size, like torch.nonzero in the reference
    ev = torch.empty((P, 2), dtype=torch.int64, device=dev)
    pair_row = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    pair_col = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    if P > 0:
        C.check(L.qc_search_fill(C.ptr(out_pts), C.ptr(in_pts), No, Ni, d, r32, C.ptr(ws), ws_bytes,
                                 C.ptr(row_ptr), C.ptr(ev), C.ptr(pair_row), C.ptr(pair_col), C.stream()),
                "qc_search_fill")
    pair_row, pair_col = pair_row[:P], pair_col[:P]
    csc_ptr, csc_pair, csc_row = _csc(pair_row, pair_col, Ni)
    # within windows of the cell-sorted orders, process segments of similar length together
    bal_out, bal_in = torch.empty_like(order_out), torch.empty_like(order_in)
    C.check(L.qc_balance_order(C.ptr(order_out), C.ptr(row_ptr), No, C.ptr(bal_out), C.stream()), "qc_balance_order")
    C.check(L.qc_balance_order(C.ptr(order_in), C.ptr(csc_ptr), Ni, C.ptr(bal_in), C.stream()), "qc_balance_order")
    return [ev, pair_row, pair_col, row_ptr, csc_ptr, csc_pair[:P], csc_row[:P], bal_out, bal_in, order_out, order_in]


@build_pairs.register_fake
def _(out_pts, in_pts, radius, same):
    ctx = torch.library.get_ctx()
    P = ctx.new_dynamic_size()
    No, Ni = out_pts.shape[0], in_pts.shape[0]
    i32 = dict(dtype=torch.int32, device=in_pts.device)
    return [in_pts.new_empty((P, 2), dtype=torch.int64), in_pts.new_empty(P, **i32), in_pts.new_empty(P, **i32),
            in_pts.new_empty(No + 1, **i32), in_pts.new_empty(Ni + 1, **i32), in_pts.new_empty(P, **i32),
            in_pts.new_empty(P, **i32), in_pts.new_empty(No, **i32), in_pts.new_empty(Ni, **i32),
            in_pts.new_empty(No, **i32), in_pts.new_empty(Ni, **i32)]


@torch.library.custom_op("quadconv_b200::pairs_to_tables", mutates_args=())
@_on_device
def pairs_to_tables(eval_indices: Tensor, n_out: int, n_in: int) -> List[Tensor]:
    """Tables for a pair list supplied by the caller (e.g. a loaded `eval_indices` Parameter).
    Returns [pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row]."""
    if eval_indices.dtype != torch.int64 or eval_indices.dim() != 2 or eval_indices.shape[1] != 2:
        raise RuntimeError("eval_indices must be int64 [P,2]")
    ev = eval_indices.contiguous()
    P = ev.shape[0]
    dev = ev.device
    counts = torch.empty(n_out, dtype=torch.int32, device=dev)
    pair_row = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    pair_col = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    C.check(C.lib().qc_pairs_to_csr(C.ptr(ev), P, n_out, n_in, C.ptr(counts), C.ptr(pair_row), C.ptr(pair_col),
                                    C.ptr(err), C.stream()), "qc_pairs_to_csr")
    if int(err.item()) != 0:
        raise RuntimeError("eval_indices must be sorted by output index and within range")
    row_ptr = _scan(counts)
    pair_row, pair_col = pair_row[:P], pair_col[:P]
    csc_ptr, csc_pair, csc_row = _csc(pair_row, pair_col, n_in)
    return [pair_row, pair_col, row_ptr, csc_ptr, csc_pair[:P], csc_row[:P]]


@pairs_to_tables.register_fake
def _(eval_indices, n_out, n_in):
    P = eval_indices.shape[0]
    i32 = dict(dtype=torch.int32, device=eval_indices.device)
    e = eval_indices.new_empty
    return [e(P, **i32), e(P, **i32), e(n_out + 1, **i32), e(n_in + 1, **i32), e(P, **i32), e(P, **i32)]


# ---------------------------------------------------------------------------------------------
# fused QuadConv layer
# ---------------------------------------------------------------------------------------------

def _flat_layout(shapes):
    """Offsets of tensors of the given shapes inside one flat fp32 buffer (16-byte aligned pieces)."""
    sizes = [math.prod(sh) for sh in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) & ~3
    return offs, sizes, max(total, 1)


def _carve(flat: Tensor, shapes):
    """Views of `flat` with the given shapes (see _flat_layout).  Custom ops may not return tensors that alias one
    another, so the ops below return the ONE zero-filled buffer their kernels accumulated into (one fill launch
    instead of one per gradient) and the autograd glue carves it."""
    offs, sizes, _ = _flat_layout(shapes)
    return [flat[o:o + n].view(sh) for o, n, sh in zip(offs, sizes, shapes)]


def _grad_shapes(hidden_ws, Ni, Ci, Co, H):
    return [tuple(w.shape) for w in hidden_ws] + [(Ni,), (Ci * Co, H)]


def _pair_phi_bwd(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, mlp, HP, dphi, nsl, dptrs, d_quad_w, csc_ptr, csc_pair,
                  Ni, st):
    """Per-pair chain rule through the hidden filter layers (qc_pair_phi_bwd_ws): the block partial sums and the per-pair
    quadrature-weight terms go through a workspace and are added in fixed orders, so the parameter gradients of the
    fp32 path are bit-reproducible run to run."""
    import ctypes
    L = C.lib()
    wsb = L.qc_pair_phi_bwd_workspace_bytes(P, d, ctypes.byref(mlp), HP)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dphi.device)
    C.check(L.qc_pair_phi_bwd_ws(C.ptr(out_pts), C.ptr(in_pts), C.ptr(quad_w), C.ptr(pair_row), C.ptr(pair_col), P, d, mlp, HP,
                                 C.ptr(dphi), nsl, dptrs, C.ptr(d_quad_w), C.ptr(csc_ptr), C.ptr(csc_pair), Ni, C.ptr(ws), wsb,
                                 st), "qc_pair_phi_bwd_ws")


def _pack(x: Tensor) -> Tensor:
    B, Cc, N = x.shape
    xp = torch.empty(_packed_shape(B, Cc, N), dtype=torch.float32, device=x.device)
    C.check(C.lib().qc_pack_features(C.ptr(x), B, Cc, N, C.ptr(xp), C.stream()), "qc_pack_features")
    return xp


def _unpack(xp: Tensor, B: int, Cc: int, N: int, bias: Optional[Tensor]) -> Tensor:
    x = torch.empty((B, Cc, N), dtype=torch.float32, device=xp.device)
    C.check(C.lib().qc_unpack_features(C.ptr(xp), B, Cc, N, C.ptr(bias), C.ptr(x), C.stream()), "qc_unpack_features")
    return x


_WIDE_MAX_BYTES = 16 << 30   # largest materialised intermediate the wide path may allocate


def _use_wide(Ci: int, Co: int, H: int, HP: int, B: int, n_rows: int) -> bool:
    """Wide layers (H > 8 or more than 32 channels on a dense side) do not fit the fused tcgen05
    kernels' shared-memory / TMEM budget: their gather stages run as qc_gather_accumulate /
    qc_gather_dot with the intermediate T in HBM and the dense stage as fp32 GEMMs (csrc/contract_wide.cu)."""
    if os.environ.get("QC_NO_WIDE"):
        return False
    if HP <= 8 and max(Ci, Co) <= 32:
        return False
    L = C.lib()
    if not (L.qc_gather_supported(Ci, H, HP, B) and L.qc_gather_supported(Co, H, HP, B)):
        return False
    return B * n_rows * max(Ci, Co) * H * 4 <= _WIDE_MAX_BYTES


def _gemm_nt(A: Tensor, Bm: Tensor, rows_per_batch: int = 0) -> Tensor:
    """D = A @ Bm^T for row-major A [M,K], Bm [N,K] on the tensor cores in 3xTF32 (csrc/gemm_tc.cu).
    rows_per_batch = R > 0 returns [M/R, N, R] (each block of R rows transposed), else [M, N]."""
    M, K = A.shape
    N = Bm.shape[0]
    if K % 4 != 0:      # 128-bit row pieces: pad the reduction dimension with zeros (odd channel counts only)
        A = torch.nn.functional.pad(A, (0, 4 - K % 4))
        Bm = torch.nn.functional.pad(Bm, (0, 4 - K % 4))
        K = A.shape[1]
    A, Bm = A.contiguous(), Bm.contiguous()
    shape = (M // rows_per_batch, N, rows_per_batch) if rows_per_batch else (M, N)
    D = torch.empty(shape, dtype=torch.float32, device=A.device)
    C.check(C.lib().qc_gemm_nt_tf32x3(C.ptr(A), C.ptr(Bm), C.ptr(D), M, N, K, rows_per_batch, C.stream()),
            "qc_gemm_nt_tf32x3")
    return D


def _gemm_tn(T: Tensor, X: Tensor) -> Tensor:
    """D[n][m] = sum_{b,k} X[b][n][k] * T[b*Kb + k][m] for T [nb*Kb, M] row-major and X [nb, N, Kb] (the [B, C, N]
    feature layout) on the tensor cores in 3xTF32 -- the GEMM that reduces over the row index of T (csrc/gemm_tc.cu)."""
    nb, N, Kb = X.shape
    M = T.shape[1]
    T, X = T.contiguous(), X.contiguous()
    L = C.lib()
    wsb = L.qc_gemm_tn_workspace_bytes(nb, Kb, M, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device=T.device) if wsb else None
    D = torch.empty((N, M), dtype=torch.float32, device=T.device)
    C.check(L.qc_gemm_tn_tf32x3(C.ptr(T), C.ptr(X), C.ptr(D), nb, Kb, M, N, C.ptr(ws), wsb, C.stream()),
            "qc_gemm_tn_tf32x3")
    return D


@torch.library.custom_op("quadconv_b200::qc_forward", mutates_args=())
@_on_device
def qc_forward(features: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor,
               hidden_ws: Sequence[Tensor], w_last: Tensor, bias: Optional[Tensor],
               pair_row: Tensor, pair_col: Tensor, row_ptr: Tensor, order_out: Optional[Tensor],
               in_channels: int, out_channels: int, mlp_bf16: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (out [B,Co,No], phi [P,HP], packed features) -- the last two are saved for backward.
    mlp_bf16: every Linear of the filter MLP takes bf16 operands (weights and layer inputs rounded to nearest
    even), fp32 accumulate -- the 'bf16 filter MLP' mode; everything else stays fp32."""
    features = _f32c(features, "features")
    out_pts, in_pts, quad_w = _f32c(out_pts, "points"), _f32c(in_pts, "points"), _f32c(quad_w, "weights")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    if mlp_bf16:
        w_last = w_last.to(torch.bfloat16).to(torch.float32)
    B, Ci, Ni = features.shape
    No, d = out_pts.shape
    Co = out_channels
    if Ci != in_channels or in_pts.shape[0] != Ni:
        raise RuntimeError(f"features {tuple(features.shape)} do not match in_channels={in_channels}, in_points={in_pts.shape[0]}")
    H = w_last.shape[1]
    if w_last.shape[0] != Ci * Co:
        raise RuntimeError("last filter layer must have in_channels*out_channels outputs")
    HP = _hp(H)
    P = pair_row.numel()
    dev = features.device
    L = C.lib()
    st = C.stream()
    mlp = C.make_mlp(d, hidden_ws, mlp_bf16)
    phi = torch.empty((max(P, 1), HP), dtype=torch.float32, device=dev)
    C.check(L.qc_pair_phi_fwd(C.ptr(out_pts), C.ptr(in_pts), C.ptr(quad_w), C.ptr(pair_row), C.ptr(pair_col),
                              P, d, mlp, HP, C.ptr(phi), st), "qc_pair_phi_fwd")
    xp = _pack(features)
    if P > 0 and _use_wide(Ci, Co, H, HP, B, max(No, Ni)):
        T = torch.empty((B, No, Ci * H), dtype=torch.float32, device=dev)
        C.check(L.qc_gather_accumulate(C.ptr(xp), Ni, Ci, C.ptr(row_ptr), C.ptr(pair_col), None, C.ptr(phi), H, HP,
                                       C.ptr(order_out), No, B, C.ptr(T), st), "qc_gather_accumulate")
        W2T = w_last.view(Ci, Co, H).permute(1, 0, 2).reshape(Co, Ci * H)     # [j, (i,h)]
        out = _gemm_nt(T.view(B * No, Ci * H), W2T, No)                       # [B, Co, No]
        if bias is not None:
            out = out + _f32c(bias, "bias").reshape(1, Co, No)
        return out, phi, xp
    CoP = (Co + 3) & ~3
    W2 = torch.empty((Ci * H, CoP), dtype=torch.float32, device=dev)
    C.check(L.qc_repack_wlast(C.ptr(w_last), Ci, Co, H, HP, 0, C.ptr(W2), st), "qc_repack_wlast")
    yp = torch.empty(_packed_shape(B, Co, No), dtype=torch.float32, device=dev)
    C.check(L.qc_contract(C.ptr(xp), Ni, Ci, C.ptr(row_ptr), C.ptr(pair_col), None, C.ptr(phi), H, HP,
                          C.ptr(W2), Co, C.ptr(order_out), No, B, C.ptr(yp), None, None, None, 0, st), "qc_contract")
    b2 = None
    if bias is not None:
        b2 = _f32c(bias, "bias").reshape(Co, No)
    out = _unpack(yp, B, Co, No, b2)
    return out, phi, xp


# ---- reduced-precision tensor-core path (csrc/contract_mma.cu) ---------------------------------

TC_TILE = 16            # destination points per tile (16 points x 8 filter components = the 128 rows of one MMA)
_TC_MAX_SMEM = 227 * 1024


def build_union_tables(seg_ptr: Tensor, src_idx: Tensor, order: Optional[Tensor], n_dst: int, n_src: int,
                       phi_idx: Optional[Tensor] = None):
    """Tile tables of the tensor-core path, built once per pair list (setup time, like the CSR/CSC tables):
    destination points are taken 16 at a time in processing order; for every tile the sorted union of its
    pairs' source points (padded with -1 to a multiple of 32 slots, at least 32) and the tile's pairs as
    entries {phi row, (point in tile << 16) | union slot}.  `phi_idx` maps a segment entry to its phi row
    (the CSC view; identity when None).
    Returns (tile_ent int32 [P,2], tile_pptr [ntiles+1], tile_uptr [ntiles+1], union_idx [total], max_ku)."""
    dev = seg_ptr.device
    ntiles = (n_dst + TC_TILE - 1) // TC_TILE
    i32 = dict(dtype=torch.int32, device=dev)
    P = int(src_idx.numel())
    rank = torch.arange(n_dst, device=dev)
    rank_of_dst = torch.empty(n_dst, dtype=torch.int64, device=dev)
    if order is None:
        rank_of_dst = rank
    else:
        rank_of_dst[order.long()] = rank
    seg_len = (seg_ptr[1:] - seg_ptr[:-1]).long()
    prank = torch.repeat_interleave(rank_of_dst, seg_len, output_size=P)     # processing rank of every pair's destination
    ptile = prank // TC_TILE
    key = ptile * n_src + src_idx.long()
    uniq, inv = torch.unique(key, return_inverse=True)          # sorted: by tile, then by source index
    utile = uniq // n_src
    cnt = torch.bincount(utile, minlength=ntiles)
    start = torch.cumsum(cnt, 0) - cnt
    padded = torch.clamp((cnt + 31) // 32 * 32, min=32)
    uptr = torch.zeros(ntiles + 1, dtype=torch.int64, device=dev)
    torch.cumsum(padded, 0, out=uptr[1:])
    slot_u = torch.arange(uniq.numel(), device=dev) - start[utile]
    total, max_ku = (int(v) for v in torch.stack([uptr[-1], padded.max()]).tolist())
    union_idx = torch.full((total,), -1, **i32)
    union_idx[uptr[utile] + slot_u] = (uniq % n_src).to(torch.int32)
    # the pairs grouped by tile (stable: destination rank, then segment order)
    perm = torch.argsort(prank, stable=True)
    phi_row = perm if phi_idx is None else phi_idx.long()[perm]
    meta = ((prank[perm] % TC_TILE) << 16) | slot_u[inv][perm]
    tile_ent = torch.stack([phi_row, meta], dim=1).to(torch.int32).contiguous()
    pptr = torch.zeros(ntiles + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(ptile, minlength=ntiles), 0, out=pptr[1:])
    return tile_ent, pptr.to(torch.int32), uptr.to(torch.int32), union_idx, max_ku


def tc_supported(Cs: int, Cd: int, H: int, B: int, max_ku: int) -> bool:
    """Shapes served by qc_contract_mma: H <= 8, 8..32 batch lanes per block, destination channels <= 64 (TMEM),
    operands within 227 KB of shared memory."""
    L = C.lib()
    if L.qc_mma_chunk_channels(Cs, B) == 0 or H > 8 or Cd > 64 or max_ku > 0xffff:
        return False
    CdP = max(16, (Cd + 15) // 16 * 16)
    if (TC_TILE * _bp(B) // 128) * CdP > 256:
        return False
    return 0 < L.qc_mma_smem_bytes(Cs, Cd, B, max_ku) <= _TC_MAX_SMEM


def tc_backward_supported(Ci: int, Co: int, H: int, B: int, max_ku_fwd: int, max_ku_bwd: int) -> bool:
    """The three tensor-core backward kernels (d features = qc_contract_mma on the CSC view, qc_dphi_mma, qc_dw_mma)
    all serve the shape; otherwise the backward runs the fp32 kernels."""
    L = C.lib()
    if not tc_supported(Co, Ci, H, B, max_ku_bwd):                       # d features: sources = out points / out channels
        return False
    CC = L.qc_mma_chunk_channels(Ci, B)
    if CC == 0 or _bp(B) * CC < 64 or max_ku_fwd > 384 or Co > 64 or L.qc_mma_chunk_channels(Co, B) == 0:
        return False
    if not 0 < L.qc_dphi_mma_smem_bytes(Ci, Co, B, max_ku_fwd) <= _TC_MAX_SMEM:
        return False
    nch = (Ci + CC - 1) // CC                                            # one TMEM accumulator per channel chunk
    if nch * ((Co + 15) // 16 * 16) > 256:
        return False
    return 0 < L.qc_dw_mma_smem_bytes(Ci, Co, B, max_ku_fwd) <= _TC_MAX_SMEM


def _pack_bf16(x: Tensor) -> Tensor:
    """[B,C,N] fp32 or bf16 -> bf16 packed [nblk][N][chunks][Bp][CC] (operand layout of the tensor-core kernels)."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"features must be float32 or bfloat16 (got {x.dtype})")
    x = x.contiguous()
    B, Cc, N = x.shape
    CC = C.lib().qc_mma_chunk_channels(Cc, B)
    xh = torch.empty((_nblk(B), N, (Cc + CC - 1) // CC, _bp(B), CC), dtype=torch.bfloat16, device=x.device)
    C.check(C.lib().qc_pack_features_bf16(C.ptr(x), int(x.dtype == torch.bfloat16), B, Cc, N, CC, C.ptr(xh), C.stream()),
            "qc_pack_features_bf16")
    return xh


@torch.library.custom_op("quadconv_b200::qc_forward_tc", mutates_args=())
@_on_device
def qc_forward_tc(features: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor,
                  hidden_ws: Sequence[Tensor], w_last: Tensor, bias: Optional[Tensor],
                  pair_row: Tensor, pair_col: Tensor, order_out: Optional[Tensor],
                  tile_ent: Tensor, tile_pptr: Tensor, tile_uptr: Tensor, union_idx: Tensor, max_ku: int,
                  in_channels: int, out_channels: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Forward of the reduced-precision mode: per-pair hidden layers in fp32 (qc_pair_phi_fwd), then both
    contraction stages on the tensor cores with bf16 operands (qc_contract_mma).  Returns (out, phi, bf16
    packed features).  `features` may already be bfloat16 (half the host-to-device / HBM read traffic)."""
    if features.dtype != torch.bfloat16:
        features = _f32c(features, "features")
    out_pts, in_pts, quad_w = _f32c(out_pts, "points"), _f32c(in_pts, "points"), _f32c(quad_w, "weights")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    B, Ci, Ni = features.shape
    No, d = out_pts.shape
    Co = out_channels
    H = w_last.shape[1]
    HP = _hp(H)
    P = pair_row.numel()
    L = C.lib()
    st = C.stream()
    mlp = C.make_mlp(d, hidden_ws, False)
    phi = torch.empty((max(P, 1), HP), dtype=torch.float32, device=features.device)
    C.check(L.qc_pair_phi_fwd(C.ptr(out_pts), C.ptr(in_pts), C.ptr(quad_w), C.ptr(pair_row), C.ptr(pair_col),
                              P, d, mlp, HP, C.ptr(phi), st), "qc_pair_phi_fwd")
    xh = _pack_bf16(features)
    yp = torch.empty(_packed_shape(B, Co, No), dtype=torch.float32, device=features.device)
    C.check(L.qc_contract_mma(C.ptr(xh), Ni, Ci, C.ptr(tile_ent), C.ptr(tile_pptr), C.ptr(tile_uptr),
                              C.ptr(union_idx), tile_uptr.numel() - 1, max_ku, C.ptr(phi), H, HP, C.ptr(w_last), Co, 0,
                              C.ptr(order_out), No, B, C.ptr(yp), st), "qc_contract_mma")
    b2 = None
    if bias is not None:
        b2 = _f32c(bias, "bias").reshape(Co, No)
    return _unpack(yp, B, Co, No, b2), phi, xh


@qc_forward_tc.register_fake
def _(features, out_pts, in_pts, quad_w, hidden_ws, w_last, bias, pair_row, pair_col, order_out,
      tile_ent, tile_pptr, tile_uptr, union_idx, max_ku, in_channels, out_channels):
    B, Ci, Ni = features.shape
    No = out_pts.shape[0]
    HP = _hp(w_last.shape[1])
    CC = 4
    return (features.new_empty((B, out_channels, No)), features.new_empty((pair_row.shape[0], HP)),
            features.new_empty((_nblk(B), Ni, (Ci + CC - 1) // CC, _bp(B), CC), dtype=torch.bfloat16))


@torch.library.custom_op("quadconv_b200::qc_backward_tc", mutates_args=())
@_on_device
def qc_backward_tc(grad_out: Tensor, xh: Tensor, phi: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor,
                   hidden_ws: Sequence[Tensor], w_last: Tensor, pair_row: Tensor, pair_col: Tensor,
                   order_out: Optional[Tensor], order_in: Optional[Tensor],
                   f_ent: Tensor, f_pptr: Tensor, f_uptr: Tensor, f_uidx: Tensor, f_max_ku: int,
                   b_ent: Tensor, b_pptr: Tensor, b_uptr: Tensor, b_uidx: Tensor, b_max_ku: int,
                   batch: int, in_channels: int, out_channels: int, has_bias: bool) -> List[Tensor]:
    """Backward of the reduced-precision mode, all three contractions on the tensor cores (bf16 operands, fp32
    accumulation): d(phi) (qc_dphi_mma), d W_last (qc_dw_mma), d features (qc_contract_mma on the CSC view); the
    per-pair chain rule through the hidden layers stays fp32 (qc_pair_phi_bwd).
    Returns [d_features, d_bias (or empty), flat] -- flat holds d_hidden_ws..., d_quad_w, d_w_last (_grad_shapes)."""
    import ctypes
    grad_out = _f32c(grad_out, "grad_out")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    B, Ci, Co = batch, in_channels, out_channels
    Ni, No, d = in_pts.shape[0], out_pts.shape[0], out_pts.shape[1]
    H = w_last.shape[1]
    HP = _hp(H)
    P = pair_row.numel()
    dev = grad_out.device
    L = C.lib()
    st = C.stream()
    f32 = dict(dtype=torch.float32, device=dev)
    gh = _pack_bf16(grad_out)
    d_bias = torch.empty(0, **f32)
    if has_bias:
        d_bias = torch.empty((1, Co, No), **f32)
        C.check(L.qc_batch_sum(C.ptr(grad_out), B, Co, No, C.ptr(d_bias), st), "qc_batch_sum")
    nft, nbt = f_uptr.numel() - 1, b_uptr.numel() - 1
    # d phi (forward view: tiles of out points, gathered feature rows) and the per-pair chain rule
    dphi = (torch.zeros if _nblk(B) > 1 else torch.empty)((1, max(P, 1), HP), **f32)
    C.check(L.qc_dphi_mma(C.ptr(xh), Ni, Ci, C.ptr(gh), Co, C.ptr(f_ent), C.ptr(f_pptr), C.ptr(f_uptr), C.ptr(f_uidx), nft,
                          f_max_ku, C.ptr(w_last), H, HP, C.ptr(order_out), No, B, C.ptr(dphi), st), "qc_dphi_mma")
    shapes = _grad_shapes(hidden_ws, Ni, Ci, Co, H)
    flat = torch.zeros(_flat_layout(shapes)[2], **f32)
    *d_hidden, d_quad_w, d_w_last = _carve(flat, shapes)
    mlp = C.make_mlp(d, hidden_ws, False)
    dptrs = (ctypes.c_void_p * max(len(d_hidden), 1))(*[C.ptr(t) for t in d_hidden])
    C.check(L.qc_pair_phi_bwd(C.ptr(out_pts), C.ptr(in_pts), C.ptr(quad_w), C.ptr(pair_row), C.ptr(pair_col),
                              P, d, mlp, HP, C.ptr(dphi), 1, dptrs, C.ptr(d_quad_w), st), "qc_pair_phi_bwd")
    # d W_last (same tiles, the stage-B result recomputed and contracted with the grad rows)
    wsb = L.qc_dw_mma_workspace_bytes(Ci, Co, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    C.check(L.qc_dw_mma(C.ptr(xh), Ni, Ci, C.ptr(gh), Co, C.ptr(f_ent), C.ptr(f_pptr), C.ptr(f_uptr), C.ptr(f_uidx), nft,
                        f_max_ku, C.ptr(phi), H, HP, C.ptr(order_out), No, B, C.ptr(ws), wsb, C.ptr(d_w_last), st), "qc_dw_mma")
    # d features (CSC view: tiles of in points, gathered grad rows, W_last transposed)
    dxp = torch.empty(_packed_shape(B, Ci, Ni), **f32)
    C.check(L.qc_contract_mma(C.ptr(gh), No, Co, C.ptr(b_ent), C.ptr(b_pptr), C.ptr(b_uptr), C.ptr(b_uidx), nbt, b_max_ku,
                              C.ptr(phi), H, HP, C.ptr(w_last), Ci, 1, C.ptr(order_in), Ni, B, C.ptr(dxp), st),
            "qc_contract_mma (d features)")
    d_features = _unpack(dxp, B, Ci, Ni, None)
    return [d_features, d_bias, flat]


@qc_backward_tc.register_fake
def _(grad_out, xh, phi, out_pts, in_pts, quad_w, hidden_ws, w_last, pair_row, pair_col, order_out, order_in,
      f_ent, f_pptr, f_uptr, f_uidx, f_max_ku, b_ent, b_pptr, b_uptr, b_uidx, b_max_ku, batch, in_channels,
      out_channels, has_bias):
    Ni, No = in_pts.shape[0], out_pts.shape[0]
    e = grad_out.new_empty
    shapes = _grad_shapes(hidden_ws, Ni, in_channels, out_channels, w_last.shape[1])
    return [e((batch, in_channels, Ni)), e((1, out_channels, No)) if has_bias else e(0), e(_flat_layout(shapes)[2])]


@qc_forward.register_fake
def _(features, out_pts, in_pts, quad_w, hidden_ws, w_last, bias, pair_row, pair_col, row_ptr, order_out,
      in_channels, out_channels, mlp_bf16=False):
    B, Ci, Ni = features.shape
    No = out_pts.shape[0]
    HP = _hp(w_last.shape[1])
    return (features.new_empty((B, out_channels, No)), features.new_empty((pair_row.shape[0], HP)),
            features.new_empty(_packed_shape(B, Ci, Ni)))


@torch.library.custom_op("quadconv_b200::qc_backward", mutates_args=())
@_on_device
def qc_backward(grad_out: Tensor, xp: Tensor, phi: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor,
                hidden_ws: Sequence[Tensor], w_last: Tensor, pair_row: Tensor, pair_col: Tensor,
                row_ptr: Tensor, csc_ptr: Tensor, csc_pair: Tensor, csc_row: Tensor,
                order_out: Optional[Tensor], order_in: Optional[Tensor], batch: int, in_channels: int,
                out_channels: int, has_bias: bool, mlp_bf16: bool = False) -> List[Tensor]:
    """Returns [d_features, d_bias (or empty), flat] -- flat holds d_hidden_ws..., d_quad_w, d_w_last (_grad_shapes)."""
    grad_out = _f32c(grad_out, "grad_out")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    if mlp_bf16:   # straight-through: gradients are taken at the rounded operands, applied to the fp32 weights
        w_last = w_last.to(torch.bfloat16).to(torch.float32)
    B, Ci, Co = batch, in_channels, out_channels
    Ni, No, d = in_pts.shape[0], out_pts.shape[0], out_pts.shape[1]
    H = w_last.shape[1]
    HP = _hp(H)
    P = pair_row.numel()
    dev = grad_out.device
    L = C.lib()
    st = C.stream()
    f32 = dict(dtype=torch.float32, device=dev)

    gp = _pack(grad_out)
    d_bias = torch.empty(0, **f32)
    if has_bias:
        d_bias = torch.empty((1, Co, No), **f32)
        C.check(L.qc_batch_sum(C.ptr(grad_out), B, Co, No, C.ptr(d_bias), st), "qc_batch_sum")

    if P > 0 and _use_wide(Ci, Co, H, HP, B, max(No, Ni)):
        wl = w_last.view(Ci, Co, H)
        # d phi: dT = G * W2^T as a GEMM, then the per-pair dots against the gathered features
        W2 = wl.permute(0, 2, 1).reshape(Ci * H, Co)                       # [(i,h), j]
        dT = _gemm_nt(grad_out.transpose(1, 2).reshape(B * No, Co), W2)    # [B*No, Ci*H]
        nsl = L.qc_gather_dot_num_slices(Ci, HP, B)
        dphi = torch.empty((nsl, P, HP), **f32)
        C.check(L.qc_gather_dot(C.ptr(xp), Ni, Ci, C.ptr(dT), C.ptr(row_ptr), C.ptr(pair_col), H, HP,
                                C.ptr(order_out), No, B, P, C.ptr(dphi), st), "qc_gather_dot")
        del dT
        d_hidden = [torch.zeros_like(w) for w in hidden_ws]
        d_quad_w = torch.zeros(Ni, **f32)
        mlp = C.make_mlp(d, hidden_ws, mlp_bf16)
        import ctypes
        dptrs = (ctypes.c_void_p * max(len(d_hidden), 1))(*[C.ptr(t) for t in d_hidden])
        _pair_phi_bwd(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, mlp, HP, dphi, nsl, dptrs, d_quad_w, csc_ptr, csc_pair,
                      Ni, st)
        del dphi
        # d features and d W_last from T'[b][n][(j,h)] = sum_p phi[p][h] G[out(p)][b][j] (CSC view)
        Tp = torch.empty((B, Ni, Co * H), **f32)
        C.check(L.qc_gather_accumulate(C.ptr(gp), No, Co, C.ptr(csc_ptr), C.ptr(csc_row), C.ptr(csc_pair),
                                       C.ptr(phi), H, HP, C.ptr(order_in), Ni, B, C.ptr(Tp), st),
                "qc_gather_accumulate (backward)")
        d_features = _gemm_nt(Tp.view(B * Ni, Co * H), wl.reshape(Ci, Co * H), Ni)   # [B, Ci, Ni]
        x = _unpack(xp, B, Ci, Ni, None)
        dWt = _gemm_tn(Tp.view(B * Ni, Co * H), x)                                    # [i, (j,h)]
        d_w_last = dWt.reshape(Ci * Co, H)
        shapes = _grad_shapes(hidden_ws, Ni, Ci, Co, H)
        flat = torch.empty(_flat_layout(shapes)[2], **f32)
        for dst, src in zip(_carve(flat, shapes), d_hidden + [d_quad_w, d_w_last]):
            dst.copy_(src)
        return [d_features, d_bias, flat]

    # d phi (grouped by out point), then back through the hidden layers per pair
    W3 = torch.empty((Co, Ci, HP), **f32)
    C.check(L.qc_repack_wlast(C.ptr(w_last), Ci, Co, H, HP, 2, C.ptr(W3), st), "qc_repack_wlast")
    nsl = L.qc_dphi_num_slices(Ni, Ci, Co, HP, B)
    dphi = torch.empty((nsl, max(P, 1), HP), **f32)
    C.check(L.qc_dphi(C.ptr(xp), Ni, Ci, C.ptr(gp), Co, C.ptr(row_ptr), C.ptr(pair_col), C.ptr(W3), H, HP,
                      C.ptr(order_out), No, B, P, C.ptr(dphi), st), "qc_dphi")
    # every accumulated output of this step lives in ONE zero-filled buffer (one fill launch instead of five)
    CiP = (Ci + 3) & ~3
    shapes = _grad_shapes(hidden_ws, Ni, Ci, Co, H) + [(Co * H, CiP)]
    flat = torch.zeros(_flat_layout(shapes)[2], **f32)
    *d_hidden, d_quad_w, d_w_last, dW2 = _carve(flat, shapes)
    mlp = C.make_mlp(d, hidden_ws, mlp_bf16)
    import ctypes
    dptrs = (ctypes.c_void_p * max(len(d_hidden), 1))(*[C.ptr(t) for t in d_hidden])
    if P > 0:
        _pair_phi_bwd(out_pts, in_pts, quad_w, pair_row, pair_col, P, d, mlp, HP, dphi, nsl, dptrs, d_quad_w, csc_ptr, csc_pair,
                      Ni, st)

    # d features (grouped by in point through the CSC tables) + d W_last in the same pass
    W2t = torch.empty((Co * H, CiP), **f32)
    C.check(L.qc_repack_wlast(C.ptr(w_last), Ci, Co, H, HP, 1, C.ptr(W2t), st), "qc_repack_wlast")
    dxp = torch.empty(_packed_shape(B, Ci, Ni), **f32)
    ws_bytes = L.qc_contract_workspace_bytes(Co, H, HP, Ci, Ni, B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    C.check(L.qc_contract(C.ptr(gp), No, Co, C.ptr(csc_ptr), C.ptr(csc_row), C.ptr(csc_pair), C.ptr(phi), H, HP,
                          C.ptr(W2t), Ci, C.ptr(order_in), Ni, B, C.ptr(dxp), C.ptr(xp), C.ptr(dW2),
                          C.ptr(ws), ws_bytes, st),
            "qc_contract (backward)")
    C.check(L.qc_unpack_dwlast(C.ptr(dW2), Ci, Co, H, C.ptr(d_w_last), st), "qc_unpack_dwlast")
    d_features = _unpack(dxp, B, Ci, Ni, None)
    return [d_features, d_bias, flat]


@qc_backward.register_fake
def _(grad_out, xp, phi, out_pts, in_pts, quad_w, hidden_ws, w_last, pair_row, pair_col, row_ptr, csc_ptr,
      csc_pair, csc_row, order_out, order_in, batch, in_channels, out_channels, has_bias, mlp_bf16=False):
    Ni, No = in_pts.shape[0], out_pts.shape[0]
    e = grad_out.new_empty
    H = w_last.shape[1]
    CiP = (in_channels + 3) & ~3
    shapes = _grad_shapes(hidden_ws, Ni, in_channels, out_channels, H) + [(out_channels * H, CiP)]
    return [e((batch, in_channels, Ni)), e((1, out_channels, No)) if has_bias else e(0), e(_flat_layout(shapes)[2])]


# ---- the whole layer in ONE C call per direction (csrc/layer.cu) ------------------------------
# Small meshes are bound by the host: the composed ops above make ~6 + ~11 C calls and a dozen allocations per layer
# step.  qc_layer_forward / qc_layer_backward run the same kernels behind one call with one workspace, which also
# carries the saved activations (packed features, phi) to the backward call.  The workspace holds the backward's
# scratch too, so this path is taken only while it stays below SINGLE_CALL_MAX_BYTES; larger layers (where the host
# cost is noise) keep the composed ops and their tighter allocation lifetimes.

SINGLE_CALL_MAX_BYTES = 64 << 20


def _layer_desc(out_pts, in_pts, quad_w, hidden_ws, w_last, bias, pair_row, pair_col, row_ptr, csc_ptr, csc_pair,
                csc_row, order_out, order_in, Ci, Co) -> C.QcLayer:
    d = C.QcLayer()
    d.d, d.Ni, d.No, d.Ci, d.Co, d.H = out_pts.shape[1], in_pts.shape[0], out_pts.shape[0], Ci, Co, w_last.shape[1]
    d.P = pair_row.numel()
    d.out_pts, d.in_pts, d.quad_w = C.ptr(out_pts), C.ptr(in_pts), C.ptr(quad_w)
    d.mlp = C.make_mlp(out_pts.shape[1], hidden_ws, False)
    d.w_last, d.bias = C.ptr(w_last), C.ptr(bias)
    d.pair_row, d.pair_col, d.row_ptr = C.ptr(pair_row), C.ptr(pair_col), C.ptr(row_ptr)
    d.csc_ptr, d.csc_pair, d.csc_row = C.ptr(csc_ptr), C.ptr(csc_pair), C.ptr(csc_row)
    d.order_out, d.order_in = C.ptr(order_out), C.ptr(order_in)
    return d


@functools.lru_cache(maxsize=256)
def single_call_workspace_bytes(B, Ci, Co, H, Ni, No, P, d=2, hidden=()) -> int:
    """Workspace of qc_layer_forward/backward for a layer of this shape (only the sizes of the descriptor are read);
    `hidden` = the hidden widths of the filter MLP (the last one is H)."""
    desc = C.QcLayer()
    desc.d, desc.Ni, desc.No, desc.Ci, desc.Co, desc.H, desc.P = d, Ni, No, Ci, Co, H, P
    desc.mlp.n_hidden = len(hidden)
    desc.mlp.dims[0] = d
    for l, h in enumerate(hidden):
        desc.mlp.dims[l + 1] = h
    import ctypes
    return int(C.lib().qc_layer_workspace_bytes(ctypes.byref(desc), B))


@torch.library.custom_op("quadconv_b200::layer_forward", mutates_args=())
@_on_device
def layer_forward(features: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor, hidden_ws: Sequence[Tensor],
                  w_last: Tensor, bias: Optional[Tensor], pair_row: Tensor, pair_col: Tensor, row_ptr: Tensor,
                  csc_ptr: Tensor, csc_pair: Tensor, csc_row: Tensor, order_out: Optional[Tensor],
                  order_in: Optional[Tensor], in_channels: int, out_channels: int) -> Tuple[Tensor, Tensor]:
    """QuadConv.forward + _integrate (quadconv.py:238-265, :213-227) as ONE C call.  Returns (out [B,Co,No], ws): the
    workspace carries the saved activations and must reach layer_backward untouched."""
    import ctypes
    features = _f32c(features, "features")
    out_pts, in_pts, quad_w = _f32c(out_pts, "points"), _f32c(in_pts, "points"), _f32c(quad_w, "weights")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    B, Ci, Ni = features.shape
    if Ci != in_channels or in_pts.shape[0] != Ni:
        raise RuntimeError(f"features {tuple(features.shape)} do not match in_channels={in_channels}, in_points={in_pts.shape[0]}")
    if w_last.shape[0] != Ci * out_channels:
        raise RuntimeError("last filter layer must have in_channels*out_channels outputs")
    _hp(w_last.shape[1])
    b2 = _f32c(bias, "bias").reshape(out_channels, out_pts.shape[0]) if bias is not None else None
    desc = _layer_desc(out_pts, in_pts, quad_w, hidden_ws, w_last, b2, pair_row, pair_col, row_ptr, csc_ptr, csc_pair,
                       csc_row, order_out, order_in, Ci, out_channels)
    L = C.lib()
    ref = ctypes.byref(desc)
    wsb = L.qc_layer_workspace_bytes(ref, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=features.device)
    out = torch.empty((B, out_channels, out_pts.shape[0]), dtype=torch.float32, device=features.device)
    C.check(L.qc_layer_forward(ref, C.ptr(features), B, C.ptr(out), C.ptr(ws), wsb, C.stream()), "qc_layer_forward")
    return out, ws


@layer_forward.register_fake
def _(features, out_pts, in_pts, quad_w, hidden_ws, w_last, bias, pair_row, pair_col, row_ptr, csc_ptr, csc_pair,
      csc_row, order_out, order_in, in_channels, out_channels):
    B, Ci, Ni = features.shape
    No = out_pts.shape[0]
    wsb = single_call_workspace_bytes(B, Ci, out_channels, w_last.shape[1], Ni, No, pair_row.shape[0], out_pts.shape[1],
                                      tuple(int(w.shape[0]) for w in hidden_ws))
    return features.new_empty((B, out_channels, No)), features.new_empty(wsb, dtype=torch.uint8)


@torch.library.custom_op("quadconv_b200::layer_backward", mutates_args=("ws",))   # the workspace's scratch regions are rewritten
@_on_device
def layer_backward(grad_out: Tensor, ws: Tensor, out_pts: Tensor, in_pts: Tensor, quad_w: Tensor,
                   hidden_ws: Sequence[Tensor], w_last: Tensor, pair_row: Tensor, pair_col: Tensor, row_ptr: Tensor,
                   csc_ptr: Tensor, csc_pair: Tensor, csc_row: Tensor, order_out: Optional[Tensor],
                   order_in: Optional[Tensor], in_channels: int, out_channels: int, has_bias: bool) -> List[Tensor]:
    """Autograd of layer_forward as ONE C call.  Returns [d_features, d_bias (or empty), flat] like qc_backward."""
    import ctypes
    grad_out = _f32c(grad_out, "grad_out")
    hidden_ws = [_f32c(w, "filter weight") for w in hidden_ws]
    w_last = _f32c(w_last, "filter weight")
    B, Co, No = grad_out.shape
    Ci, Ni = in_channels, in_pts.shape[0]
    dev = grad_out.device
    f32 = dict(dtype=torch.float32, device=dev)
    desc = _layer_desc(out_pts, in_pts, quad_w, hidden_ws, w_last, None, pair_row, pair_col, row_ptr, csc_ptr, csc_pair,
                       csc_row, order_out, order_in, Ci, Co)
    shapes = _grad_shapes(hidden_ws, Ni, Ci, Co, w_last.shape[1])
    flat = torch.empty(_flat_layout(shapes)[2], **f32)          # every piece is overwritten by the C call
    *d_hidden, d_quad_w, d_w_last = _carve(flat, shapes)
    d_features = torch.empty((B, Ci, Ni), **f32)
    d_bias = torch.empty((1, Co, No), **f32) if has_bias else torch.empty(0, **f32)
    dptrs = (ctypes.c_void_p * max(len(d_hidden), 1))(*[C.ptr(t) for t in d_hidden])
    C.check(C.lib().qc_layer_backward(ctypes.byref(desc), C.ptr(grad_out), B, C.ptr(d_features), C.ptr(d_quad_w),
                                      C.ptr(d_w_last), dptrs, C.ptr(d_bias) if has_bias else None, C.ptr(ws),
                                      ws.numel(), C.stream()), "qc_layer_backward")
    return [d_features, d_bias, flat]


@layer_backward.register_fake
def _(grad_out, ws, out_pts, in_pts, quad_w, hidden_ws, w_last, pair_row, pair_col, row_ptr, csc_ptr, csc_pair,
      csc_row, order_out, order_in, in_channels, out_channels, has_bias):
    Ni, No = in_pts.shape[0], out_pts.shape[0]
    e = grad_out.new_empty
    shapes = _grad_shapes(hidden_ws, Ni, in_channels, out_channels, w_last.shape[1])
    return [e((grad_out.shape[0], in_channels, Ni)), e((1, out_channels, No)) if has_bias else e(0),
            e(_flat_layout(shapes)[2])]


def _single_call_ok(B, Ci, Co, H, Ni, No, P, d, mlp_bf16, hidden=()) -> bool:
    if mlp_bf16 or P <= 0 or SINGLE_CALL_MAX_BYTES <= 0 or H > 32:
        return False
    if _use_wide(Ci, Co, H, _hp(H), B, max(No, Ni)):
        return False
    return single_call_workspace_bytes(B, Ci, Co, H, Ni, No, P, d, hidden) <= SINGLE_CALL_MAX_BYTES


class _QuadConvFn(torch.autograd.Function):
    """Autograd glue: forward/backward are the custom ops above."""

    @staticmethod
    def forward(ctx, features, quad_w, w_last, bias, tables, dims, *hidden_ws):
        out_pts, in_pts, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in = tables[:10]
        Ci, Co, mlp_bf16, tc = dims
        if tc is not None:
            tile_ent, tile_pptr, tile_uptr, union_idx, max_ku = tc["fwd"]
            out, phi, xh = _call(qc_forward_tc, features, out_pts, in_pts, quad_w, list(hidden_ws), w_last, bias, pair_row,
                                         pair_col, tc["order_out"], tile_ent, tile_pptr, tile_uptr, union_idx, max_ku, Ci, Co)
            # the tensor-core backward reads the bf16 packed features; shapes it does not serve re-pack the fp32 input
            ctx.save_for_backward(xh if tc.get("bwd") is not None else features, phi, quad_w, w_last, *hidden_ws)
        elif _single_call_ok(features.shape[0], Ci, Co, w_last.shape[1], in_pts.shape[0], out_pts.shape[0],
                             pair_row.numel(), out_pts.shape[1], mlp_bf16, tuple(int(w.shape[0]) for w in hidden_ws)):
            out, ws = _call(layer_forward, features, out_pts, in_pts, quad_w, list(hidden_ws), w_last, bias, pair_row, pair_col,
                                    row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in, Ci, Co)
            ctx.save_for_backward(ws, quad_w, w_last, *hidden_ws)
            ctx.single = True
        else:
            out, phi, xp = _call(qc_forward, features, out_pts, in_pts, quad_w, list(hidden_ws), w_last, bias, pair_row,
                                      pair_col, row_ptr, order_out, Ci, Co, mlp_bf16)
            ctx.save_for_backward(xp, phi, quad_w, w_last, *hidden_ws)
        ctx.tables = tables
        ctx.tc = tc
        ctx.dims = (features.shape[0], Ci, Co, bias is not None, mlp_bf16)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        out_pts, in_pts, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in = ctx.tables[:10]
        B, Ci, Co, has_bias, mlp_bf16 = ctx.dims
        if getattr(ctx, "single", False):
            ws, quad_w, w_last, *hidden_ws = ctx.saved_tensors
            res = _call(layer_backward, grad_out, ws, out_pts, in_pts, quad_w, list(hidden_ws), w_last, pair_row, pair_col,
                                 row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in, Ci, Co, has_bias)
            return _QuadConvFn._unpack_grads(res, hidden_ws, in_pts.shape[0], Ci, Co, w_last.shape[1], has_bias)
        xp, phi, quad_w, w_last, *hidden_ws = ctx.saved_tensors
        if ctx.tc is not None and ctx.tc.get("bwd") is not None:
            tc = ctx.tc
            res = _call(qc_backward_tc, grad_out, xp, phi, out_pts, in_pts, quad_w, list(hidden_ws), w_last, pair_row, pair_col,
                                 tc["order_out"], tc["order_in"], *tc["fwd"], *tc["bwd"], B, Ci, Co, has_bias)
            return _QuadConvFn._unpack_grads(res, hidden_ws, in_pts.shape[0], Ci, Co, w_last.shape[1], has_bias)
        if ctx.tc is not None:
            with torch.cuda.device(xp.device):
                xp = _pack(_f32c(xp, "features"))
        res = _call(qc_backward, grad_out, xp, phi, out_pts, in_pts, quad_w, list(hidden_ws), w_last, pair_row, pair_col,
                          row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in, B, Ci, Co, has_bias, mlp_bf16)
        return _QuadConvFn._unpack_grads(res, hidden_ws, in_pts.shape[0], Ci, Co, w_last.shape[1], has_bias)

    @staticmethod
    def _unpack_grads(res, hidden_ws, Ni, Ci, Co, H, has_bias):
        d_features, d_bias, flat = res
        *d_hidden, d_quad_w, d_w_last = _carve(flat, _grad_shapes(hidden_ws, Ni, Ci, Co, H))
        return (d_features, d_quad_w, d_w_last, d_bias if has_bias else None, None, None, *d_hidden)


# ---- batch folding for layers with few channels ------------------------------------------------
# The per-pair overhead of the fused kernels (pair index, phi row, gather address, loop control) is paid once per
# (pair, batch lane) and amortised over the channels a lane owns: a 4 -> 8 channel layer runs at a quarter of the
# per-FLOP rate of a 32 -> 32 one, and with batch 256 it walks the pair list eight times (one pass per block of 32
# lanes).  A layer with few channels and a large batch is therefore evaluated as the EQUIVALENT layer with `s` batch
# entries folded into the channel axis:
#     x [B, Ci, N] viewed as [B/s, s*Ci, N]  (free),   W_last' = I_s (x) W_last  (block diagonal, [(s,i),(s',j),h]),
#     out' [B/s, s*Co, N] viewed as [B, Co, N]  (free).
# Stage B (the gather, 60 % of the work) does exactly the same sums per (batch entry, channel) with s times fewer lanes;
# the dense stages multiply by the zero blocks as well, which the tensor cores absorb.  Values are unchanged up to the
# summation order of d W_last over the s diagonal blocks.  FOLD_MIN_BATCH = the smallest folded batch (lanes per point).
FOLD_MAX_CHANNELS = 32
FOLD_MIN_BATCH = 32
_EYES = {}


def fold_factor(B: int, Ci: int, Co: int, H: int) -> int:
    if H > 8 or FOLD_MIN_BATCH <= 0:
        return 1
    f = 1
    while 2 * f * max(Ci, Co) <= FOLD_MAX_CHANNELS and B % (2 * f) == 0 and B // (2 * f) >= FOLD_MIN_BATCH:
        f *= 2
    return f


def _eye(f: int, device) -> Tensor:
    key = (f, str(device))
    e = _EYES.get(key)
    if e is None:
        e = _EYES[key] = torch.eye(f, device=device, dtype=torch.float32).view(f, 1, f, 1, 1)
    return e


def quadconv_apply(features, quad_w, hidden_ws, w_last, bias, tables, in_channels, out_channels, mlp_bf16=False,
                   tc=None):
    B, Ci, Co, H = features.shape[0], in_channels, out_channels, w_last.shape[1]
    f = 1
    if tc is None and not mlp_bf16 and features.dim() == 3 and features.shape[1] == Ci and features.dtype == torch.float32:
        f = fold_factor(B, Ci, Co, H)
    if f > 1:
        N = features.shape[2]
        No = tables[0].shape[0]
        xf = features.contiguous().view(B // f, f * Ci, N)
        wf = (_eye(f, w_last.device) * w_last.view(1, Ci, 1, Co, H)).view(f * Ci * f * Co, H)
        bf = bias.repeat(1, f, 1) if bias is not None else None
        out = _QuadConvFn.apply(xf, quad_w, wf, bf, tables, (f * Ci, f * Co, False, None), *hidden_ws)
        return out.view(B, Co, No)
    return _QuadConvFn.apply(features, quad_w, w_last, bias, tables,
                             (in_channels, out_channels, bool(mlp_bf16), tc), *hidden_ws)


# ---------------------------------------------------------------------------------------------
# learned quadrature weights (MeshHandler.weight_map, mesh_handler.py:101-112)
# ---------------------------------------------------------------------------------------------

def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[C.ptr(t) for t in tensors])


@torch.library.custom_op("quadconv_b200::weight_map", mutates_args=())
@_on_device
def weight_map(points: Tensor, adj32: Tensor, inc_ptr: Tensor, inc_idx: Tensor, params: Sequence[Tensor]) -> Tensor:
    """weights[N] = scatter_add of the per-simplex MLP outputs; params = [W0,b0,W1,b1,W2,b2,W3,b3]."""
    points = _f32c(points, "points")
    params = [_f32c(p, "weight-network parameter") for p in params]
    M, E = adj32.shape
    N, d = points.shape
    el_w = torch.empty((M, E), dtype=torch.float32, device=points.device)
    C.check(C.lib().qc_weight_map_fwd(C.ptr(points), C.ptr(adj32), M, E, d, _ptr_array(params), C.ptr(el_w), C.stream()),
            "qc_weight_map_fwd")
    w = torch.empty(N, dtype=torch.float32, device=points.device)
    C.check(C.lib().qc_vertex_sum(C.ptr(el_w), C.ptr(inc_ptr), C.ptr(inc_idx), N, C.ptr(w), C.stream()), "qc_vertex_sum")
    return w


@weight_map.register_fake
def _(points, adj32, inc_ptr, inc_idx, params):
    return points.new_empty(points.shape[0])


@torch.library.custom_op("quadconv_b200::weight_map_bwd", mutates_args=())
@_on_device
def weight_map_bwd(grad_w: Tensor, points: Tensor, adj32: Tensor, params: Sequence[Tensor]) -> Tensor:
    grad_w = _f32c(grad_w, "grad")
    points = _f32c(points, "points")
    params = [_f32c(p, "weight-network parameter") for p in params]
    M, E = adj32.shape
    shapes = [tuple(p.shape) for p in params]
    flat = torch.zeros(_flat_layout(shapes)[2], dtype=torch.float32, device=grad_w.device)
    grads = _carve(flat, shapes)
    wsb = C.lib().qc_weight_map_bwd_workspace_bytes(M)     # per-block partial sums, added in block order (bit-reproducible)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=grad_w.device)
    C.check(C.lib().qc_weight_map_bwd_ws(C.ptr(points), C.ptr(adj32), M, E, points.shape[1], _ptr_array(params),
                                         C.ptr(grad_w), _ptr_array(grads), C.ptr(ws), wsb, C.stream()), "qc_weight_map_bwd_ws")
    return flat


@weight_map_bwd.register_fake
def _(grad_w, points, adj32, params):
    return grad_w.new_empty(_flat_layout([tuple(p.shape) for p in params])[2])


class _WeightMapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, on_backward, points, adj32, inc_ptr, inc_idx, *params):
        ctx.save_for_backward(points, adj32, *params)
        ctx.on_backward = on_backward
        return _call(weight_map, points, adj32, inc_ptr, inc_idx, list(params))

    @staticmethod
    def backward(ctx, grad_w):
        points, adj32, *params = ctx.saved_tensors
        if ctx.on_backward is not None:
            ctx.on_backward()          # the graph behind the memoised weights is being consumed (MeshHandler.weight_map)
        flat = _call(weight_map_bwd, grad_w, points, adj32, list(params))
        return (None, None, None, None, None, *_carve(flat, [tuple(p.shape) for p in params]))


def weight_map_apply(points, adj32, inc_ptr, inc_idx, params, on_backward=None):
    return _WeightMapFn.apply(on_backward, points, adj32, inc_ptr, inc_idx, *params)


# ---------------------------------------------------------------------------------------------
# Mesh_MaxPool
# ---------------------------------------------------------------------------------------------

@torch.library.custom_op("quadconv_b200::segmax", mutates_args=())
@_on_device
def segmax(x: Tensor, group: Tensor, grp_ptr: Tensor, grp_mem: Tensor, n_out: int, act: int = 0) -> Tensor:
    """Segmented max over the pool groups (mesh_pool.py:88-89); act = 1: of sin(x) (fused activation, csrc/pool.cu)."""
    x = _f32c(x, "pool input")
    B, Cc, Nin = x.shape
    out = torch.empty((B, Cc, n_out), dtype=torch.float32, device=x.device)
    C.check(C.lib().qc_segmax_act_fwd(C.ptr(x), C.ptr(grp_ptr), C.ptr(grp_mem), B * Cc, Nin, n_out, act, C.ptr(out),
                                      C.stream()), "qc_segmax_act_fwd")
    return out


@segmax.register_fake
def _(x, group, grp_ptr, grp_mem, n_out, act=0):
    return x.new_empty((x.shape[0], x.shape[1], n_out))


@torch.library.custom_op("quadconv_b200::segmax_bwd", mutates_args=())
@_on_device
def segmax_bwd(x: Tensor, out: Tensor, grad_out: Tensor, group: Tensor, grp_ptr: Tensor, grp_mem: Tensor,
               act: int = 0) -> Tensor:
    x, out, grad_out = _f32c(x, "x"), _f32c(out, "out"), _f32c(grad_out, "grad")
    B, Cc, Nin = x.shape
    gx = torch.empty_like(x)
    C.check(C.lib().qc_segmax_act_bwd(C.ptr(x), C.ptr(out), C.ptr(grad_out), C.ptr(group), C.ptr(grp_ptr),
                                      C.ptr(grp_mem), B * Cc, Nin, out.shape[2], act, C.ptr(gx), C.stream()),
            "qc_segmax_act_bwd")
    return gx


@segmax_bwd.register_fake
def _(x, out, grad_out, group, grp_ptr, grp_mem, act=0):
    return torch.empty_like(x)


@torch.library.custom_op("quadconv_b200::gather_pool", mutates_args=())
@_on_device
def gather_pool(x: Tensor, index: Tensor, act: int = 0) -> Tensor:
    x = _f32c(x, "pool input")
    B, Cc, Nin = x.shape
    n_out = index.numel()
    out = torch.empty((B, Cc, n_out), dtype=torch.float32, device=x.device)
    C.check(C.lib().qc_gather_act_fwd(C.ptr(x), C.ptr(index), B * Cc, Nin, n_out, act, C.ptr(out), C.stream()),
            "qc_gather_act_fwd")
    return out


@gather_pool.register_fake
def _(x, index, act=0):
    return x.new_empty((x.shape[0], x.shape[1], index.numel()))


@torch.library.custom_op("quadconv_b200::gather_pool_bwd", mutates_args=())
@_on_device
def gather_pool_bwd(grad_out: Tensor, grp_ptr: Tensor, grp_mem: Tensor, n_coarse: int, x: Optional[Tensor] = None,
                    act: int = 0) -> Tensor:
    grad_out = _f32c(grad_out, "grad")
    B, Cc, Nf = grad_out.shape
    gx = torch.empty((B, Cc, n_coarse), dtype=torch.float32, device=grad_out.device)
    xp = C.ptr(_f32c(x, "pool input")) if (act and x is not None) else None
    C.check(C.lib().qc_gather_act_bwd(C.ptr(grad_out), xp, C.ptr(grp_ptr), C.ptr(grp_mem), B * Cc, Nf, n_coarse, act,
                                      C.ptr(gx), C.stream()), "qc_gather_act_bwd")
    return gx


@gather_pool_bwd.register_fake
def _(grad_out, grp_ptr, grp_mem, n_coarse, x=None, act=0):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n_coarse))


def _segmax_setup(ctx, inputs, output):
    x, group, grp_ptr, grp_mem, n_out, act = inputs
    ctx.save_for_backward(x, output, group, grp_ptr, grp_mem)
    ctx.act = act


def _segmax_backward(ctx, grad):
    x, out, group, grp_ptr, grp_mem = ctx.saved_tensors
    return segmax_bwd(x, out, grad, group, grp_ptr, grp_mem, ctx.act), None, None, None, None, None


segmax.register_autograd(_segmax_backward, setup_context=_segmax_setup)


class _SegmaxFn(torch.autograd.Function):
    """Same autograd as the registered one above, with the op bodies called directly in eager mode (_call)."""

    @staticmethod
    def forward(ctx, x, group, grp_ptr, grp_mem, n_out, act):
        out = _call(segmax, x, group, grp_ptr, grp_mem, n_out, act)
        ctx.save_for_backward(x, out, group, grp_ptr, grp_mem)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, grad):
        x, out, group, grp_ptr, grp_mem = ctx.saved_tensors
        return _call(segmax_bwd, x, out, grad, group, grp_ptr, grp_mem, ctx.act), None, None, None, None, None


def segmax_apply(x, group, grp_ptr, grp_mem, n_out, act=0):
    return _SegmaxFn.apply(x, group, grp_ptr, grp_mem, n_out, act)


class _GatherPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, index, grp_ptr, grp_mem, act):
        if act:
            ctx.save_for_backward(grp_ptr, grp_mem, x)
        else:
            ctx.save_for_backward(grp_ptr, grp_mem)
        ctx.n_coarse, ctx.act = x.shape[2], act
        return _call(gather_pool, x, index, act)

    @staticmethod
    def backward(ctx, grad):
        grp_ptr, grp_mem, *rest = ctx.saved_tensors
        return _call(gather_pool_bwd, grad, grp_ptr, grp_mem, ctx.n_coarse, rest[0] if rest else None, ctx.act), None, None, None, None


def gather_pool_apply(x, index, grp_ptr, grp_mem, act=0):
    return _GatherPoolFn.apply(x, index, grp_ptr, grp_mem, act)
