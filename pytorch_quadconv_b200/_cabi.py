"""
ctypes binding of libquadconv_b200.so (include/quadconv_b200.h).

This is the whole FFI surface: raw device pointers, sizes and a stream handle.  There is NO CPU
fallback -- if the shared library is missing or a kernel returns an error the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquadconv_b200.so")
QC_MAX_MLP_LAYERS = 8


class QcMlp(ctypes.Structure):
    """qc_mlp_t"""
    _fields_ = [("n_hidden", c_int),
                ("dims", c_int * (QC_MAX_MLP_LAYERS + 1)),
                ("w", c_void_p * QC_MAX_MLP_LAYERS),
                ("operand_bf16", c_int)]


class QcLayer(ctypes.Structure):
    """qc_layer_t"""
    _fields_ = [("d", c_int), ("Ni", c_int), ("No", c_int), ("Ci", c_int), ("Co", c_int), ("H", c_int),
                ("P", c_int64),
                ("out_pts", c_void_p), ("in_pts", c_void_p), ("quad_w", c_void_p),
                ("mlp", QcMlp),
                ("w_last", c_void_p), ("bias", c_void_p),
                ("pair_row", c_void_p), ("pair_col", c_void_p), ("row_ptr", c_void_p),
                ("csc_ptr", c_void_p), ("csc_pair", c_void_p), ("csc_row", c_void_p),
                ("order_out", c_void_p), ("order_in", c_void_p)]


_SIGNATURES = {
    "qc_version": (c_int, []),
    "qc_search_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "qc_search_count": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_size_t,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "qc_search_fill": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_size_t,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qc_exclusive_scan_i32": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "qc_pairs_to_csr": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "qc_csc_count": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "qc_csc_fill": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    "qc_balance_order": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "qc_pack_features": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_unpack_features": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "qc_batch_sum": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_pair_phi_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                ctypes.POINTER(QcMlp), c_int, c_void_p, c_void_p]),
    "qc_pair_phi_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                ctypes.POINTER(QcMlp), c_int, c_void_p, c_int, ctypes.POINTER(c_void_p),
                                c_void_p, c_void_p]),
    "qc_contract_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "qc_contract": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_size_t, c_void_p]),
    "qc_dphi_num_slices": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "qc_dphi": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                        c_int, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "qc_gather_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "qc_gather_dot_num_slices": (c_int, [c_int, c_int, c_int]),
    "qc_gather_accumulate": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "qc_gather_dot": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                              c_int, c_int64, c_void_p, c_void_p]),
    "qc_gemm_nt_tf32x3": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, ctypes.c_longlong, c_void_p]),
    "qc_gemm_tn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "qc_gemm_tn_tf32x3": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "qc_repack_wlast": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_unpack_dwlast": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_segmax_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_segmax_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_int, c_void_p, c_void_p]),
    "qc_gather_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_gather_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_segmax_act_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_segmax_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "qc_gather_act_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_gather_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_mma_get_variant": (c_int, []),
    "qc_mma_set_variant": (c_int, [c_int]),
    "qc_mma_chunk_channels": (c_int, [c_int, c_int]),
    "qc_pack_features_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "qc_mma_smem_bytes": (c_int, [c_int, c_int, c_int, c_int]),
    "qc_contract_mma": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                c_void_p]),
    "qc_dw_mma_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "qc_dw_mma_smem_bytes": (c_int, [c_int, c_int, c_int, c_int]),
    "qc_dw_mma": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                          c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "qc_dphi_mma_smem_bytes": (c_int, [c_int, c_int, c_int, c_int]),
    "qc_dphi_mma": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "qc_layer_workspace_bytes": (c_size_t, [ctypes.POINTER(QcLayer), c_int]),
    "qc_layer_forward": (c_int, [ctypes.POINTER(QcLayer), c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "qc_layer_backward": (c_int, [ctypes.POINTER(QcLayer), c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_size_t, c_void_p]),
    "qc_tc_selftest": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "qc_tc_selftest_ss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "qc_tc_probe_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "qc_tc_mma_bench": (c_int, [c_int] * 13 + [c_void_p, c_void_p]),
    "qc_weight_map_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_void_p, c_void_p]),
    "qc_vertex_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "qc_weight_map_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_void_p,
                                  ctypes.POINTER(c_void_p), c_void_p]),
    "qc_weight_map_bwd_workspace_bytes": (c_size_t, [c_int]),
    "qc_weight_map_bwd_ws": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_void_p,
                                     ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "qc_pair_phi_bwd_workspace_bytes": (c_size_t, [c_int64, c_int, ctypes.POINTER(QcMlp), c_int]),
    "qc_pair_phi_bwd_ws": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, ctypes.POINTER(QcMlp), c_int,
                                   c_void_p, c_int, ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p, c_size_t, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pytorch_quadconv_b200.build` "
                "(nvcc, sm_100a). pytorch_quadconv_b200 has no CPU or eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


_ERRORS = {-1: "QC_EINVAL (bad argument)", -2: "QC_EUNSUP (shape not implemented by the kernels)",
           -3: "QC_EWORKSPACE (workspace too small)"}


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = _ERRORS.get(rc, f"cudaError {rc}")
        raise RuntimeError(f"libquadconv_b200: {what} failed: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL).  The tensor must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pytorch_quadconv_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor handed to the C ABI")
    return t.data_ptr()


def stream() -> int:
    """Raw handle of the current stream of the current device (what torch.cuda.current_stream().cuda_stream returns,
    without building a Stream object: this runs once per op call on a host-bound path)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def make_mlp(d: int, hidden_ws, operand_bf16: bool = False) -> QcMlp:
    m = QcMlp()
    m.operand_bf16 = 1 if operand_bf16 else 0
    n = len(hidden_ws)
    if n > QC_MAX_MLP_LAYERS:
        raise RuntimeError(f"filter MLP with {n} hidden layers exceeds QC_MAX_MLP_LAYERS={QC_MAX_MLP_LAYERS}")
    m.n_hidden = n
    m.dims[0] = d
    for l, W in enumerate(hidden_ws):
        if W.shape[1] != m.dims[l]:
            raise RuntimeError("filter MLP weight shapes do not chain")
        m.dims[l + 1] = W.shape[0]
        m.w[l] = ptr(W)
    return m
