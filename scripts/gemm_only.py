"""Run the three NT GEMMs and the TN GEMM of the wide path once each at the C4 shapes (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_quadconv_b200 import _cabi as C  # noqa: E402

L = C.lib()
for rep in range(2):
    for (M, N, K, rpb) in [(80000, 128, 2048, 10000), (80000, 2048, 128, 0), (80000, 64, 4096, 10000)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        D = torch.empty(M * N, device="cuda")
        C.check(L.qc_gemm_nt_tf32x3(C.ptr(A), C.ptr(B), C.ptr(D), M, N, K, rpb, C.stream()), "nt")
    nb, Kb, M, N = 8, 10000, 4096, 64
    T = torch.randn(nb * Kb, M, device="cuda")
    X = torch.randn(nb, N, Kb, device="cuda")
    D = torch.empty(N, M, device="cuda")
    wsb = L.qc_gemm_tn_workspace_bytes(nb, Kb, M, N)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device="cuda")
    C.check(L.qc_gemm_tn_tf32x3(C.ptr(T), C.ptr(X), C.ptr(D), nb, Kb, M, N, C.ptr(ws), wsb, C.stream()), "tn")
    torch.cuda.synchronize()
