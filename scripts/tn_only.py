"""Run the TN GEMM (dW_last of the wide path) at the C4 shape a few times (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_quadconv_b200 import _cabi as C  # noqa: E402

L = C.lib()
nb, Kb, M, N = 8, 10000, 4096, 64
T = torch.randn(nb * Kb, M, device="cuda")
X = torch.randn(nb, N, Kb, device="cuda")
D = torch.empty(N, M, device="cuda")
wsb = L.qc_gemm_tn_workspace_bytes(nb, Kb, M, N)
ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device="cuda")
for _ in range(3):
    C.check(L.qc_gemm_tn_tf32x3(C.ptr(T), C.ptr(X), C.ptr(D), nb, Kb, M, N, C.ptr(ws), wsb, C.stream()), "tn")
torch.cuda.synchronize()
