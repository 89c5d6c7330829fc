"""Time the dense GEMMs of the wide path on the C4 shapes: qc_gemm_nt_tf32x3 with the TMA-fed warp-specialised kernel
(default) and with the cp.async kernel (QC_GEMM_NO_TMA=1), qc_gemm_tn_tf32x3 (dW_last), and torch.matmul (cuBLAS
fp32, TF32 off) as the library yardstick."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_quadconv_b200 import _cabi as C  # noqa: E402

L = C.lib()


def t(f, n=10):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (M, N, K, rpb) in [(80000, 128, 2048, 10000), (80000, 2048, 128, 0), (80000, 64, 4096, 10000)]:
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    D = torch.empty(M * N, device="cuda")
    call = lambda: C.check(L.qc_gemm_nt_tf32x3(C.ptr(A), C.ptr(B), C.ptr(D), M, N, K, rpb, C.stream()), "g")
    os.environ.pop("QC_GEMM_NO_TMA", None)
    tma = t(call)
    os.environ["QC_GEMM_NO_TMA"] = "1"
    cpa = t(call)
    os.environ.pop("QC_GEMM_NO_TMA", None)
    ref = t(lambda: torch.matmul(A, B.t()))
    gb = (M * K + N * K + M * N) * 4 / 1e9
    print((M, N, K, rpb), "tma %.3f ms (%.0f GB/s)  cp.async %.3f ms  torch %.3f ms" % (tma, gb / tma * 1e3, cpa, ref), flush=True)

# dW_last of C4: T' [B*Ni, Co*H] = [80000, 4096], x [B, Ci, Ni] = [8, 64, 10000]
nb, Kb, M, N = 8, 10000, 4096, 64
T = torch.randn(nb * Kb, M, device="cuda")
X = torch.randn(nb, N, Kb, device="cuda")
D = torch.empty(N, M, device="cuda")
wsb = L.qc_gemm_tn_workspace_bytes(nb, Kb, M, N)
ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device="cuda")
tn = t(lambda: C.check(L.qc_gemm_tn_tf32x3(C.ptr(T), C.ptr(X), C.ptr(D), nb, Kb, M, N, C.ptr(ws), wsb, C.stream()), "tn"))
Xt = X.permute(1, 0, 2).reshape(N, nb * Kb).contiguous()
ref = t(lambda: torch.matmul(Xt, T))
print("TN", (nb, Kb, M, N), "ours %.3f ms (%.0f GB/s)  torch %.3f ms" % (tn, T.numel() * 4 / 1e9 / tn * 1e3, ref), flush=True)
