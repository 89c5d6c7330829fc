"""Time qc_gemm_nt_tf32x3 against torch.matmul (fp32, TF32 off) on the C4 shapes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_quadconv_b200 import _cabi as C
L = C.lib()
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K, rpb) in [(80000, 128, 2048, 10000), (80000, 2048, 128, 0), (80000, 64, 4096, 10000)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    D = torch.empty(M * N, device="cuda")
    ours = t(lambda: C.check(L.qc_gemm_nt_tf32x3(C.ptr(A), C.ptr(B), C.ptr(D), M, N, K, rpb, C.stream()), "g"))
    ref = t(lambda: torch.matmul(A, B.t()))
    print((M, N, K, rpb), "ours %.3f ms  torch %.3f ms" % (ours, ref))
