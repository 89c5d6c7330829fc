"""GPU: known-answer test of the tcgen05 (tensor-core) building blocks: TMEM-resident A operand,
K-major shared-memory B descriptor, 3xTF32 error-compensated product, TMEM read-back."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_3xtf32_gemm(cuda_lib):
    from pytorch_quadconv_b200 import _cabi as C
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 64, generator=g).cuda()
    B = torch.randn(32, 64, generator=g).cuda()
    D = torch.full((128, 32), float("nan"), device="cuda")
    C.check(cuda_lib.qc_tc_selftest(C.ptr(A), C.ptr(B), C.ptr(D), C.stream()), "qc_tc_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    err = float((D.double() - ref).norm() / ref.norm())
    assert err < 2e-6, err          # plain TF32 would give ~5e-4
    # structured operands catch layout mix-ups that random data can hide
    A2 = torch.zeros(128, 64, device="cuda"); A2[5, 7] = 1.0; A2[100, 63] = 2.0
    B2 = torch.arange(32 * 64, dtype=torch.float32, device="cuda").reshape(32, 64)
    C.check(cuda_lib.qc_tc_selftest(C.ptr(A2), C.ptr(B2), C.ptr(D), C.stream()), "qc_tc_selftest")
    ref2 = A2.double() @ B2.double().T
    assert float((D.double() - ref2).abs().max()) < 1e-3
    # both operands through shared-memory descriptors (SS mode)
    D.fill_(float("nan"))
    C.check(cuda_lib.qc_tc_selftest_ss(C.ptr(A), C.ptr(B), C.ptr(D), C.stream()), "qc_tc_selftest_ss")
    torch.cuda.synchronize()
    assert float((D.double() - ref).norm() / ref.norm()) < 2e-6


@pytest.mark.parametrize("M,N,K,rpb", [(1000, 40, 68, 0), (1000, 40, 68, 250), (384, 128, 2048, 128), (700, 200, 128, 0),
                                       (130, 64, 36, 65), (50, 7, 4, 0)])
def test_tcgen05_gemm_nt_3xtf32(cuda_lib, M, N, K, rpb):
    """Dense stage of the wide path (csrc/gemm_tc.cu): D = A * B^T in 3xTF32, row-major or per-batch transposed
    output, ragged M / N / K tiles."""
    from pytorch_quadconv_b200 import _cabi as C
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = torch.full((M * N,), float("nan"), device="cuda")
    C.check(cuda_lib.qc_gemm_nt_tf32x3(C.ptr(A), C.ptr(B), C.ptr(D), M, N, K, rpb, C.stream()), "qc_gemm_nt_tf32x3")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T                                    # [M, N]
    if rpb:
        ref = ref.reshape(M // rpb, rpb, N).transpose(1, 2).contiguous()   # [M/R, N, R]
    err = float((D.double().reshape(ref.shape) - ref).norm() / ref.norm())
    assert err < 2e-6, err


def test_tcgen05_bf16_operand_majorness(cuda_lib):
    """Building block for a reduced-precision gather stage: with bf16 operands (kind::f16) the shared-memory
    operands may be MN-major (contiguous along M / N -- the order in which gathered feature rows arrive) as well
    as K-major; all four combinations give the exact bf16 product.  (The same descriptors with kind::tf32 return
    zeros for every MN-major variant, see profiles/README.md.)"""
    from pytorch_quadconv_b200 import _cabi as C
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 32, generator=g).cuda()
    B = torch.randn(32, 32, generator=g).cuda()
    ref = A.bfloat16().double() @ B.bfloat16().double().T
    for mode in range(4):
        D = torch.full((128, 32), float("nan"), device="cuda")
        C.check(cuda_lib.qc_tc_probe_bf16(C.ptr(A), C.ptr(B), C.ptr(D), mode, C.stream()), "qc_tc_probe_bf16")
        torch.cuda.synchronize()
        assert float((D.double() - ref).norm() / ref.norm()) < 1e-6, mode
