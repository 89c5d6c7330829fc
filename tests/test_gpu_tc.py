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
                                       (130, 64, 36, 65), (50, 7, 4, 0), (4000, 128, 1024, 0)])
@pytest.mark.parametrize("loader", ["tma", "cp_async"])
def test_tcgen05_gemm_nt_3xtf32(cuda_lib, M, N, K, rpb, loader, monkeypatch):
    """Dense stage of the wide path (csrc/gemm_tc.cu): D = A * B^T in 3xTF32, row-major or per-batch transposed
    output, ragged M / N / K tiles; the TMA-fed warp-specialised kernel (default) and the cp.async kernel (fallback)."""
    from pytorch_quadconv_b200 import _cabi as C
    if loader == "cp_async":
        monkeypatch.setenv("QC_GEMM_NO_TMA", "1")
    else:
        monkeypatch.delenv("QC_GEMM_NO_TMA", raising=False)
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


@pytest.mark.parametrize("nb,Kb,M,N", [(3, 100, 200, 40), (8, 1000, 512, 64), (2, 77, 130, 70), (1, 5000, 1024, 64),
                                        (5, 3922, 256, 7), (1, 1, 128, 1)])
@pytest.mark.parametrize("loader", ["tma", "registers"])
def test_tcgen05_gemm_tn_3xtf32(cuda_lib, nb, Kb, M, N, loader, monkeypatch):
    """dW_last of the wide path (csrc/gemm_tc.cu, k_gemm_tn_tf32x3): D[n][m] = sum_{b,k} X[b][n][k] T[b*Kb+k][m] in
    3xTF32 with T as the TMEM operand; ragged tiles, chunks that end inside a batch block, K splits, bounded
    accumulation chains (Kb = 5000: 79 chunks per CTA set), and run-to-run bit reproducibility."""
    from pytorch_quadconv_b200 import _cabi as C
    if loader == "registers":
        monkeypatch.setenv("QC_GEMM_NO_TMA", "1")       # the register-loading kernel (also the fallback for M % 4 != 0)
    else:
        monkeypatch.delenv("QC_GEMM_NO_TMA", raising=False)
    g = torch.Generator().manual_seed(nb + Kb + M + N)
    T = torch.randn(nb * Kb, M, generator=g).cuda()
    X = torch.randn(nb, N, Kb, generator=g).cuda()
    wsb = cuda_lib.qc_gemm_tn_workspace_bytes(nb, Kb, M, N)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(2):
        D = torch.full((N, M), float("nan"), device="cuda")
        C.check(cuda_lib.qc_gemm_tn_tf32x3(C.ptr(T), C.ptr(X), C.ptr(D), nb, Kb, M, N, C.ptr(ws), wsb, C.stream()),
                "qc_gemm_tn_tf32x3")
        torch.cuda.synchronize()
        outs.append(D)
    ref = X.double().permute(1, 0, 2).reshape(N, nb * Kb) @ T.double()
    err = float((outs[0].double() - ref).norm() / ref.norm())
    assert err < 2e-6, err
    assert torch.equal(outs[0], outs[1])


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


def test_single_call_layer_entry_points(cuda_lib):
    """qc_layer_forward / qc_layer_backward: the whole layer through two C calls on [B,C,N] tensors with a
    caller-owned workspace -- no ops.py composition -- against the float64 oracle (1e-5) for the output and every
    gradient.  The pair tables come from the search entry points, as a non-Python host would build them."""
    import ctypes
    import numpy as np
    from conftest import relerr
    from oracle import quadconv_oracle as O
    from pytorch_quadconv_b200 import _cabi as C, ops
    L = C.lib()
    g = torch.Generator().manual_seed(2024)
    d, n, ci, co, B, fs, r = 2, 700, 6, 10, 5, [8, 8], 0.1
    pts = torch.rand(n, d, generator=g)
    x = torch.randn(B, ci, n, generator=g)
    qw = torch.rand(n, generator=g) + 0.5
    dims = [d] + fs + [ci * co]
    ws_ = [torch.randn(dims[l + 1], dims[l], generator=g) / np.sqrt(dims[l]) for l in range(len(dims) - 1)]
    bias = torch.randn(co, n, generator=g)
    go = torch.randn(B, co, n, generator=g)
    dev = "cuda"
    res = [t for t in ops.build_pairs(pts.to(dev), pts.to(dev), r, True)]      # tables only (search kernels)
    ev, pair_row, pair_col, row_ptr, csc_ptr, csc_pair, csc_row, order_out, order_in = res[:9]
    dp = [t.to(dev).contiguous() for t in (pts, x, qw, bias, go)]
    wd = [w.to(dev).contiguous() for w in ws_]
    desc = C.QcLayer()
    desc.d, desc.Ni, desc.No, desc.Ci, desc.Co, desc.H, desc.P = d, n, n, ci, co, fs[-1], ev.shape[0]
    desc.out_pts = desc.in_pts = C.ptr(dp[0])
    desc.quad_w = C.ptr(dp[2])
    desc.mlp = C.make_mlp(d, wd[:-1])
    desc.w_last, desc.bias = C.ptr(wd[-1]), C.ptr(dp[3])
    for k, t in (("pair_row", pair_row), ("pair_col", pair_col), ("row_ptr", row_ptr), ("csc_ptr", csc_ptr),
                 ("csc_pair", csc_pair), ("csc_row", csc_row), ("order_out", order_out), ("order_in", order_in)):
        setattr(desc, k, C.ptr(t.contiguous()))
    wsb = L.qc_layer_workspace_bytes(ctypes.byref(desc), B)
    assert wsb > 0
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty(B, co, n, device=dev)
    st = C.stream()
    assert L.qc_layer_forward(ctypes.byref(desc), C.ptr(dp[1]), B, C.ptr(out), C.ptr(ws), wsb - 1, st) == -3   # QC_EWORKSPACE
    C.check(L.qc_layer_forward(ctypes.byref(desc), C.ptr(dp[1]), B, C.ptr(out), C.ptr(ws), wsb, st), "qc_layer_forward")
    dx = torch.empty_like(dp[1])
    dqw = torch.empty(n, device=dev)
    dwl = torch.empty_like(wd[-1])
    dhid = [torch.empty_like(w) for w in wd[:-1]]
    dbias = torch.empty(co, n, device=dev)
    dptrs = (ctypes.c_void_p * len(dhid))(*[C.ptr(t) for t in dhid])
    C.check(L.qc_layer_backward(ctypes.byref(desc), C.ptr(dp[4]), B, C.ptr(dx), C.ptr(dqw), C.ptr(dwl), dptrs, C.ptr(dbias),
                                C.ptr(ws), wsb, st), "qc_layer_backward")
    torch.cuda.synchronize()
    # float64 oracle of the reference formulas
    x64, qw64 = x.double().requires_grad_(), qw.double().requires_grad_()
    w64 = [w.double().requires_grad_() for w in ws_]
    b64 = bias.double().unsqueeze(0).requires_grad_()
    y64 = O.quadconv_forward(pts, pts, qw64, w64, x64, ev.cpu(), ci, co, b64)
    (y64 * go.double()).sum().backward()
    assert relerr(out.cpu().numpy(), y64.detach().numpy()) < 1e-5
    assert relerr(dx.cpu().numpy(), x64.grad.numpy()) < 1e-5
    assert relerr(dqw.cpu().numpy(), qw64.grad.numpy()) < 1e-5
    assert relerr(dwl.cpu().numpy(), w64[-1].grad.numpy()) < 1e-5
    for a_, b_ in zip(dhid, w64[:-1]):
        assert relerr(a_.cpu().numpy(), b_.grad.numpy()) < 1e-5
    assert relerr(dbias.cpu().numpy(), b64.grad[0].numpy()) < 1e-5
