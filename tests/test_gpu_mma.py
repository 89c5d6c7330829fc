"""GPU: the reduced-precision tensor-core mode, QuadConv(compute_dtype=torch.bfloat16) (csrc/contract_mma.cu:
both contraction stages on tcgen05 with bf16 operands, fp32 accumulation).

Stated bounds (north_star "with a stated bf16 bound"), normwise relative:
  * vs the oracle doing the SAME operand roundings in float64 (quadconv_forward_bf16tc): <= 2e-4
    (what is left is fp32 accumulation order and roundings that flip at bf16 boundaries);
  * vs the exact fp32-mode layer (float64 oracle of the reference formulas):              <= 1e-2 (measured ~3e-3),
    and > 1e-5, i.e. the mode really is a different arithmetic.
fp32 mode stays the parity mode (tests/test_gpu_layer.py, 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from test_gpu_layer import _StubMesh, _oracle_case

pytestmark = pytest.mark.gpu
TOL_SAME_ROUNDING = 2e-4
TOL_VS_FP32 = 1e-2


def _layer(c, d, n, ci, co, fs, r, bias, No, same):
    import pytorch_quadconv_b200 as q
    layer = q.QuadConv(spatial_dim=d, in_points=n, out_points=No, in_channels=ci, out_channels=co,
                       filter_seq=fs, bias=bias, output_same=same, decay_param=r,
                       compute_dtype=torch.bfloat16).cuda()
    with torch.no_grad():
        lin = [m for m in layer.filter if isinstance(m, torch.nn.Linear)]
        for m, w in zip(lin, c["ws"]):
            m.weight.copy_(w.detach())
        if bias:
            layer.bias.copy_(c["b"].detach())
    return layer, lin


def test_stage_b_exposed(cuda_lib):
    """Diagnostic known-answer test: with 4 input channels (one column chunk) and W_last chosen so that output
    channel j = h*4 + i picks T[o,b,i,h], the layer output IS the gather stage's result T (rounded to bf16)."""
    from oracle import quadconv_oracle as O
    d, n, ci, co, B, fs, r = 2, 600, 4, 32, 32, [8, 8], 0.1
    c = _oracle_case(d, n, ci, co, B, fs, r, False, seed=77)
    wl = torch.zeros(ci * co, 8)
    for i in range(ci):
        for h in range(8):
            wl[i * co + (h * 4 + i), h] = 1.0
    c["ws"][-1] = wl
    layer, _ = _layer(c, d, n, ci, co, fs, r, False, n, True)
    assert layer.compute_dtype == torch.bfloat16
    mesh = _StubMesh(c["pts"].cuda(), c["out_pts"].cuda(), c["qw"].detach().cuda())
    y = layer(mesh, c["x"].detach().cuda())
    assert layer.tc_active, "the tensor-core path did not run"
    ref = O.quadconv_forward_bf16tc(c["out_pts"], c["pts"], c["qw"].detach(), [w.detach() for w in c["ws"]],
                                    c["x"].detach(), c["pairs"], ci, co)
    err = relerr(y.detach().cpu().numpy(), ref.numpy())
    assert err < TOL_SAME_ROUNDING, err


@pytest.mark.parametrize("d,n,ci,co,B,fs,r,bias,n_out", [
    (3, 1500, 32, 32, 32, [8, 8], 0.16, True, None),    # C3-like: 8 column chunks of 4 channels, 4 row blocks
    (2, 900, 16, 16, 8, [4, 4], 0.1, False, None),      # C1-like: batch 8 -> 16 channels per chunk, H = 4 (phi rows of 4)
    (2, 1000, 4, 8, 8, [8, 8], 0.08, True, None),       # C2's first layer: 4 channels, 32-column chunks
    (2, 800, 8, 12, 16, [8, 8], 0.09, True, 300),       # batch 16 -> 8 channels per chunk; cross-level (No != Ni)
    (2, 700, 20, 24, 40, [8], 0.1, False, None),        # batch 40: two batch blocks; partial last chunk; one hidden layer
    (2, 1200, 8, 8, 8, [8, 8], 0.012, True, 500),       # tiny radius: most out points have no pair, empty tiles
    (2, 50, 8, 8, 32, [8, 8], 0.3, False, None),        # fewer points than four tiles
    (3, 900, 32, 64, 32, [8, 5], 0.12, False, None),    # 64 output channels (TMEM full), H = 5
    (2, 600, 6, 5, 9, [6], 0.15, True, None),           # odd channels, batch 9 -> 16 lanes
])
def test_tc_forward_matches_same_rounding_oracle(cuda_lib, d, n, ci, co, B, fs, r, bias, n_out):
    from oracle import quadconv_oracle as O
    c = _oracle_case(d, n, ci, co, B, fs, r, bias, seed=7000 + n + ci, n_out=n_out)
    same = n_out is None
    No = c["out_pts"].shape[0]
    layer, lin = _layer(c, d, n, ci, co, fs, r, bias, No, same)
    w_dev = c["qw"].detach().cuda().requires_grad_()
    mesh = _StubMesh(c["pts"].cuda(), c["out_pts"].cuda(), w_dev)
    x = c["x"].detach().cuda().requires_grad_()
    y = layer(mesh, x)
    assert layer.tc_active and layer._tc["fwd"][4] % 32 == 0
    assert torch.equal(layer.eval_indices.cpu(), c["pairs"])                # the pair list does not depend on the mode
    ref = O.quadconv_forward_bf16tc(c["out_pts"], c["pts"], c["qw"].detach(), [w.detach() for w in c["ws"]],
                                    c["x"].detach(), c["pairs"], ci, co, c["b"].detach() if bias else None)
    e_same = relerr(y.detach().cpu().numpy(), ref.numpy())
    e_fp32 = relerr(y.detach().cpu().numpy(), c["y"].numpy())
    assert e_same < TOL_SAME_ROUNDING, (e_same, e_fp32)
    assert 1e-5 < e_fp32 < TOL_VS_FP32, e_fp32
    # backward of the mode: d(phi), d W_last and d features on the tensor cores where the shape is served (the C2-first-
    # layer shape with 32-column chunks and 64 output channels fall back to the fp32 kernels); exact-gradient bound
    assert layer.tc_backward_active == (not (ci == 4 and B == 8) and co <= 32 or (d, n) == (2, 700)), (layer.tc_backward_active,)
    (y * c["go"].cuda()).sum().backward()
    assert relerr(x.grad.cpu().numpy(), c["x"].grad.numpy()) < TOL_VS_FP32
    assert relerr(w_dev.grad.cpu().numpy(), c["qw"].grad.numpy()) < TOL_VS_FP32
    for m, w in zip(lin, c["ws"]):
        assert relerr(m.weight.grad.cpu().numpy(), w.grad.numpy()) < TOL_VS_FP32


@pytest.mark.parametrize("d,n,ci,co,B,fs,r", [(3, 1500, 32, 32, 32, [8, 8], 0.16), (2, 800, 8, 12, 16, [8, 8], 0.09),
                                             (2, 700, 20, 24, 40, [8], 0.1)])
def test_tma_gather4_producers_match_cp_async(cuda_lib, d, n, ci, co, B, fs, r):
    """The union rows fetched by TMA (cp.async.bulk.tensor.2d tile::gather4 into the same 128-byte-swizzled stages, out
    of bounds rows for the padding slots) feed the same MMAs: output and every gradient are bit-identical to the
    cp.async producers'."""
    res = []
    c = _oracle_case(d, n, ci, co, B, fs, r, True, seed=4100 + n)
    # ONE layer for both runs: the union lists are built once (slot order inside a tile is not fixed across builds)
    layer, lin = _layer(c, d, n, ci, co, fs, r, True, n, True)
    w_dev = c["qw"].detach().cuda().requires_grad_()
    mesh = _StubMesh(c["pts"].cuda(), c["out_pts"].cuda(), w_dev)
    x = c["x"].detach().cuda().requires_grad_()
    old = cuda_lib.qc_mma_set_variant(-1)
    try:
        for var in (4, 12):
            cuda_lib.qc_mma_set_variant(var)
            assert cuda_lib.qc_mma_get_variant() == var
            for t in [x, w_dev] + [m.weight for m in lin]:
                t.grad = None
            y = layer(mesh, x)
            assert layer.tc_active and layer.tc_backward_active
            (y * c["go"].cuda()).sum().backward()
            torch.cuda.synchronize()
            res.append(([y.detach().clone(), x.grad.clone(), lin[-1].weight.grad.clone()],
                        [w_dev.grad.clone()] + [m.weight.grad.clone() for m in lin[:-1]]))
    finally:
        cuda_lib.qc_mma_set_variant(old)
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][1], res[1][1]):      # accumulated with float atomics downstream of the (identical) d(phi)
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) < 1e-5


def test_tc_falls_back_outside_its_shapes(cuda_lib):
    """H = 16 (wide filter) and batch 3 are not served by the tensor-core kernels: the layer runs the fp32 path."""
    import pytorch_quadconv_b200 as q
    for (ci, co, B, fs) in ((8, 8, 8, [16]), (8, 8, 3, [8, 8])):
        c = _oracle_case(2, 400, ci, co, B, fs, 0.12, False, seed=5)
        layer, _ = _layer(c, 2, 400, ci, co, fs, 0.12, False, 400, True)
        mesh = _StubMesh(c["pts"].cuda(), c["out_pts"].cuda(), c["qw"].detach().cuda())
        y = layer(mesh, c["x"].detach().cuda())
        assert not layer.tc_active
        assert relerr(y.detach().cpu().numpy(), c["y"].numpy()) < 1e-5


def _rb(t):
    return t.float().to(torch.bfloat16).double()


@pytest.mark.parametrize("d,n,ci,co,B,H,r,n_out", [
    (3, 1500, 32, 32, 32, 8, 0.16, None),     # C3-like
    (2, 900, 16, 16, 8, 4, 0.1, None),        # C1-like: batch 8, 16 channels per chunk, phi rows of 4
    (2, 800, 8, 12, 16, 8, 0.09, 300),        # batch 16, cross-level
    (2, 700, 20, 24, 40, 8, 0.1, None),       # two batch blocks (atomic accumulation), partial last chunk
    (2, 500, 8, 8, 8, 5, 0.1, None),          # batch 8 with 8 channels: 64-column chunks, H = 5
])
def test_dphi_mma_kernel(cuda_lib, d, n, ci, co, B, H, r, n_out):
    """qc_dphi_mma through the C ABI against the same sums in float64 with the same operand roundings (features,
    grad_out, W_last and dT rounded to bf16): <= 2e-3 (rounding flips of dT), and <= 1e-2 from the exact value."""
    from pytorch_quadconv_b200 import _cabi as C, ops
    g = torch.Generator().manual_seed(31 + n)
    pts = torch.rand(n, d, generator=g)
    out_pts = pts if n_out is None else pts[torch.randperm(n, generator=g)[:n_out]].contiguous()
    No = out_pts.shape[0]
    x = torch.randn(B, ci, n, generator=g)
    go = torch.randn(B, co, No, generator=g)
    wl = torch.randn(ci * co, H, generator=g) / np.sqrt(H)
    res = ops.build_pairs(out_pts.cuda(), pts.cuda(), float(r), n_out is None)
    ev, pair_row, pair_col, row_ptr = res[0], res[1], res[2], res[3]
    cell_out = res[9]
    ent, pptr, uptr, uidx, max_ku = ops.build_union_tables(row_ptr, pair_col, cell_out, No, n)
    P = ev.shape[0]
    HP = 4 if H <= 4 else 8
    xh, gh = ops._pack_bf16(x.cuda()), ops._pack_bf16(go.cuda())
    dphi = torch.zeros(P, HP, device="cuda")
    L = C.lib()
    assert 0 < L.qc_dphi_mma_smem_bytes(ci, co, B, max_ku) <= 227 * 1024
    C.check(L.qc_dphi_mma(C.ptr(xh), n, ci, C.ptr(gh), co, C.ptr(ent), C.ptr(pptr), C.ptr(uptr), C.ptr(uidx),
                          uptr.numel() - 1, max_ku, C.ptr(wl.cuda()), H, HP, C.ptr(cell_out), No, B, C.ptr(dphi),
                          C.stream()), "qc_dphi_mma")
    torch.cuda.synchronize()
    io, ii = ev[:, 0].cpu(), ev[:, 1].cpu()
    W = wl.view(ci, co, H)

    def ref(rounded):
        f = (lambda t: _rb(t)) if rounded else (lambda t: t.double())
        dT = torch.einsum("bjo,ijh->obih", f(go), f(W))
        if rounded:
            dT = _rb(dT)
        out = torch.empty(P, H, dtype=torch.float64)
        xs = f(x)
        for s in range(0, P, 4096):
            out[s:s + 4096] = torch.einsum("bip,pbih->ph", xs[:, :, ii[s:s + 4096]], dT[io[s:s + 4096]])
        return out
    got = dphi[:, :H].cpu()
    assert relerr(got.numpy(), ref(True).numpy()) < 2e-3
    assert relerr(got.numpy(), ref(False).numpy()) < 1e-2
    if HP > H:
        assert float(dphi[:, H:].abs().max()) == 0.0


@pytest.mark.parametrize("d,n,ci,co,B,H,r,n_out", [
    (3, 1500, 32, 32, 32, 8, 0.16, None),     # C3-like: chunk pairs (2 x 32 k-values = one M = 64 accumulator), 4 groups
    (2, 900, 16, 16, 8, 4, 0.1, None),        # batch 8, 16 channels per chunk: one M = 128 accumulator
    (2, 800, 8, 12, 16, 8, 0.09, 300),        # batch 16, 8 channels per chunk (M = 64 single chunk), cross-level
    (2, 700, 20, 24, 40, 8, 0.1, None),       # two batch blocks, odd number of chunks (last group half empty)
    (2, 1000, 4, 8, 8, 8, 0.08, None),        # 4 channels, 32-column chunks
])
def test_dw_mma_kernel(cuda_lib, d, n, ci, co, B, H, r, n_out):
    """qc_dw_mma through the C ABI: d W_last = sum_{o,b} T[o;b,i,h] g[b,j,o] against float64 with the same operand
    roundings (features, phi, T and grad_out rounded to bf16): <= 2e-3, and <= 1e-2 from the exact value."""
    from pytorch_quadconv_b200 import _cabi as C, ops
    g = torch.Generator().manual_seed(57 + n)
    pts = torch.rand(n, d, generator=g)
    out_pts = pts if n_out is None else pts[torch.randperm(n, generator=g)[:n_out]].contiguous()
    No = out_pts.shape[0]
    x = torch.randn(B, ci, n, generator=g)
    go = torch.randn(B, co, No, generator=g)
    res = ops.build_pairs(out_pts.cuda(), pts.cuda(), float(r), n_out is None)
    ev, pair_row, pair_col, row_ptr, cell_out = res[0], res[1], res[2], res[3], res[9]
    P = ev.shape[0]
    HP = 4 if H <= 4 else 8
    phi = torch.zeros(P, HP)
    phi[:, :H] = torch.randn(P, H, generator=g)
    ent, pptr, uptr, uidx, max_ku = ops.build_union_tables(row_ptr, pair_col, cell_out, No, n)
    xh, gh = ops._pack_bf16(x.cuda()), ops._pack_bf16(go.cuda())
    L = C.lib()
    wsb = L.qc_dw_mma_workspace_bytes(ci, co, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    dw = torch.full((ci * co, H), float("nan"), device="cuda")
    C.check(L.qc_dw_mma(C.ptr(xh), n, ci, C.ptr(gh), co, C.ptr(ent), C.ptr(pptr), C.ptr(uptr), C.ptr(uidx),
                        uptr.numel() - 1, max_ku, C.ptr(phi.cuda()), H, HP, C.ptr(cell_out), No, B, C.ptr(ws), wsb,
                        C.ptr(dw), C.stream()), "qc_dw_mma")
    torch.cuda.synchronize()
    io, ii = ev[:, 0].cpu(), ev[:, 1].cpu()

    def ref(rounded):
        f = (lambda t: _rb(t)) if rounded else (lambda t: t.double())
        T = torch.zeros(No, B, ci, H, dtype=torch.float64)
        xs, ph = f(x), f(phi[:, :H])
        for s in range(0, P, 2048):
            T.index_add_(0, io[s:s + 2048], torch.einsum("ph,bip->pbih", ph[s:s + 2048], xs[:, :, ii[s:s + 2048]]))
        if rounded:
            T = _rb(T)
        return torch.einsum("obih,bjo->ijh", T, f(go)).reshape(ci * co, H)
    got = dw.cpu()
    assert torch.isfinite(got).all()
    assert relerr(got.numpy(), ref(True).numpy()) < 2e-3
    assert relerr(got.numpy(), ref(False).numpy()) < 1e-2
