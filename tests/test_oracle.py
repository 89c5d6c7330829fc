"""CPU: pin the oracle (oracle/) against the golden vectors produced by the live reference
(tests/golden/make_golden.py).  Pairs must be bit-exact; floats within 1e-6 normwise."""
import numpy as np
import pytest
import torch

from conftest import elim_map_from, golden_cases, load_golden, relerr
from oracle import quadconv_oracle as O


def test_search_c_matches_reference_pairs():
    g = load_golden("pairs")
    for name in golden_cases(g):
        p = O.eval_indices_c(g[name + "/out_pts"], g[name + "/in_pts"], float(g[name + "/r"]))
        assert p.dtype == np.int64
        assert np.array_equal(p, g[name + "/pairs"]), name


def test_search_torch_blocked_matches_reference_pairs():
    g = load_golden("pairs")
    for name in golden_cases(g):
        p = O.eval_indices_torch(torch.from_numpy(g[name + "/out_pts"]), torch.from_numpy(g[name + "/in_pts"]),
                                 float(g[name + "/r"]), block=97)
        assert torch.equal(p, torch.from_numpy(g[name + "/pairs"])), name


@pytest.mark.parametrize("name", ["layer_2d", "layer_2d_wide", "layer_3d", "layer_1to2"])
def test_forward_backward_matches_reference(name):
    L = load_golden(name)
    t = lambda k: torch.from_numpy(L[k])
    ws = [t(f"W{l}").requires_grad_() for l in range(10) if f"W{l}" in L.files]
    x, qw = t("x").requires_grad_(), t("quad_w").requires_grad_()
    bias = t("bias").requires_grad_() if "bias" in L.files else None
    y = O.quadconv_forward(t("pts"), t("pts"), qw, ws, x, t("pairs"), int(L["ci"]), int(L["co"]), bias,
                           block_pairs=1500)
    (y * t("grad_out")).sum().backward()
    tol = 1e-6
    assert relerr(y.detach().numpy(), L["y"]) < tol
    assert relerr(x.grad.numpy(), L["dx"]) < tol
    assert relerr(qw.grad.numpy(), L["dquad_w"]) < tol
    for l, w in enumerate(ws):
        assert relerr(w.grad.numpy(), L[f"dW{l}"]) < 2e-6
    if bias is not None:
        assert relerr(bias.grad.numpy(), L["dbias"]) < tol


def test_pool_index_and_values_match_reference():
    g = load_golden("pool")
    em0, em1 = elim_map_from(g, "elim0"), elim_map_from(g, "elim1")
    n0, n1, n2 = [int(v) for v in g["point_seq"]]
    assert np.array_equal(O.pool_group_index(em0, n0), g["index_pool0"])
    assert np.array_equal(O.pool_group_index(em1, n1), g["index_pool1"])
    assert np.array_equal(O.pool_adjoint_index(em1, n1), g["index_up1"])
    assert np.array_equal(O.pool_adjoint_index(em0, n0), g["index_up0"])
    x = torch.from_numpy(g["x"])
    y0 = O.maxpool_forward(x, torch.from_numpy(g["index_pool0"]), n1)
    assert torch.equal(y0, torch.from_numpy(g["y0"]))
    assert np.array_equal(O.segmax_c(g["x"], g["index_pool0"], n1), g["y0"])
    y1 = O.maxpool_forward(y0, torch.from_numpy(g["index_pool1"]), n2)
    assert torch.equal(y1, torch.from_numpy(g["y1"]))
    z1 = O.maxpool_adjoint(y1, torch.from_numpy(g["index_up1"]))
    assert torch.equal(z1, torch.from_numpy(g["z1"]))


def test_bf16_operand_mode_of_the_oracle():
    """The product's `filter_dtype=torch.bfloat16` mode is not in the reference; its oracle (filter_mlp with
    operand_bf16=True) must be exactly 'every Linear sees bf16-rounded weights and inputs, accumulation in the
    working precision', with unrounded (straight-through) gradients."""
    from oracle import quadconv_oracle as O
    g = torch.Generator().manual_seed(3)
    z = torch.randn(50, 2, generator=g, dtype=torch.float64)
    ws = [torch.randn(8, 2, generator=g, dtype=torch.float64), torch.randn(8, 8, generator=g, dtype=torch.float64),
          torch.randn(6, 8, generator=g, dtype=torch.float64)]
    rnd = lambda t: t.float().bfloat16().double()
    h = z
    for l, W in enumerate(ws):
        h = rnd(h) @ rnd(W).t()
        if l + 1 < len(ws):
            h = torch.sin(h)
    got = O.filter_mlp(ws, z, operand_bf16=True)
    assert torch.equal(got, h)
    assert not torch.equal(got, O.filter_mlp(ws, z))                 # the mode really rounds
    # straight-through: the gradient w.r.t. the last weight is the (rounded) last activation, not a bf16 value
    wl = ws[-1].clone().requires_grad_()
    O.filter_mlp(ws[:-1] + [wl], z, operand_bf16=True).sum().backward()
    a_last = z
    for W in ws[:-1]:
        a_last = torch.sin(rnd(a_last) @ rnd(W).t())
    expect = rnd(a_last).sum(0).expand(6, 8)
    assert torch.allclose(wl.grad, expect, rtol=1e-12, atol=1e-12)
    assert (wl.grad != rnd(wl.grad)).any()                           # not quantised to bf16


def test_float64_composition_agrees_with_live_reference_golden():
    """tests/ref64.py (the float64 checker the GPU tests hold the 1e-5 bound against) reproduces the live
    reference's fp32 autoencoder slice -- output, loss and every gradient -- to the reference's own fp32 noise."""
    import torch
    import ref64
    from oracle import quadconv_oracle as O
    A = load_golden("autoencoder")
    mp = ref64.mesh_params64({k[5:]: A[k] for k in A.files if k.startswith("mesh/")})
    pts = [torch.from_numpy(A[f"mesh/_points.{l}"]) for l in range(2)]
    adj = [torch.from_numpy(A[f"mesh/_adjacency.{l}"]) for l in range(2)]
    qw = [O.weight_map(pts[l], adj[l], mp) for l in range(2)]
    elim = elim_map_from(A, "elim0")
    n0, n1 = pts[0].shape[0], pts[1].shape[0]
    grp = torch.from_numpy(O.pool_group_index(elim, n0))
    upi = torch.from_numpy(O.pool_adjoint_index(elim, n0))
    x64 = torch.from_numpy(A["x"]).double().requires_grad_()
    P64 = {}

    def conv64(nm, lvl, h, ci, co):
        ws = [A[f"{nm}/filter.{2 * l}.weight"] for l in range(3)]
        y_, ws64, b64 = ref64.layer64(pts[lvl], pts[lvl], qw[lvl], ws, h, torch.from_numpy(A[f"{nm}/eval_indices"]),
                                      ci, co, A[f"{nm}/bias"])
        for l in range(3):
            P64[f"{nm}grad/filter.{2 * l}.weight"] = ws64[l]
        P64[f"{nm}grad/bias"] = b64
        return y_

    h = torch.sin(conv64("e0", 0, x64, 2, 4))
    h = O.maxpool_forward(h, grp, n1)
    h = torch.sin(conv64("e1", 1, h, 4, 4))
    h = O.maxpool_adjoint(h, upi)
    y = conv64("d0", 0, h, 4, 2)
    loss = torch.nn.functional.mse_loss(y, x64.detach())
    loss.backward()
    assert relerr(y.detach().numpy(), A["y"]) < 1e-5
    assert abs(float(loss) - float(A["loss"])) < 1e-5 * abs(float(A["loss"]))
    assert relerr(x64.grad.numpy(), A["dx"]) < 5e-5
    for k, p in P64.items():
        assert relerr(p.grad.numpy(), A[k]) < 5e-5, k
    for k, p in zip(ref64.mesh_param_names(), mp):
        assert relerr(p.grad.numpy(), A["meshgrad/" + k]) < 2e-4, k


def test_c_oracle_reproduces_c3_pair_hash():
    """The committed SHA-256 of the 100k-point C3 pair list (tests/golden/pairs_c3_sha256.json, written after the
    reference's blocked ATen ops and the C oracle agreed) is what the C oracle computes today (~16 s on 8 cores)."""
    import hashlib
    import json
    import os
    import torch
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "pairs_c3_sha256.json")))
    torch.manual_seed(ref["seed"])
    pts = torch.rand(ref["N"], ref["d"]).numpy()
    p = O.eval_indices_c(pts, pts, ref["radius"])
    assert p.shape == (ref["P"], 2)
    assert hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() == ref["sha256"]
