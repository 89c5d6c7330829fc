"""GPU: BASELINE.json's configurations at THEIR OWN sizes (C2 was covered by a 400-point
slice, C4 by a 500-point wide layer, C5 not at all).

* C1 / C4 (single layers, C4 = the wide path: 64 -> 128 channels, H = 32, ~200 neighbours per point): output and every
  gradient against the defining sums evaluated in float64 on the GPU in pair blocks (the oracle's formulas,
  oracle/quadconv_oracle.py:126-148, without the [P, Ci, Co] tensor ever being held whole); pairs bit-exact vs the C oracle.
* C2 (10k-point autoencoder, 4 pooling stages, 8 QuadConv + 8 pool layers): output, loss, d x and every parameter
  gradient (filters, biases, the mesh's weight network) against the float64 oracle composition on the CPU.
* C5 (the same model, batch 256): the gradients of 4 batch shards averaged (what the NCCL all-reduce computes) equal
  the full-batch gradients; the 2-rank NCCL version of this statement is tests/test_gpu_dist.py.
Tolerance: north_star's 1e-5 normwise."""
import numpy as np
import pytest
import torch

from conftest import relerr
from test_gpu_layer import _StubMesh

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _layer_vs_float64(N, d, Ci, Co, B, fs, r, block):
    import pytorch_quadconv_b200 as q
    from oracle import quadconv_oracle as O
    torch.manual_seed(0)
    H = fs[-1]
    pts = torch.rand(N, d)
    layer = q.QuadConv(spatial_dim=d, in_points=N, out_points=N, in_channels=Ci, out_channels=Co,
                       filter_seq=fs, output_same=True, decay_param=r, bias=True).cuda()
    with torch.no_grad():
        layer.bias.normal_()
    w = (torch.rand(N).cuda() + 0.5).requires_grad_()
    pc = pts.cuda()
    mesh = _StubMesh(pc, pc, w)
    x = torch.randn(B, Ci, N, device="cuda", requires_grad=True)
    y = layer(mesh, x)
    go = torch.randn_like(y)
    (y * go).sum().backward()
    ev = layer.eval_indices
    assert np.array_equal(ev.cpu().numpy(), O.eval_indices_c(pts.numpy(), pts.numpy(), layer.decay_param))

    lin = [m for m in layer.filter if isinstance(m, torch.nn.Linear)]
    hid = [m.weight.detach().double().requires_grad_() for m in lin[:-1]]
    w64 = w.detach().double().requires_grad_()
    wl = lin[-1].weight.detach().double().view(Ci, Co, H)
    xd, god = x.detach().double(), go.double()
    y64 = torch.zeros(B, Co, N, dtype=torch.float64, device="cuda")
    dx64 = torch.zeros(B, Ci, N, dtype=torch.float64, device="cuda")
    ref_last = torch.zeros(Ci, Co, H, dtype=torch.float64, device="cuda")
    for s0 in range(0, ev.shape[0], block):
        io, ii = ev[s0:s0 + block, 0], ev[s0:s0 + block, 1]
        h = (pc[io] - pc[ii]).double()                                       # quadconv.py:250-253 (fp32 subtract)
        for W in hid:
            h = torch.sin(h @ W.t())
        phi = w64[ii].unsqueeze(1) * h                                       # [c, H]: quadrature weight x hidden activations
        with torch.no_grad():
            G = torch.einsum("ch,ijh->cij", phi, wl)                         # the materialised filter block (quadconv.py:257)
            y64.index_add_(2, io, torch.einsum("cij,bic->bjc", G, xd[:, :, ii]))          # :216-225
            dx64.index_add_(2, ii, torch.einsum("cij,bjc->bic", G, god[:, :, io]))
            cp = torch.einsum("bic,bjc->cij", xd[:, :, ii], god[:, :, io])
            ref_last += torch.einsum("cij,ch->ijh", cp, phi)
            dphi = torch.einsum("cij,ijh->ch", cp, wl)
            del G, cp
        (phi * dphi).sum().backward()
    y64 += layer.bias.detach().double()
    assert _rel(y.detach(), y64) < TOL
    assert _rel(x.grad, dx64) < TOL
    assert _rel(lin[-1].weight.grad.view(Ci, Co, H), ref_last) < TOL
    for m, W in zip(lin[:-1], hid):
        assert _rel(m.weight.grad, W.grad) < TOL
    assert _rel(w.grad, w64.grad) < TOL
    assert _rel(layer.bias.grad, go.double().sum(dim=0, keepdim=True)) < TOL               # bias is [1, Co, No] (quadconv.py:68-72)
    return int(ev.shape[0])


def test_c1_layer_full_size_vs_float64(cuda_lib):
    """configs[0]: 2-D 2500 -> 2500 points, 16 -> 16 channels, batch 8, default radius (py_scripts/quadconv_layer_profile.py)."""
    _layer_vs_float64(2500, 2, 16, 16, 8, [4, 4], None, 1 << 16)


def test_c4_wide_layer_full_size_vs_float64(cuda_lib):
    """configs[3]: 64 -> 128 channels, ~200 neighbours per point, H = 32: the wide path (csrc/contract_wide.cu + gemm_tc.cu)."""
    N = 10_000
    P = _layer_vs_float64(N, 2, 64, 128, 8, [32, 32], float(np.sqrt(200 / (np.pi * N))), 1 << 14)
    assert 150 * N < P < 210 * N


def _autoencoder(B, seed=0, fused=True):
    import pytorch_quadconv_b200 as q
    from bench import QuadConvAutoencoder
    torch.manual_seed(seed)
    N = 10_000
    pts = torch.rand(N, 2)
    x = torch.randn(B, 4, N)
    mesh = q.MeshHandler(pts).construct([N, 0, 0, 0, 0], mirror=True, quad_map="mfnus")
    model = QuadConvAutoencoder(mesh, [4, 8, 16, 32, 32], [8, 8], fuse_activation=fused).cuda()
    return model, x.cuda()


@pytest.mark.parametrize("fused", [False, True])
def test_c2_autoencoder_full_size_vs_float64(cuda_lib, fused):
    """configs[1]: the 10k-point autoencoder (8 QuadConv + 4 pool + 4 adjoint-pool layers, sin activations, MSE) at full
    size, batch 2, against the same composition evaluated in float64 by the oracle on the CPU.  fused: the sin
    activations evaluated inside the pool kernels (Mesh_MaxPool(activation='sin')) instead of torch.sin launches."""
    import ref64
    from oracle import quadconv_oracle as O
    model, x = _autoencoder(2, fused=fused)
    mesh = model.mesh
    x.requires_grad_()
    y = model(x)
    loss = torch.nn.functional.mse_loss(y, x.detach())
    loss.backward()
    lv = mesh.point_seq
    nl = len(lv) - 1
    assert nl == 4 and lv[0] == 10_000

    sd = {k: v.detach().cpu().numpy() for k, v in mesh.state_dict().items() if k.startswith("_weight_network")}
    mp = ref64.mesh_params64(sd)
    pts = [mesh._points[l].detach().cpu() for l in range(nl + 1)]
    adj = [mesh._adjacency[l].detach().cpu() for l in range(nl + 1)]
    qw = {}

    def weights(l):
        if l not in qw:
            qw[l] = O.weight_map(pts[l], adj[l], mp)
        return qw[l]

    P64 = []

    def conv64(m, lvl, h):
        ev = m.eval_indices.cpu()
        assert np.array_equal(ev.numpy(), O.eval_indices_c(pts[lvl].numpy(), pts[lvl].numpy(), m.decay_param)), lvl
        ws = [mm.weight.detach().cpu().numpy() for mm in m.filter if isinstance(mm, torch.nn.Linear)]
        y_, ws64, b64 = ref64.layer64(pts[lvl], pts[lvl], weights(lvl), ws, h, ev, m.in_channels, m.out_channels,
                                      m.bias.detach().cpu().numpy())
        P64.append((m, ws64, b64))
        return y_

    x64 = x.detach().cpu().double().requires_grad_()
    h = x64
    for l, (conv, pool) in enumerate(zip(model.enc, model.pools)):
        h = O.maxpool_forward(torch.sin(conv64(conv, l, h)), pool.index.cpu().long(), lv[l + 1])
    for i, (up, conv) in enumerate(zip(model.unpools, model.dec)):
        l = nl - 1 - i
        h = conv64(conv, l, O.maxpool_adjoint(h, up.index.cpu().long()))
        if i + 1 < nl:
            h = torch.sin(h)
    loss64 = torch.nn.functional.mse_loss(h, x64.detach())
    loss64.backward()
    assert relerr(y.detach().cpu().numpy(), h.detach().numpy()) < TOL
    assert abs(float(loss) - float(loss64)) < TOL * abs(float(loss64))
    assert relerr(x.grad.cpu().numpy(), x64.grad.numpy()) < TOL
    for m, ws64, b64 in P64:
        lin = [mm for mm in m.filter if isinstance(mm, torch.nn.Linear)]
        for mm, W in zip(lin, ws64):
            assert relerr(mm.weight.grad.cpu().numpy(), W.grad.numpy()) < TOL
        assert relerr(m.bias.grad.cpu().numpy(), b64.grad.numpy()) < TOL
    named = dict(mesh.named_parameters())
    for k, p64 in zip(ref64.mesh_param_names(), mp):
        assert relerr(named[k].grad.cpu().numpy(), p64.grad.numpy()) < TOL, k


def test_c5_batch256_sharded_gradients_equal_full_batch(cuda_lib):
    """configs[4] at its own size: batch 256 (eight batch blocks of 32 lanes) over the 10k-point autoencoder; four shards
    of 64 run one after the other, their flat gradient buffers averaged as ncclAllReduce(AVG) does."""
    from pytorch_quadconv_b200.distributed import FlatGradAllReduce, shard_batch, trainable_parameters
    model, x = _autoencoder(256)
    params = trainable_parameters([model])

    def grads(xb):
        for p in params:
            p.grad = None
        loss = torch.nn.functional.mse_loss(model(xb), xb)
        loss.backward()
        return float(loss)

    loss_full = grads(x)
    full = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    red = FlatGradAllReduce(params)
    total = torch.zeros_like(red.flat)
    world, loss_sum = 4, 0.0
    for r in range(world):
        loss_sum += grads(shard_batch(x, r, world))
        red.pack()
        total += red.flat
    total /= world
    assert abs(loss_sum / world - loss_full) < TOL * abs(loss_full)
    assert _rel(total, full) < TOL
    off = 0
    for p in params:
        n = p.numel()
        assert _rel(total[off:off + n], full[off:off + n]) < TOL
        off += n
