"""
tests/golden/make_golden.py -- generates tests/golden/*.npz from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Every array the parity tests need (inputs, parameters, reference outputs and autograd gradients) is
stored, so neither the tests nor the oracle ever import the reference at run time.  The reference is
imported unmodified from /root/reference (torch_quadconv/__init__.py:4-6).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from torch_quadconv import QuadConv, MeshHandler, Mesh_MaxPool  # noqa: E402
from torch_quadconv.utils.MFNUS2 import MFNUS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def ref_pairs(out_pts, in_pts, r, same):
    """Drives QuadConv._compute_eval_indices (quadconv.py:165-199) through a stub mesh so that any
    point sets (cross-level, 1-D, duplicates) can be used without MeshHandler.construct."""
    class _M:
        def input_points(self):
            return in_pts

        def output_points(self):
            return out_pts
    qc = QuadConv(spatial_dim=in_pts.shape[1], in_points=in_pts.shape[0], out_points=out_pts.shape[0],
                  in_channels=1, out_channels=1, filter_seq=[2], decay_param=r, output_same=same,
                  cache=False)
    return qc._compute_eval_indices(_M())


def golden_pairs():
    cases = {}
    g = torch.Generator().manual_seed(1234)
    # random clouds, d = 1, 2, 3, default radius 4/sqrt(N) (quadconv.py:59-62)
    for d, n in ((1, 300), (2, 700), (3, 900)):
        p = torch.rand(n, d, generator=g)
        r = 4 / np.sqrt(n) if d < 3 else 0.2
        cases[f"rand{d}d"] = (p, p, r, True)
    # lattice: many distances tie exactly with r
    xs = torch.linspace(0, 1, 21)
    lat = torch.stack(torch.meshgrid(xs, xs, indexing="ij"), -1).reshape(-1, 2).contiguous()
    cases["lattice_tie"] = (lat, lat, 0.1, True)
    cases["lattice_tie2"] = (lat, lat, float(np.float32(0.05) * 3), True)
    # duplicated points
    p = torch.rand(200, 2, generator=g)
    p = torch.cat([p, p[:50]], 0)
    cases["dups"] = (p, p, 0.15, True)
    # r -> 0 : self pairs (and duplicates) only
    cases["r0"] = (p, p, 0.0, True)
    # r >= diameter : all pairs
    q = torch.rand(60, 3, generator=g)
    cases["allpairs"] = (q, q, 2.0, True)
    # cross level No != Ni
    a = torch.rand(500, 2, generator=g)
    b = a[torch.randperm(500, generator=g)[:170]].contiguous()
    cases["cross_down"] = (b, a, 0.12, False)
    cases["cross_up"] = (a, b, 0.2, False)
    # offset / negative coordinates, anisotropic box
    c = torch.rand(400, 3, generator=g) * torch.tensor([4.0, 0.5, 2.0]) - torch.tensor([2.0, 0.1, -3.0])
    cases["box3d"] = (c, c, 0.35, True)
    arrs = {}
    for name, (o, i, r, same) in cases.items():
        idx = ref_pairs(o, i, r, same)
        arrs[name + "/out_pts"] = o
        arrs[name + "/in_pts"] = i
        arrs[name + "/r"] = np.float64(r)
        arrs[name + "/pairs"] = idx
    save("pairs", **arrs)


def golden_layer(name, d, n, ci, co, B, fs, bias, seed, weight_args=None, r=None):
    torch.manual_seed(seed)
    pts = torch.rand(n, d)
    x = torch.randn(B, ci, n, requires_grad=True)
    if weight_args is None:
        mesh = MeshHandler(pts).construct([n])
    else:
        mesh = MeshHandler(pts).construct([n], quad_args={"point_args": {}, "weight_args": weight_args})
    qc = QuadConv(spatial_dim=d, in_points=n, out_points=n, in_channels=ci, out_channels=co,
                  filter_seq=fs, bias=bias, output_same=True, decay_param=r)
    w = mesh.weights()
    w.retain_grad()
    # run the reference forward with the same quadrature weights tensor so d/dw can be read off
    mesh_w = mesh.weight_map
    mesh.weight_map = lambda points: w
    y = qc(mesh, x)
    mesh.weight_map = mesh_w
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 1))
    (y * g).sum().backward()
    arrs = dict(pts=pts, x=x, quad_w=w, grad_out=g, y=y, dx=x.grad, dquad_w=w.grad,
                pairs=qc.eval_indices, r=np.float64(qc.decay_param), ci=ci, co=co)
    ws = [m.weight for m in qc.filter if hasattr(m, "weight")]
    for l, W in enumerate(ws):
        arrs[f"W{l}"] = W
        arrs[f"dW{l}"] = W.grad
    if bias:
        arrs["bias"] = qc.bias
        arrs["dbias"] = qc.bias.grad
    for k, v in mesh.state_dict().items():
        arrs["mesh/" + k] = v
    for k, p in mesh.named_parameters():
        if p.grad is not None:
            arrs["meshgrad/" + k] = p.grad
    save(name, **arrs)


def golden_nested(name="layer_nested", d=2, n=200, ci=2, co=3, B=2, fs=(4,), seed=9):
    """filter_mode='nested' (quadconv.py:105-114): Ci*Co scalar MLPs whose outputs are concatenated along dim 0 and
    reshaped -- the reference's (scrambled) layout is part of the contract, so it is pinned here."""
    torch.manual_seed(seed)
    pts = torch.rand(n, d)
    x = torch.randn(B, ci, n, requires_grad=True)
    mesh = MeshHandler(pts).construct([n])
    qc = QuadConv(spatial_dim=d, in_points=n, out_points=n, in_channels=ci, out_channels=co, filter_seq=list(fs),
                  filter_mode='nested', bias=True, output_same=True)
    y = qc(mesh, x)
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 1))
    (y * g).sum().backward()
    arrs = dict(pts=pts, x=x, grad_out=g, y=y, dx=x.grad, pairs=qc.eval_indices, r=np.float64(qc.decay_param), ci=ci, co=co,
                bias=qc.bias, dbias=qc.bias.grad)
    for k, v in qc.state_dict().items():
        if k.startswith("filter."):
            arrs["sd/" + k] = v
    for k, p_ in qc.named_parameters():
        if k.startswith("filter."):
            arrs["grad/" + k] = p_.grad
    for k, v in mesh.state_dict().items():
        arrs["mesh/" + k] = v
    for k, p_ in mesh.named_parameters():
        if p_.grad is not None:
            arrs["meshgrad/" + k] = p_.grad
    save(name, **arrs)


def golden_pool():
    torch.manual_seed(7)
    pts = torch.rand(600, 2)
    mesh = MeshHandler(pts).construct([600, 0, 0], mirror=True, quad_map="mfnus")
    arrs = {"pts": pts, "point_seq": np.array(mesh.point_seq)}
    for l, p in enumerate(mesh._points):
        arrs[f"level{l}"] = p
    for l, dm in enumerate(mesh._downsample_map):
        keys = np.array(list(dm.keys()), dtype=np.int64)
        lens = np.array([len(v) for v in dm.values()], dtype=np.int64)
        flat = np.array([int(k) for v in dm.values() for k in v], dtype=np.int64)
        arrs[f"elim{l}/keys"], arrs[f"elim{l}/lens"], arrs[f"elim{l}/flat"] = keys, lens, flat
    B, C = 3, 5
    x = torch.randn(B, C, 600, requires_grad=True)
    # quantise so that ties occur inside groups (tie -> even gradient split)
    with torch.no_grad():
        x.copy_(torch.round(x * 2) / 2)
    pool0, pool1 = Mesh_MaxPool(), Mesh_MaxPool()
    up1, up0 = Mesh_MaxPool(adjoint=True), Mesh_MaxPool(adjoint=True)
    mesh.reset()
    cur = [mesh._current_index]
    y0 = pool0(mesh, x); cur.append(mesh._current_index)
    y1 = pool1(mesh, y0); cur.append(mesh._current_index)
    z1 = up1(mesh, y1); cur.append(mesh._current_index)
    z0 = up0(mesh, z1); cur.append(mesh._current_index)
    g = torch.randn(z0.shape, generator=torch.Generator().manual_seed(8))
    gy0 = torch.randn(y0.shape, generator=torch.Generator().manual_seed(9))
    ((z0 * g).sum() + (y0 * gy0).sum()).backward()
    arrs.update(x=x, y0=y0, y1=y1, z1=z1, z0=z0, g=g, gy0=gy0, dx=x.grad, cursor=np.array(cur),
                index_pool0=pool0.index[0, 0], index_pool1=pool1.index[0, 0],
                index_up1=up1.index[0, 0], index_up0=up0.index[0, 0])
    save("pool", **arrs)


def golden_mfnus():
    torch.manual_seed(11)
    arrs = {}
    for name, pts in (("u2d", torch.rand(500, 2)), ("u3d", torch.rand(400, 3))):
        kept, dm = MFNUS(pts.clone(), fc=1.65, K=10)
        arrs[name + "/pts"] = pts
        arrs[name + "/kept"] = kept
        arrs[name + "/keys"] = np.array(list(dm.keys()), dtype=np.int64)
        arrs[name + "/lens"] = np.array([len(v) for v in dm.values()], dtype=np.int64)
        arrs[name + "/flat"] = np.array([int(k) for v in dm.values() for k in v], dtype=np.int64)
    save("mfnus", **arrs)


def golden_autoencoder():
    """Two-level mirrored mesh, QuadConv(output_same) + pool + QuadConv + unpool + QuadConv, the
    C2 structure in miniature; exercises the shared MeshHandler cursor (mesh_handler.py:55-66)."""
    torch.manual_seed(21)
    pts = torch.rand(400, 2)
    mesh = MeshHandler(pts).construct([400, 0], mirror=True, quad_map="mfnus")
    n0, n1 = mesh.point_seq
    mk = lambda n, ci, co: QuadConv(spatial_dim=2, in_points=n, out_points=n, in_channels=ci,
                                    out_channels=co, filter_seq=[8, 8], output_same=True, bias=True)
    e0, e1, d0 = mk(n0, 2, 4), mk(n1, 4, 4), mk(n0, 4, 2)
    pool, up = Mesh_MaxPool(), Mesh_MaxPool(adjoint=True)
    x = torch.randn(2, 2, n0, requires_grad=True)
    mesh.reset()
    cur = [mesh._current_index]
    h = torch.sin(e0(mesh, x)); cur.append(mesh._current_index)
    h = pool(mesh, h); cur.append(mesh._current_index)
    h = torch.sin(e1(mesh, h)); cur.append(mesh._current_index)
    h = up(mesh, h); cur.append(mesh._current_index)
    y = d0(mesh, h); cur.append(mesh._current_index)
    loss = torch.nn.functional.mse_loss(y, x.detach())
    loss.backward()
    arrs = dict(pts=pts, x=x, y=y, loss=loss, dx=x.grad, cursor=np.array(cur),
                point_seq=np.array(mesh.point_seq))
    for nm, m in (("e0", e0), ("e1", e1), ("d0", d0)):
        for k, v in m.state_dict().items():
            arrs[f"{nm}/{k}"] = v
        for k, p in m.named_parameters():
            if p.grad is not None:
                arrs[f"{nm}grad/{k}"] = p.grad
    for k, v in mesh.state_dict().items():
        arrs["mesh/" + k] = v
    for k, p in mesh.named_parameters():
        if p.grad is not None:
            arrs["meshgrad/" + k] = p.grad
    dm = mesh._downsample_map[0]
    arrs["elim0/keys"] = np.array(list(dm.keys()), dtype=np.int64)
    arrs["elim0/lens"] = np.array([len(v) for v in dm.values()], dtype=np.int64)
    arrs["elim0/flat"] = np.array([int(k) for v in dm.values() for k in v], dtype=np.int64)
    save("autoencoder", **arrs)


if __name__ == "__main__":
    golden_pairs()
    golden_layer("layer_2d", d=2, n=300, ci=3, co=5, B=2, fs=[4, 4], bias=True, seed=3)
    golden_layer("layer_2d_wide", d=2, n=256, ci=8, co=8, B=4, fs=[8, 8], bias=False, seed=4)
    golden_layer("layer_3d", d=3, n=350, ci=4, co=2, B=3, fs=[8], bias=False, seed=5, r=0.22,
                 weight_args={"element_size": 4, "dimension": 3, "const": False})
    golden_layer("layer_1to2", d=2, n=100, ci=1, co=2, B=1, fs=[5, 6, 7], bias=True, seed=6)
    golden_pool()
    golden_mfnus()
    golden_autoencoder()
    golden_nested()
