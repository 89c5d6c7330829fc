"""CPU: host-side mirror of the reference API -- construction parity, state_dict layout, cursor
walk, MFNUS coarsening and the pool index, all against golden data from the live reference."""
import numpy as np
import pytest
import torch

from conftest import elim_map_from, load_golden

import pytorch_quadconv_b200 as q
from pytorch_quadconv_b200.mesh_pool import _group_of_fine_points
from pytorch_quadconv_b200.utils.mfnus import mfnus_eliminate


def test_mfnus_matches_reference():
    g = load_golden("mfnus")
    for nm in ("u2d", "u3d"):
        kept, dm = mfnus_eliminate(g[nm + "/pts"], fc=1.65, K=10)
        assert np.array_equal(kept, g[nm + "/kept"])
        assert np.array_equal(np.array(list(dm.keys())), g[nm + "/keys"])
        assert np.array_equal(np.array([len(v) for v in dm.values()]), g[nm + "/lens"])
        assert np.array_equal(np.array([int(k) for v in dm.values() for k in v]), g[nm + "/flat"])


def test_pool_group_index_matches_reference():
    g = load_golden("pool")
    n0, n1, n2 = [int(v) for v in g["point_seq"]]
    assert np.array_equal(_group_of_fine_points(elim_map_from(g, "elim0"), n0), g["index_pool0"])
    assert np.array_equal(_group_of_fine_points(elim_map_from(g, "elim1"), n1), g["index_pool1"])
    # adjoint uses the same last-keeper-wins table (mesh_pool.py:55-78)
    assert np.array_equal(_group_of_fine_points(elim_map_from(g, "elim1"), n1), g["index_up1"])
    assert np.array_equal(_group_of_fine_points(elim_map_from(g, "elim0"), n0), g["index_up0"])


def test_mesh_construct_and_cursor_match_reference():
    g = load_golden("pool")
    mesh = q.MeshHandler(torch.from_numpy(g["pts"])).construct([600, 0, 0], mirror=True, quad_map="mfnus")
    assert mesh.point_seq == [int(v) for v in g["point_seq"]]
    for l in range(3):
        assert np.array_equal(mesh._points[l].numpy(), g[f"level{l}"])
    # mirrored walk of the cursor, as recorded from the reference after each pool layer
    mesh.reset()
    walk = [mesh._current_index]
    for _ in range(4):
        mesh.step()
        walk.append(mesh._current_index)
    assert walk == [int(v) for v in g["cursor"]]
    mesh.reset()
    seen = []
    for _ in range(4):
        seen.append((mesh.input_points().shape[0], mesh.output_points().shape[0]))
        mesh.step()
    n0, n1, n2 = mesh.point_seq
    assert seen == [(n0, n1), (n1, n2), (n2, n1), (n1, n0)]


def test_single_level_step_raises_like_reference():
    mesh = q.MeshHandler(torch.rand(30, 2)).construct([30])
    with pytest.raises(ZeroDivisionError):     # mesh_handler.py:60-63 with one level
        mesh.step()


@pytest.mark.parametrize("name,kw", [
    ("layer_2d", dict(d=2, n=300, ci=3, co=5, fs=[4, 4], bias=True, seed=3)),
    ("layer_1to2", dict(d=2, n=100, ci=1, co=2, fs=[5, 6, 7], bias=True, seed=6)),
])
def test_same_seed_gives_reference_parameters(name, kw):
    """Same RNG stream as the reference: rand points, randn features, MeshHandler, then QuadConv."""
    L = load_golden(name)
    torch.manual_seed(kw["seed"])
    pts = torch.rand(kw["n"], kw["d"])
    _x = torch.randn(int(L["x"].shape[0]), kw["ci"], kw["n"])
    mesh = q.MeshHandler(pts).construct([kw["n"]])
    layer = q.QuadConv(spatial_dim=kw["d"], in_points=kw["n"], out_points=kw["n"], in_channels=kw["ci"],
                       out_channels=kw["co"], filter_seq=kw["fs"], bias=kw["bias"], output_same=True)
    assert np.array_equal(pts.numpy(), L["pts"])
    sd = layer.state_dict()
    assert list(sd.keys()) == ["bias"] + [f"filter.{2 * l}.weight" for l in range(len(kw["fs"]) + 1)]
    for l in range(len(kw["fs"]) + 1):
        assert np.array_equal(sd[f"filter.{2 * l}.weight"].numpy(), L[f"W{l}"])
    assert np.array_equal(sd["bias"].numpy(), L["bias"])
    msd = mesh.state_dict()
    for k in msd:
        assert np.array_equal(msd[k].numpy(), L["mesh/" + k]), k
    assert float(layer.decay_param) == float(L["r"])
    # quadrature weights from the learned map (mesh_handler.py:101-112) equal the reference's
    assert np.allclose(mesh.weights().detach().numpy(), L["quad_w"], rtol=0, atol=1e-6)


def test_ctor_validation():
    with pytest.raises(AssertionError):
        q.QuadConv(spatial_dim=0, in_points=4, out_points=4, in_channels=1, out_channels=1, filter_seq=[2])
    with pytest.raises(ValueError):
        q.QuadConv(spatial_dim=2, in_points=4, out_points=4, in_channels=1, out_channels=1, filter_seq=[2],
                   filter_mode="bogus")
    layer = q.QuadConv(spatial_dim=2, in_points=16, out_points=16, in_channels=1, out_channels=1, filter_seq=[2])
    assert float(layer.decay_param) == 1.0 and layer.bias is None and layer.cached is False
    assert q.MeshMaxPool is q.Mesh_MaxPool


def test_union_tile_tables_host_logic():
    """ops.build_union_tables (tables of the tensor-core path) against a brute-force construction: every pair's
    slot points at its own source point, unions are sorted, padded to multiples of 16 with -1, tiles follow
    the processing order."""
    import numpy as np
    import torch
    from pytorch_quadconv_b200 import ops
    rng = np.random.default_rng(0)
    n_dst, n_src = 70, 50
    lens = rng.integers(0, 9, n_dst)
    lens[5] = 0
    seg = np.zeros(n_dst + 1, dtype=np.int32)
    np.cumsum(lens, out=seg[1:])
    src = np.concatenate([np.sort(rng.choice(n_src, l, replace=False)) for l in lens]).astype(np.int32)
    order = rng.permutation(n_dst).astype(np.int32)
    for od in (None, order):
        ent, pptr, uptr, uidx, max_ku = ops.build_union_tables(torch.from_numpy(seg), torch.from_numpy(src),
                                                               None if od is None else torch.from_numpy(od), n_dst, n_src)
        ent, pptr, uptr, uidx = ent.numpy(), pptr.numpy(), uptr.numpy(), uidx.numpy()
        ntiles = (n_dst + 15) // 16
        assert uptr.shape == (ntiles + 1,) and uptr[0] == 0 and np.all(np.diff(uptr) % 32 == 0) and np.all(np.diff(uptr) >= 32)
        assert max_ku == np.diff(uptr).max()
        assert pptr[0] == 0 and pptr[-1] == len(src) and ent.shape == (len(src), 2)
        proc = np.arange(n_dst) if od is None else od
        seen = np.zeros(len(src), dtype=bool)
        for t in range(ntiles):
            members = proc[t * 16:(t + 1) * 16]
            want = np.unique(np.concatenate([src[seg[m]:seg[m + 1]] for m in members] + [np.zeros(0, np.int32)]))
            got = uidx[uptr[t]:uptr[t + 1]]
            assert np.array_equal(got[:len(want)], want) and np.all(got[len(want):] == -1)
            assert pptr[t + 1] - pptr[t] == sum(seg[m + 1] - seg[m] for m in members)
            for e, meta in ent[pptr[t]:pptr[t + 1]]:
                ol, slot = meta >> 16, meta & 0xffff
                m = members[ol]
                assert seg[m] <= e < seg[m + 1] and got[slot] == src[e] and not seen[e]
                seen[e] = True
        assert seen.all()


def test_fold_factor_rules():
    """Batch folding (ops.fold_factor): largest power of two s with s*max(Ci,Co) <= 32, B % s == 0 and B/s >= 32;
    never for wide filters (H > 8)."""
    from pytorch_quadconv_b200 import ops
    assert ops.fold_factor(256, 4, 8, 8) == 4        # 4 -> 8 channels, batch 256: [64, 16 -> 32]
    assert ops.fold_factor(256, 8, 4, 8) == 4
    assert ops.fold_factor(256, 16, 32, 8) == 1      # already 32 channels on one side
    assert ops.fold_factor(128, 4, 8, 8) == 4 and ops.fold_factor(64, 4, 8, 8) == 2 and ops.fold_factor(32, 4, 8, 8) == 1
    assert ops.fold_factor(96, 3, 5, 8) == 2         # 96 / 4 = 24 lanes would fall below 32
    assert ops.fold_factor(72, 4, 4, 8) == 2         # 72 % 4 == 0 but 72 / 4 = 18 < 32
    assert ops.fold_factor(256, 4, 8, 32) == 1       # wide filter: the wide path is not folded
    saved = ops.FOLD_MIN_BATCH
    try:
        ops.FOLD_MIN_BATCH = 0
        assert ops.fold_factor(256, 4, 8, 8) == 1    # switch
    finally:
        ops.FOLD_MIN_BATCH = saved


def test_flat_gradient_layout_roundtrip():
    """ops._flat_layout / _carve: 16-byte aligned pieces of one flat buffer, shapes preserved, no overlap."""
    import torch
    from pytorch_quadconv_b200 import ops
    shapes = [(8, 3), (8, 8), (100,), (1024, 8), ()]
    offs, sizes, total = ops._flat_layout(shapes)
    assert all(o % 4 == 0 for o in offs) and total >= sum(sizes)
    for (o, n), (o2, _) in zip(zip(offs, sizes), list(zip(offs, sizes))[1:]):
        assert o + n <= o2
    flat = torch.arange(total, dtype=torch.float32)
    views = ops._carve(flat, shapes)
    assert [tuple(v.shape) for v in views] == [tuple(s) for s in shapes]
    for v, o in zip(views, offs):
        assert v.data_ptr() == flat.data_ptr() + 4 * o
