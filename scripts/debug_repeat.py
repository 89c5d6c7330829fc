import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pytorch_quadconv_b200 as q
from test_gpu_layer import _oracle_case, _StubMesh

def run(reps, d, n, ci, co, B, fs, r, bias, n_out=None, bwd=True):
    c = _oracle_case(d, n, ci, co, B, fs, r, bias, seed=1000 + n + ci, n_out=n_out)
    No = c["out_pts"].shape[0]
    errs = []
    first = None
    for rep in range(reps):
        layer = q.QuadConv(spatial_dim=d, in_points=n, out_points=No, in_channels=ci, out_channels=co,
                           filter_seq=fs, bias=bias, output_same=n_out is None, decay_param=r).cuda()
        with torch.no_grad():
            lin = [m for m in layer.filter if isinstance(m, torch.nn.Linear)]
            for m, w in zip(lin, c["ws"]):
                m.weight.copy_(w.detach())
            if bias: layer.bias.copy_(c["b"].detach())
        w_dev = c["qw"].detach().cuda().requires_grad_()
        mesh = _StubMesh(c["pts"].cuda(), c["out_pts"].cuda(), w_dev)
        x = c["x"].detach().cuda().requires_grad_()
        y = layer(mesh, x)
        if bwd:
            (y * c["go"].cuda()).sum().backward()
        yd = y.detach().cpu()
        e = float((yd.double() - c["y"].double()).norm() / c["y"].double().norm())
        errs.append(e)
        if first is None: first = yd
        elif not torch.equal(first, yd):
            dd = (first - yd).abs()
            print("   rep", rep, "differs from rep 0: max abs", float(dd.max()), "count", int((dd > 0).sum()))
        if e > 1e-6:
            ea = (yd.double() - c["y"].double()).abs()
            top = ea.flatten().topk(8)
            print("   rep", rep, "rel", e, [(np.unravel_index(int(i), ea.shape), float(v)) for v, i in zip(top.values, top.indices)])
    print("cfg", d, n, ci, co, B, fs, "reps", reps, "max rel %.3e min rel %.3e" % (max(errs), min(errs)))

if len(sys.argv) > 1 and sys.argv[1] == "small":
    run(1, 2, 300, 16, 16, 8, [4, 4], 0.15, True)
    run(1, 3, 300, 32, 8, 32, [8, 8], 0.3, False)
    print("small done")
else:
    run(30, 2, 900, 16, 16, 8, [4, 4], 0.1, False, bwd=False)
    run(10, 3, 1500, 32, 32, 32, [8, 8], 0.16, True, bwd=False)
    run(10, 2, 800, 8, 12, 6, [8, 8], 0.09, True, 300, bwd=False)
