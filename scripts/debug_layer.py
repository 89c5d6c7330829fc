import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pytorch_quadconv_b200 as q
from test_gpu_layer import _oracle_case, _StubMesh

def run(d, n, ci, co, B, fs, r, bias, n_out=None):
    c = _oracle_case(d, n, ci, co, B, fs, r, bias, seed=1000 + n + ci, n_out=n_out)
    No = c["out_pts"].shape[0]
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
    yd = y.detach().cpu().double(); yr = c["y"].double()
    # float64 oracle for an unbiased reference
    from oracle import quadconv_oracle as O
    y64 = O.quadconv_forward(c["out_pts"].double(), c["pts"].double(), c["qw"].detach().double(), [w.detach().double() for w in c["ws"]], c["x"].detach().double(), c["pairs"], ci, co, None if not bias else c["b"].detach().double())
    e = (yd - yr)
    print("cfg", d, n, ci, co, B, fs, "rel(ours,oracle32)=%.3e rel(ours,f64)=%.3e rel(oracle32,f64)=%.3e" % (float(e.norm()/yr.norm()), float((yd-y64).norm()/y64.norm()), float((yr-y64).norm()/y64.norm())))
    ea = e.abs()
    flat = ea.flatten().topk(5)
    for v, i in zip(flat.values, flat.indices):
        b, j, o = np.unravel_index(int(i), ea.shape)
        print("   top err %.3e at b=%d j=%d o=%d  ours=%.7f ref=%.7f f64=%.7f" % (float(v), b, j, o, float(yd[b,j,o]), float(yr[b,j,o]), float(y64[b,j,o])))
    (y * c["go"].cuda()).sum().backward()
    def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm())
    print("   dx %.3e dqw %.3e" % (rel(x.grad.cpu(), c["x"].grad), rel(w_dev.grad.cpu(), c["qw"].grad)), " dW", ["%.3e" % rel(m.weight.grad.cpu(), w.grad) for m, w in zip(lin, c["ws"])])

for cfg in [(2, 900, 16, 16, 8, [4, 4], 0.1, False), (3, 1500, 32, 32, 32, [8, 8], 0.16, True), (2, 600, 5, 7, 3, [6], 0.12, True),
            (2, 500, 4, 40, 40, [8, 8], 0.1, False), (2, 400, 64, 16, 2, [32, 32], 0.15, False), (1, 300, 2, 3, 4, [4, 4], 0.02, False),
            (2, 700, 6, 4, 5, [], 0.08, False), (3, 400, 3, 2, 1, [16, 4], 0.3, False)]:
    try:
        run(*cfg)
    except Exception as ex:
        print("cfg", cfg, "EXC", repr(ex)[:300])
run(2, 800, 8, 12, 6, [8, 8], 0.09, True, 300)
