"""Compare the tensor-core filter-MLP kernels with the scalar ones gradient by gradient (debug aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_quadconv_b200 as q  # noqa: E402

torch.manual_seed(5)
dev = torch.device("cuda:0")
d, fs, ci, co, n = 2, [32, 32], 8, 8, 777
pts = torch.rand(n, d)
mesh = q.MeshHandler(pts).construct([n]).to(dev)
layer = q.QuadConv(spatial_dim=d, in_points=n, out_points=n, in_channels=ci, out_channels=co, filter_seq=fs,
                   output_same=True, bias=False, decay_param=0.11).to(dev)
x = torch.randn(8, ci, n, device=dev)
wgt = torch.randn(8, co, n, device=dev)
res = []
for no_tc in (False, True):
    if no_tc:
        os.environ["QC_MLP_NO_TC"] = "1"
    else:
        os.environ.pop("QC_MLP_NO_TC", None)
    mesh.reset()
    ps = list(layer.named_parameters()) + list(mesh.named_parameters())
    for _, p in ps:
        p.grad = None
    xi = x.clone().requires_grad_()
    y = layer(mesh, xi)
    (y * wgt).sum().backward()
    torch.cuda.synchronize()
    res.append((y.detach().cpu(), xi.grad.cpu(), {k: p.grad.cpu() for k, p in ps if p.grad is not None}))
(y1, dx1, g1), (y0, dx0, g0) = res
rel = lambda a, b: float((a - b).norm() / b.norm())
print("y", rel(y1, y0), "dx", rel(dx1, dx0))
for k in g0:
    print(k, tuple(g0[k].shape), rel(g1[k], g0[k]))
a, b = g1["filter.0.weight"], g0["filter.0.weight"]
print(torch.cat([a, b], 1))
a, b = g1["filter.2.weight"], g0["filter.2.weight"]
print(((a - b).abs().amax(1) / b.abs().amax(1)))
