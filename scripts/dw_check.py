"""Full-size C3 check of d(W_last): tcgen05 accumulators vs the FP32-pipe kernel (QC_NO_TC=1 in a second process)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.workload("c3")
dev = torch.device("cuda:0")
mesh, layer, x_host = bench.build_problem(w, dev)
x = x_host.to(dev).requires_grad_()
torch.manual_seed(1)
y = layer(mesh, x)
go = torch.randn_like(y)
(y * go).sum().backward()
lin = [m for m in layer.filter if isinstance(m, torch.nn.Linear)]
out = {"dwl": lin[-1].weight.grad.cpu().numpy(), "dw0": lin[0].weight.grad.cpu().numpy(), "dx": x.grad[:2, :, :2000].cpu().numpy(),
       "y": y[:2, :, :2000].detach().cpu().numpy()}
np.savez(sys.argv[1], **out)
print("saved", sys.argv[1])
