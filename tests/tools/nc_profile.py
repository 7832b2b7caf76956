import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from mma_b200.node_classification.layers import MMA
from oracle import restate
dev = torch.device("cuda", 0)
ORDER = ["moment_3", "sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4", "max", "max2", "max3",
         "max4", "min", "min2", "min3", "min4", "softmax", "softmin", "std", "normalized_mean"]
topo = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "planetoid_topology.pt"))
which, Fd, C, names, p = (sys.argv[1] if len(sys.argv) > 1 else "pubmed"), 16, 3, ["min", "min2", "min3", "min4"], 0.5
if which == "cora":
    Fd, C, names, p = 64, 7, ["mean", "mean2"], 0.75
rowptr, col = topo[which]["rowptr"], topo[which]["col"]
n = rowptr.numel() - 1
g = torch.Generator().manual_seed(42)
x = torch.randn(n, Fd, generator=g)
add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
ps = {nm: torch.nn.Parameter(torch.randn(2 * Fd, Fd, generator=g).to(dev) * 0.1) for nm in ORDER}
W = torch.nn.Parameter((torch.randn(Fd, C, generator=g) * 0.1).to(dev)); b = torch.nn.Parameter(torch.zeros(C, device=dev))
L = MMA(add_all, "new_sigmoid", 2, Fd, C, W, b, *[ps[nm] for nm in ORDER], p, names, dev)
adj = restate.csr_to_sparse_adj(rowptr, col, n).to(dev)
xg = x.to(dev).requires_grad_(); gy = torch.randn(n, C, generator=g).to(dev)
plist = [xg, W, b] + [ps[nm] for nm in names]
def step():
    y = L(xg, adj)
    return torch.autograd.grad(y, plist, gy)
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
