import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mma_b200 import MMAConv, Graph
torch.manual_seed(0)
n, E, Fd = 3000, 40000, 64
g = torch.Generator().manual_seed(21)
src = torch.randint(0, n, (E,), generator=g); dst = torch.randint(0, n - 3, (E,), generator=g)
hist = torch.bincount(torch.bincount(dst, minlength=n))
for aggr, fold in ((["mean", "max", "std"], 32), (["mean", "sum", "min", "max", "std"], 32), (["mean", "max", "std"], 100000), (["mean", "max", "std"], 1)):
    conv = MMAConv(Fd, Fd, aggr, ["identity", "amplification"], hist, towers=1, strict_reference=False).cuda()
    conv.fold_min_rows = fold
    conv.device_seed = True
    graph = Graph(src.cuda(), dst.cuda(), n, sort_rows=True)
    x = torch.randn(n, Fd).cuda().requires_grad_()
    gy = torch.randn(n, Fd).cuda()
    params = list(conv.parameters()) + conv.mask_parameters()
    def fwd():
        return conv(x, graph)
    def step():
        y = conv(x, graph)
        return (y,) + torch.autograd.grad(y, [x] + params, gy)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for name, fn in (("fwd", fwd), ("fwd+bwd", step)):
        try:
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                o = fn()
            cg.replay(); torch.cuda.synchronize()
            print(aggr, fold, name, "OK")
        except Exception as e:
            print(aggr, fold, name, "FAILED:", str(e).split("\n")[0])
            torch.cuda.synchronize()
