"""Kernel-level breakdown of one MMAConv fwd+bwd step of bench.py's config 4 (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mma_b200
from torch.profiler import profile, ProfilerActivity

N, E, F = 2_000_000, 32_000_000, 128
AGGR = ["mean", "sum", "min", "max", "std"]
SCAL = ["identity", "amplification", "attenuation", "linear"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(42)
src = torch.randint(0, N, (E,), generator=gen, device=dev)
dst = torch.randint(0, N, (E,), generator=gen, device=dev)
hist = torch.bincount(torch.bincount(dst, minlength=N)).cpu()
conv = mma_b200.MMAConv(F, F, AGGR, SCAL, hist, towers=1, strict_reference=False).to(dev)
graph = mma_b200.Graph(src, dst, N, sort_rows=True)
del src, dst
x = torch.randn(N, F, device=dev, generator=gen).requires_grad_()
gy = torch.randn(N, F, device=dev, generator=gen)
params = list(conv.parameters()) + conv.mask_parameters()


def step():
    y = conv(x, graph)
    return torch.autograd.grad(y, [x] + params, gy)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
