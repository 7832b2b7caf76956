"""Config-5-shaped load-balance probe on one GPU: power-law in-degrees (alpha = 2.1), hidden 64."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mma_b200
from mma_b200 import _lib

def powerlaw_graph(N, E, cap, dev, seed=42):
    g = torch.Generator(device=dev).manual_seed(seed)
    rank = torch.arange(1, N + 1, device=dev, dtype=torch.float64)
    w = rank.pow(-1.0 / (2.1 - 1.0))
    deg = (w / w.sum() * E)
    deg = deg.clamp(max=cap)
    deg = (deg * (E / deg.sum())).clamp(max=cap).round().long()
    dst = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    dst = dst[torch.randperm(dst.numel(), device=dev, generator=g)]
    src = torch.randint(0, N, (dst.numel(),), device=dev, generator=g)
    return src, dst, deg

dev = torch.device("cuda", 0)
N, E, F, cap = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
src, dst, deg = powerlaw_graph(N, E, cap, dev)
E = dst.numel()
print(f"N={N} E={E} max deg={int(deg.max())} rows with deg>4096: {int((deg > 4096).sum())} deg0 rows: {int((deg == 0).sum())}")
hist = torch.bincount(deg).cpu()
AGGR = ["mean", "sum", "min", "max", "std"]; SCAL = ["identity", "amplification", "attenuation", "linear"]
torch.manual_seed(0)
conv = mma_b200.MMAConv(F, F, AGGR, SCAL, hist, towers=1, strict_reference=False).to(dev)
graph = mma_b200.Graph(src, dst, N, sort_rows=True)
x = torch.randn(N, F, device=dev).requires_grad_(); gy = torch.randn(N, F, device=dev)
params = list(conv.parameters()) + conv.mask_parameters()
def step():
    y = conv(x, graph); return torch.autograd.grad(y, [x] + params, gy)
for _ in range(3): step()
torch.cuda.synchronize()
_lib.reset_counters(); _lib.enable_timing(True)
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"step {dt*1e3:.2f} ms  ({E/dt/1e9:.3f} G edges/s)")
for k, (c, ms) in _lib.timing_summary().items():
    print(f"   {k:28s} {ms:8.3f} ms x {c // 5}")
