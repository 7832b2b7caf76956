"""Kernel-level breakdown of one MMAConv fwd+bwd step of bench.py's config 4 (torch.profiler, CUDA
activities).  Under torchrun it profiles rank 0 of the destination-range sharded run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import mma_b200
from mma_b200.parallel import ShardedGraph, allreduce_grads
from torch.profiler import profile, ProfilerActivity

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
N, E, F = 2_000_000, 32_000_000, 128
AGGR = ["mean", "sum", "min", "max", "std"]
SCAL = ["identity", "amplification", "attenuation", "linear"]
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gen = torch.Generator(device=dev).manual_seed(42)
src = torch.randint(0, N, (E,), generator=gen, device=dev)
dst = torch.randint(0, N, (E,), generator=gen, device=dev)
hist = torch.bincount(torch.bincount(dst, minlength=N)).cpu()
torch.manual_seed(42)
conv = mma_b200.MMAConv(F, F, AGGR, SCAL, hist, towers=1, strict_reference=False).to(dev)
if world > 1:
    graph = ShardedGraph(src, dst, N, rank, world, balance="nodes")
    rows = graph.rows
    graph.local.build_transpose()
else:
    graph = mma_b200.Graph(src, dst, N, sort_rows=True)
    rows = N
del src, dst
x = torch.randn(rows, F, device=dev, generator=gen).requires_grad_()
gy = torch.randn(rows, F, device=dev, generator=gen)
params = list(conv.parameters()) + conv.mask_parameters()


def step():
    y = conv(x, graph)
    g = torch.autograd.grad(y, [x] + params, gy)
    if world > 1:
        for p, gg in zip(params, g[1:]):
            p.grad = gg
        allreduce_grads(params)
    return g


use_graph = "--graph" in sys.argv
conv.device_seed = use_graph
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
run = step
if use_graph:
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        step()
    run = cg.replay
    for _ in range(3):
        run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=60))
if world > 1:
    torch.cuda.synchronize(); dist.barrier()
    if use_graph:
        cg.reset()
    sys.stdout.flush()
    os._exit(0)
