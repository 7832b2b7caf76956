"""all-gather / reduce-scatter bandwidth at the sizes of the sharded layer (torchrun, one rank per GPU)."""
import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, F = 2_000_000 // world, 128
x = torch.randn(rows, F, device=dev); out = torch.empty(world * rows, F, device=dev)
big = torch.randn(world * rows, F, device=dev); rs = torch.empty(rows, F, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t_ag = timeit(lambda: dist.all_gather_into_tensor(out, x))
t_rs = timeit(lambda: dist.reduce_scatter_tensor(rs, big))
if rank == 0:
    gb = out.numel() * 4 / 1e9
    print(f"world {world} PROTO={os.environ.get('NCCL_PROTO')} ALGO={os.environ.get('NCCL_ALGO')}: all_gather {gb:.2f} GB out: {t_ag:.3f} ms "
          f"({gb * (world - 1) / world / t_ag * 1e3:.0f} GB/s rx per GPU) | reduce_scatter: {t_rs:.3f} ms "
          f"({gb * (world - 1) / world / t_rs * 1e3:.0f} GB/s)")
dist.destroy_process_group()
