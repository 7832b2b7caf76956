"""Times the two short-K GEMM shapes of the benchmark layer (mask projection: K=128 -> 384 columns; dgrad of the
post transform: K=128 -> 640 columns) at M = 2M rows.  MMA_GEMM_ARES=0 selects the streaming kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mma_b200 import tc_gemm as tg

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for (N, K) in [(384, 128), (640, 128)]:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    hi, lo = tg.split_weight(W)
    out = torch.empty(M, N, device="cuda")
    for _ in range(2):
        tg.linear(A, hi, lo, N, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tg.linear(A, hi, lo, N, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 4 * M * (N + K) / 1e9
    print(f"M={M} N={N} K={K}: {ms:.3f} ms  {gb / ms:.0f} GB/s  ares={os.environ.get('MMA_GEMM_ARES', '1')} "
          f"tma_store={os.environ.get('MMA_GEMM_TMA_STORE', '1')}")
