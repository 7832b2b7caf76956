"""A/B of the streaming GEMM kernel (gemm_nt_kernel) at the two shapes the benchmark layer runs it on -- the grouped
post transform (K = 512 -> 128 columns) and the mask dgrad ([dPQ | dOut]: K = 256 + 128 -> 128 columns) -- at M = 2M
rows, with the 3xTF32 error against fp64 on the first rows.  Environment switches of the library (MMA_GEMM_N256, ...)
are read once per process: run this once per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mma_b200 import tc_gemm as tg

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tag = " ".join(f"{k}={os.path.basename(v)}" for k, v in sorted(os.environ.items()) if k.startswith(("MMA_GEMM", "MMA_B200_LIB")))
torch.manual_seed(0)
for (K0, K1, N) in [(512, 0, 128), (256, 128, 128)]:
    A0 = torch.randn(M, K0, device="cuda")
    A1 = torch.randn(M, K1, device="cuda") if K1 else None
    W = torch.randn(N, K0 + K1, device="cuda") / (K0 + K1) ** 0.5
    hi, lo = tg.split_weight(W)
    out = torch.empty(M, N, device="cuda")
    for _ in range(3):
        tg.linear(A0, hi, lo, N, A1=A1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tg.linear(A0, hi, lo, N, A1=A1, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = 4096
    Af = A0[:n].double() if A1 is None else torch.cat([A0[:n], A1[:n]], 1).double()
    ref = Af @ W.double().t()
    err = ((out[:n].double() - ref).abs().max() / ref.abs().max()).item()
    gb = 4 * M * (N + K0 + K1) / 1e9
    print(f"M={M} K={K0}+{K1} N={N}: {ms:.3f} ms  {1000 * gb / ms:.0f} GB/s  relerr {err:.2e}  [{tag}]", flush=True)

# the tensor-memory-resident kernel (K = 128: mask projection -> 384 columns, post dgrad -> 512 columns)
for (K, N) in [(128, 384), (128, 512)]:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    hi, lo = tg.split_weight(W)
    out = torch.empty(M, N, device="cuda")
    for _ in range(3):
        tg.linear(A, hi, lo, N, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tg.linear(A, hi, lo, N, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ref = A[:4096].double() @ W.double().t()
    err = ((out[:4096].double() - ref).abs().max() / ref.abs().max()).item()
    print(f"M={M} K={K} N={N} (resident A): {ms:.3f} ms  {4 * M * (N + K) / 1e6 / ms:.0f} GB/s  relerr {err:.2e}  [{tag}]", flush=True)

# the weight-gradient kernel (post transform: dO [M,128]^T Z [M,512]; mask: [dPQ | dOut] [M,256+128]^T x [M,128])
for (N0, N1, K) in [(128, 0, 512), (256, 128, 128)]:
    G0 = torch.randn(M, N0, device="cuda")
    G1 = torch.randn(M, N1, device="cuda") if N1 else None
    A = torch.randn(M, K, device="cuda")
    for _ in range(3):
        dW = tg.wgrad(G0, A, G1=G1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dW = tg.wgrad(G0, A, G1=G1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    m = 200_000
    Gf = G0[:m] if G1 is None else torch.cat([G0[:m], G1[:m]], 1)
    ref = Gf.double().t() @ A[:m].double()
    chk = tg.wgrad(G0[:m], A[:m], G1=None if G1 is None else G1[:m])
    err = ((chk.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"M={M} wgrad N={N0}+{N1} K={K}: {ms:.3f} ms (incl. slab sum)  {4 * M * (N0 + N1 + K) / 1e6 / ms:.0f} GB/s  relerr {err:.2e}  [{tag}]", flush=True)
