"""Error map of the tensor-memory-resident GEMM variant per (row block, 64-column sub-tile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mma_b200 import tc_gemm as tg

for (M, N, K) in [(77777, 640, 128), (77696, 640, 128), (128 * 148 * 3, 640, 128), (128 * 148 * 4, 640, 128), (128 * 148 * 5 + 128, 384, 128), (200000, 384, 128)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    hi, lo = tg.split_weight(W)
    C = tg.linear(A, hi, lo, N)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    err = (C.double() - ref).abs()
    mt, ns = (M + 127) // 128, (N + 63) // 64
    pad = torch.zeros(mt * 128, ns * 64, dtype=torch.float64, device="cuda")
    pad[:M, :N] = err
    emap = pad.view(mt, 128, ns, 64).amax(dim=(1, 3)) / ref.abs().max()
    bad = (emap > 1e-5).nonzero()
    print(f"M={M} N={N} K={K}: max rel err {emap.max().item():.2e}; bad (row block, sub-tile) pairs: {bad.shape[0]} of {mt * ns}")
    if bad.shape[0]:
        print("  first bad:", bad[:12].tolist(), " bad row blocks mod 148:", sorted(set((bad[:, 0] // 148).tolist()))[:10],
              " bad sub-tiles:", sorted(set(bad[:, 1].tolist())))
        r, c = bad[0].tolist()
        blk = pad[r * 128:(r + 1) * 128, c * 64:(c + 1) * 64] / ref.abs().max()
        print("  rows bad in first block:", (blk.amax(1) > 1e-5).nonzero().flatten().tolist()[:40])
        print("  cols bad in first block:", (blk.amax(0) > 1e-5).nonzero().flatten().tolist()[:70])
