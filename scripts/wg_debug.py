import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mma_b200 import tc_gemm as tg
dev = "cuda:0"
torch.manual_seed(0)
M, N, K = 32, 128, 128
G = torch.zeros(M, N, device=dev); A = torch.zeros(M, K, device=dev)
G[3, 5] = 1.0; G[10, 70] = 2.0
A[3] = torch.arange(K, device=dev).float() + 1
A[10] = -(torch.arange(K, device=dev).float() + 1)
slabs = tg.make_slabs(M, 1, dev)
print("slabs", slabs.tolist())
for mode in (2, 0):
    part = tg.wgrad_partials(G, A, slabs, slabs.shape[0], mode=mode)
    torch.cuda.synchronize()
    p = part[0]
    nz = p.nonzero()
    print("mode", mode, "absmax", p.abs().max().item(), "nnz", nz.shape[0], "expected nnz", 2 * K)
    print(" rows with nonzeros:", sorted(set(nz[:, 0].tolist()))[:20])
    for r in sorted(set(nz[:, 0].tolist()))[:4]:
        print("  row", r, p[r, :12].tolist(), "...", p[r, 60:68].tolist())
ref = G.t() @ A
print("ref rows 5,70:", ref[5, :6].tolist(), ref[70, :6].tolist())
G = torch.randn(M, N, device=dev); A = torch.randn(M, K, device=dev)
part = tg.wgrad_partials(G, A, slabs, 1, mode=0)
ref = G.double().t() @ A.double()
print("random: absmax out", part.abs().max().item(), "ref", ref.abs().max().item(), "err", (part[0].double() - ref).abs().max().item())
