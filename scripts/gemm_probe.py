"""GPU probe of the tcgen05 3xTF32 GEMMs: accuracy vs fp64 and timing vs cuBLAS fp32 (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mma_b200
from mma_b200 import tc_gemm as tg

dev = "cuda:0"
torch.manual_seed(0)


def relerr(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def check_nt(M, N, K, modes=(0, 1, 2)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    ref = A.double() @ W.double().t() + b.double()
    hi, lo = tg.split_weight(W)
    out = {}
    for m in modes:
        try:
            C = tg.linear(A, hi, lo, N, bias=b, mode=m)
            torch.cuda.synchronize()
            out[m] = relerr(C, ref)
        except Exception as e:
            out[m] = repr(e)[:200]
    cub = relerr(torch.addmm(b, A, W.t()), ref)
    print(f"NT  M={M} N={N} K={K}: relerr mode0/1/2 = {out}, cuBLAS fp32 = {cub:.2e}", flush=True)


def check_wgrad(M, N, K, modes=(0, 1, 2)):
    G = torch.randn(M, N, device=dev)
    A = torch.randn(M, K, device=dev)
    ref = G.double().t() @ A.double()
    out = {}
    for m in modes:
        try:
            dW = tg.wgrad(G, A, mode=m)
            torch.cuda.synchronize()
            out[m] = relerr(dW, ref)
        except Exception as e:
            out[m] = repr(e)[:200]
    cub = relerr(G.t() @ A, ref)
    print(f"WG  M={M} N={N} K={K}: relerr mode0/1/2 = {out}, cuBLAS fp32 = {cub:.2e}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "nt"):
    check_nt(128, 128, 32)
    check_nt(256, 128, 64)
    check_nt(1000, 128, 128)
    check_nt(4096, 256, 640)
    check_nt(5000, 384, 136)
    check_nt(333, 20, 100)
if which in ("all", "wg"):
    check_wgrad(32, 128, 128)
    check_wgrad(64, 128, 128)
    check_wgrad(4096, 128, 128)
    check_wgrad(100000, 128, 640)
    check_wgrad(77777, 384, 128)
    check_wgrad(5003, 20, 100)
if which in ("all", "scatter"):
    # grouped + scatter + add + two A sources
    M, K0, K1, N = 3000, 64, 96, 128
    A0 = torch.randn(M, K0, device=dev); A1 = torch.randn(M, K1, device=dev)
    Ws = torch.randn(3, N, K0 + K1, device=dev) / 12
    perm = torch.randperm(M, device=dev).int()
    add = torch.randn(M, N, device=dev)
    segs = [(0, 1000, 0), (1000, 1100, 2), (1100, 3000, 1)]
    tab = []
    for lo_, hi_, w in segs:
        r = lo_
        while r < hi_:
            tab.append((r, hi_, w * N, 0)); r += 128
    tab = torch.tensor(tab, dtype=torch.int32, device=dev)
    hi, lo = tg.split_weight(Ws.view(3 * N, K0 + K1))
    C = tg.linear(A0, hi, lo, N, A1=A1, tile_tab=tab, out_map=perm, add=add, mode=0)
    Acat = torch.cat([A0, A1], 1).double()
    ref = torch.empty(M, N, dtype=torch.float64, device=dev)
    for lo_, hi_, w in segs:
        ref[lo_:hi_] = Acat[lo_:hi_] @ Ws[w].double().t()
    full = torch.empty_like(ref); full[perm.long()] = ref
    full += add.double()
    print("grouped+scatter+add relerr:", relerr(C, full), flush=True)
if which in ("all", "time"):
    M = 2_000_000
    for (N, K) in ((384, 128), (128, 640), (128, 128), (640, 128)):
        A = torch.randn(M, K, device=dev)
        W = torch.randn(N, K, device=dev) / K ** 0.5
        hi, lo = tg.split_weight(W)
        C = torch.empty(M, N, device=dev)
        res = {}
        for m in (0, 1, 2):
            res[m] = timeit(lambda: tg.linear(A, hi, lo, N, out=C, mode=m))
        t_cub = timeit(lambda: torch.mm(A, W.t(), out=C))
        gb = (M * K + M * N) * 4 / 1e9
        print(f"time NT M={M} N={N} K={K}: mode0 {res[0]:.3f} ms mode1 {res[1]:.3f} ms mode2 {res[2]:.3f} ms | cuBLAS {t_cub:.3f} ms"
              f" | {gb:.2f} GB -> mode0 {gb / res[0]:.2f} TB/s, {2 * M * N * K / res[0] / 1e9:.1f} TFLOP/s", flush=True)
        del A, C
    for (N, K) in ((128, 640), (384, 128), (128, 128)):
        G = torch.randn(M, N, device=dev)
        A = torch.randn(M, K, device=dev)
        res = {}
        for m in (0, 1, 2):
            res[m] = timeit(lambda: tg.wgrad(G, A, mode=m))
        t_cub = timeit(lambda: torch.mm(G.t(), A))
        gb = (M * K + M * N) * 4 / 1e9
        print(f"time WG M={M} N={N} K={K}: mode0 {res[0]:.3f} ms mode1 {res[1]:.3f} ms mode2 {res[2]:.3f} ms | cuBLAS {t_cub:.3f} ms"
              f" | {gb:.2f} GB -> mode0 {gb / res[0]:.2f} TB/s", flush=True)
        del G, A
