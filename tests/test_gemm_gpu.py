"""tcgen05 3xTF32 GEMMs (csrc/gemm_tf32x3.cu) against an fp64 product: they stand in for the reference's
fp32 Linears (graph_regression/mask_aggr.py:68, mma_conv.py:132-136), so the bar is fp32-level accuracy:
1e-5 of the largest magnitude (north_star tolerance), and in practice within ~4x of cuBLAS fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu
REL = 1e-5


def relerr(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1000, 128, 128), (4096, 256, 640), (5000, 384, 136),
                                   (333, 20, 100), (1, 4, 4), (129, 132, 36),
                                   # K <= 128 and N >= 256: the variant with the activation tile resident in tensor memory
                                   (4096, 384, 128), (77777, 640, 128), (1000, 256, 96), (300, 320, 100), (129, 260, 32),
                                   (200000, 384, 128)])
def test_linear_vs_fp64(M, N, K):
    from mma_b200 import tc_gemm as tg
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    hi, lo = tg.split_weight(W)
    C = tg.linear(A, hi, lo, N, bias=b)
    ref = A.double() @ W.double().t() + b.double()
    assert relerr(C, ref) < REL
    # plain TF32 would NOT pass: the 3x split is what buys fp32 accuracy
    if K >= 32:
        assert relerr(tg.linear(A, hi, lo, N, bias=b, mode=2), ref) > 10 * REL


@pytest.mark.parametrize("M,N,K", [(32, 128, 128), (4096, 128, 128), (100000, 128, 640), (77777, 384, 128),
                                   (5003, 20, 100), (1, 4, 4), (31, 8, 12)])
def test_wgrad_vs_fp64(M, N, K):
    from mma_b200 import tc_gemm as tg
    g = torch.Generator().manual_seed(M + N + K)
    G = torch.randn(M, N, generator=g).cuda()
    A = torch.randn(M, K, generator=g).cuda()
    dW = tg.wgrad(G, A)
    ref = G.double().t() @ A.double()
    assert relerr(dW, ref) < REL
    # deterministic: fixed slab order, no atomics
    assert torch.equal(dW, tg.wgrad(G, A))


def test_grouped_scatter_add_two_sources():
    """Grouped GEMM (one weight per row range), second A source concatenated along K, row-indexed
    addend and output-row scatter fused in the epilogue; wgrad with two G sources and ragged slabs."""
    from mma_b200 import tc_gemm as tg
    torch.manual_seed(0)
    M, K0, K1, N = 3000, 64, 96, 128
    A0, A1 = torch.randn(M, K0).cuda(), torch.randn(M, K1).cuda()
    Ws = (torch.randn(3, N, K0 + K1) / 12).cuda()
    perm = torch.randperm(M).cuda().int()
    add = torch.randn(M, N).cuda()
    segs = [(0, 1000, 0), (1000, 1100, 2), (1100, 3000, 1)]
    tab = [(r, hi_, w * N, 0) for lo_, hi_, w in segs for r in range(lo_, hi_, 128)]
    tab = torch.tensor(tab, dtype=torch.int32).cuda()
    hi, lo = tg.split_weight(Ws.view(3 * N, K0 + K1))
    C = tg.linear(A0, hi, lo, N, A1=A1, tile_tab=tab, out_map=perm, add=add)
    Acat = torch.cat([A0, A1], 1).double()
    ref = torch.empty(M, N, dtype=torch.float64).cuda()
    for lo_, hi_, w in segs:
        ref[lo_:hi_] = Acat[lo_:hi_] @ Ws[w].double().t()
    full = torch.empty_like(ref)
    full[perm.long()] = ref
    full += add.double()
    assert relerr(C, full) < REL
    # wgrad: G = [G0 | G1], slabs that end inside a 32-row block, one slot per slab
    G0, G1 = torch.randn(M, 128).cuda(), torch.randn(M, 40).cuda()
    slabs = torch.tensor([(0, 1000, 0, 0), (1000, 1003, 1, 0), (1003, 3000, 2, 0)], dtype=torch.int32).cuda()
    part = tg.wgrad_partials(G0, A1, slabs, 3, G1=G1)
    Gc = torch.cat([G0, G1], 1).double()
    for (lo_, hi_, slot, _) in slabs.tolist():
        assert relerr(part[slot], Gc[lo_:hi_].t() @ A1[lo_:hi_].double()) < REL
    seg = tg.reduce_slabs_segmented(part, torch.tensor([0, 2, 3], dtype=torch.int32).cuda())
    assert relerr(seg[0], Gc[:1003].t() @ A1[:1003].double()) < REL
    assert relerr(seg[1], Gc[1003:].t() @ A1[1003:].double()) < REL


def test_grouped_scatter_add_resident_a():
    """The tensor-memory-resident variant (K <= 128, N >= 256) with everything the epilogue can fuse: one weight
    per row range (tile table), bias, row-indexed addend (by output and by input row) and output-row scatter."""
    from mma_b200 import tc_gemm as tg
    torch.manual_seed(1)
    M, K, N = 3000, 128, 640
    A = torch.randn(M, K).cuda()
    Ws = (torch.randn(3, N, K) / 11).cuda()
    b = torch.randn(N).cuda()
    perm = torch.randperm(M).cuda().int()
    add = torch.randn(M, N).cuda()
    segs = [(0, 1000, 0), (1000, 1100, 2), (1100, 3000, 1)]
    tab = torch.tensor([(r, hi_, w * N, 0) for lo_, hi_, w in segs for r in range(lo_, hi_, 128)], dtype=torch.int32).cuda()
    hi, lo = tg.split_weight(Ws.view(3 * N, K))
    ref = torch.empty(M, N, dtype=torch.float64).cuda()
    for lo_, hi_, w in segs:
        ref[lo_:hi_] = A[lo_:hi_].double() @ Ws[w].double().t()
    ref += b.double()
    for by_input in (False, True):
        C = tg.linear(A, hi, lo, N, tile_tab=tab, out_map=perm, add=add, bias=b, add_by_input_row=by_input)
        full = torch.empty_like(ref)
        full[perm.long()] = ref + (add.double() if by_input else 0)
        if not by_input:
            full += add.double()
        assert relerr(C, full) < REL
    # same numbers as the streaming kernel to fp32 rounding, and deterministic
    C1 = tg.linear(A, hi, lo, N, tile_tab=tab)
    assert torch.equal(C1, tg.linear(A, hi, lo, N, tile_tab=tab))
    assert relerr(C1, ref - b.double()) < REL


def test_no_cpu_path():
    from mma_b200 import tc_gemm as tg
    with pytest.raises(RuntimeError):
        tg.split_weight(torch.randn(4, 4))
