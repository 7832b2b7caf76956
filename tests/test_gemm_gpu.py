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


# ---------------------------------------------------------------------------------------------
# weight-space algebra of the fused layer (csrc/weight_prep.cu)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,ta,tb,splits", [(128, 2688, 128, False, False, 1), (128, 128, 2688, False, True, 21),
                                                (128, 2688, 128, True, False, 1), (75, 33, 50, False, False, 1),
                                                (5, 7, 300, True, True, 4), (64, 64, 16, False, False, 1)])
def test_small_gemm_vs_fp64(M, N, K, ta, tb, splits):
    from mma_b200 import tc_gemm as tg
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g).cuda()
    C = tg.small_gemm(A, B, trans_a=ta, trans_b=tb, k_splits=splits)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    assert C.shape == (M, N)
    assert relerr(C, ref) < 2e-6
    # operands that are column slices of wider matrices (W_x = W_post[:, :F])
    Aw = torch.randn((K, M + 8) if ta else (M, K + 8), generator=g).cuda()
    Av = Aw[:, :M] if ta else Aw[:, :K]
    C2 = tg.small_gemm(Av, B, trans_a=ta, trans_b=tb, k_splits=splits)
    ref2 = (Av.double().t() if ta else Av.double()) @ (B.double().t() if tb else B.double())
    assert relerr(C2, ref2) < 2e-6


@pytest.mark.parametrize("nb,S,names,F,Co", [(37, 4, ("mean", "sum", "min", "max", "std"), 128, 128),
                                             (3, 3, ("min", "max"), 20, 12), (1, 1, ("mean",), 4, 4)])
def test_compose_post_weight_and_its_gradient_vs_einsum(nb, S, names, F, Co):
    """W_c = sum_{s, a -> m} coef W_lin W_{s,a} (split for the 3xTF32 GEMMs, plus the transpose) and the gradient
    D = sum_b coef dW_c against the torch statement of the same algebra (fused_layer.py, MMA_WPREP=0 path)."""
    from mma_b200 import _lib, tc_gemm as tg
    from mma_b200.fused_layer import fold_blocks
    akinds = tuple(_lib.AGGR_KINDS[a] for a in names)
    mat, block_of, inv = fold_blocks(akinds)
    A, Am = len(akinds), len(mat)
    g = torch.Generator().manual_seed(nb + S + F)
    coef = torch.zeros(nb, S, A, Am)
    for a in range(A):
        coef[:, :, a, block_of[a]] = torch.rand(nb, S, generator=g) + 0.1
    coef = coef.cuda()
    WlW = torch.randn(Co, (S * A + 1) * F, generator=g).cuda()
    hi, lo, hiT, loT = tg.compose_post_weight(coef, WlW, F, F, True)
    ref = torch.einsum("bsam,scaf->bcmf", coef.double(), WlW[:, F:].double().view(Co, S, A, F).permute(1, 0, 2, 3)).reshape(nb, Co, Am * F)
    assert relerr(hi + lo, ref) < 1e-6
    assert torch.equal(hiT, hi.transpose(1, 2)) and torch.equal(loT, lo.transpose(1, 2))
    # hi is a TF32 number, lo the TF32 of the residual (low 13 mantissa bits clear)
    assert torch.equal((hi.view(torch.int32) & 0x1FFF), torch.zeros_like(hi, dtype=torch.int32))
    assert torch.equal((lo.view(torch.int32) & 0x1FFF), torch.zeros_like(lo, dtype=torch.int32))
    dWc = torch.randn(nb, Co, Am * F, generator=g).cuda()
    dX = torch.randn(Co, F, generator=g).cuda()
    D = tg.compose_post_wgrad(coef, torch.tensor(block_of, dtype=torch.int32).cuda(), dWc, dX, F, F)
    refD = torch.einsum("bsam,bcmf->csaf", coef.double(), dWc.double().view(nb, Co, Am, F)).reshape(Co, S * A * F)
    assert torch.equal(D[:, :F], dX)
    assert relerr(D[:, F:], refD) < 2e-6
    D0 = tg.compose_post_wgrad(coef, torch.tensor(block_of, dtype=torch.int32).cuda(), dWc, None, F, F)
    assert float(D0[:, :F].abs().max()) == 0.0 and torch.equal(D0[:, F:], D[:, F:])


def test_weight_prep_kernels_match_the_torch_formulation_of_the_layer(monkeypatch):
    """The fused layer with the library's weight-space kernels against the same layer with the torch formulation
    (MMA_WPREP=0): output and every gradient at 5e-6 (the summation order of the small products differs, and with it
    the last bits of the weights the 3xTF32 GEMMs see)."""
    import mma_b200
    from mma_b200 import fused_layer
    from mma_b200.synthetic import degree_histogram
    torch.manual_seed(5)
    N, E, F = 3000, 40000, 128
    ei = torch.randint(0, N, (2, E))
    conv = mma_b200.MMAConv(F, F, ["mean", "sum", "min", "max", "std"], ["identity", "amplification", "attenuation", "linear"],
                            degree_histogram(ei, N), towers=1, strict_reference=False).cuda()
    params = list(conv.parameters()) + conv.mask_parameters()
    graph = mma_b200.Graph(ei[0].cuda(), ei[1].cuda(), N, sort_rows=True)
    x = torch.randn(N, F).cuda().requires_grad_()
    gy = torch.randn(N, F).cuda()
    res = {}
    for flag in (True, False):
        monkeypatch.setattr(fused_layer, "WPREP", flag)
        conv._calls = 0
        y = conv(x, graph)
        res[flag] = [y.detach()] + [t.detach() for t in torch.autograd.grad(y, [x] + params, gy)]
    for a, b in zip(res[True], res[False]):
        assert relerr(a, b.double()) < 5e-6
