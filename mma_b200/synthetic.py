"""Synthetic inputs of BASELINE.json's configurations: pure index generation, CPU or CUDA, no
reference arithmetic.  The reference itself downloads ZINC (graph_regression/mma.py:46-54) and unpickles Planetoid
(node_classification/utils.py:33-119); there is no network here, so the benchmark and the tests draw graphs of the
same SHAPES from these generators (the Cora / Pubmed topologies are committed fixtures, tests/golden/)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor


def zinc_like_batch(num_graphs: int = 128, seed: int = 42) -> Tuple[Tensor, Tensor]:
    """Config 2: a block-diagonal batch of molecule-shaped graphs -- n_g = clip(round(N(23.2, 4.5^2)), 9, 37) nodes,
    a random spanning tree plus ring-closure edges up to ~1.07 n_g bonds, maximum degree 4, both directions.
    Returns (edge_index [2, E] int64, batch [N] int64), nodes numbered graph by graph."""
    g = torch.Generator().manual_seed(seed)
    srcs, dsts, batch = [], [], []
    base = 0
    for gi in range(num_graphs):
        n = int(torch.clamp(torch.round(torch.randn((), generator=g) * 4.5 + 23.2), 9, 37))
        degc = [0] * n
        und = set()
        for v in range(1, n):
            for _ in range(64):
                u = int(torch.randint(0, v, (), generator=g))
                if degc[u] < 4:
                    break
            else:
                u = min(range(v), key=lambda k: degc[k])
            und.add((u, v)); degc[u] += 1; degc[v] += 1
        extra = max(0, int(round(1.07 * n)) - (n - 1))
        tries = 0
        while extra > 0 and tries < 200:
            tries += 1
            u = int(torch.randint(0, n, (), generator=g)); v = int(torch.randint(0, n, (), generator=g))
            if u == v:
                continue
            a, b = min(u, v), max(u, v)
            if (a, b) in und or degc[a] >= 4 or degc[b] >= 4:
                continue
            und.add((a, b)); degc[a] += 1; degc[b] += 1; extra -= 1
        for (u, v) in sorted(und):
            srcs += [base + u, base + v]; dsts += [base + v, base + u]
        batch += [gi] * n
        base += n
    return torch.tensor([srcs, dsts], dtype=torch.int64), torch.tensor(batch, dtype=torch.int64)


def degree_histogram(edge_index: Tensor, n: int) -> Tensor:
    """The `deg` argument of MMAConv as graph_regression/mma.py:57-60 builds it: bincount of the in-degrees."""
    return torch.bincount(torch.bincount(edge_index[1], minlength=n))


def powerlaw_edges(N: int, E: int, dev, seed: int = 42, alpha: float = 2.1) -> Tuple[Tensor, Tensor]:
    """Config 5: in-degree of the node of rank r proportional to r^(-1/(alpha-1)), scaled to ~E edges, largest degree
    capped at E/200 (10^6 for config 5); the ranks are dealt to RANDOM node ids (ids carry no locality, as in a hashed
    id space), sources uniform.  Returns (src, dst) int64 on `dev`."""
    g = torch.Generator(device=dev).manual_seed(seed)
    cap = max(E // 200, 1)
    r = torch.arange(1, N + 1, device=dev, dtype=torch.float64)
    w = r.pow(-1.0 / (alpha - 1.0))
    deg = (w / w.sum() * E).clamp(max=cap)
    deg = (deg * (E / deg.sum())).clamp(max=cap).round().long()
    ids = torch.randperm(N, device=dev, generator=g)
    dst = torch.repeat_interleave(ids, deg)
    del r, w, deg, ids
    dst = dst[torch.randperm(dst.numel(), device=dev, generator=g)]
    src = torch.randint(0, N, (dst.numel(),), device=dev, generator=g)
    return src, dst


def uniform_edges(N: int, E: int, dev, seed: int = 42) -> Tuple[Tensor, Tensor]:
    """Config 4: src, dst ~ U[0, N) iid (multi-edges and self loops kept)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g, device=dev)
    dst = torch.randint(0, N, (E,), generator=g, device=dev)
    return src, dst


def csr_to_sparse_adj(rowptr: Tensor, col: Tensor, n: int) -> Tensor:
    """Binary sparse COO adjacency [n, n] of a neighbour-list CSR (what node_classification/utils.py:139-146 hands to
    the layers)."""
    row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), (rowptr[1:] - rowptr[:-1]).to(torch.int64))
    return torch.sparse_coo_tensor(torch.stack([row, col.to(torch.int64)]), torch.ones(col.numel()), (n, n)).coalesce()
