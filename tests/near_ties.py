"""Test helper (CPU, plain torch): where is a min/max selection of the reference NUMERICALLY AMBIGUOUS?

At the layer level the messages of the CUDA path differ from the reference's by fp32 rounding (the mask Linear
is evaluated as P[dst] + Q[src] on 3xTF32 tensor cores instead of one sgemm over cat([x_i, x_j]),
mma_conv.py:146-152).  The forward is continuous in the messages, but the BACKWARD of min / max routes the
gradient to the arg edge (torch_scatter): when the best and the second-best message of a
(row, column) are closer than that rounding, the two evaluations may pick different edges and the gradients of
the two sources differ by O(1) of one edge's contribution -- for the reference against ITSELF in another fp32
summation order just as much (measured on CPU: config-4 shape, 3000 nodes / 48000 edges / hidden 128: one flip
in 6 M messages, two messages 2 ulp apart from different sources, moves dx by 1.3e-3 of its largest entry).

`ambiguous_entries` finds those entries on the oracle's side so that a whole-layer parity test can be strict
everywhere else.  Exact ties are NOT ambiguous: equal values arise from dropped elements (+-0, equal under the
strict comparison) and duplicate edges, are equal on both sides, and are resolved by the same first-occurrence
rule.
"""
import torch


def _second_best(m, index, n, arg, is_max):
    """Best value of every (row, col) after removing its arg edge.  m [E, F], arg [n, F] (E = empty)."""
    E, F = m.shape
    fill = float("-inf") if is_max else float("inf")
    work = m.clone()
    cols = torch.arange(F).expand(n, F)
    has = arg < E
    work[arg[has], cols[has]] = fill
    out = torch.full((n, F), fill, dtype=m.dtype)
    out.scatter_reduce_(0, index.view(-1, 1).expand(E, F), work, "amax" if is_max else "amin", include_self=True)
    return out


def ambiguous_entries(m, index, n, arg_min, arg_max, tol):
    """m [E, F] messages (oracle side, after dropout), index [E] destination of every edge, arg_* [n, F] the
    oracle's arg edges (E for empty rows).  Returns (rows, edges): destination rows that hold an entry whose best
    and second-best message differ by less than `tol` (absolute) without being equal, and the edges (best and
    runner-up candidates) whose sources' gradients a flip would change."""
    E, F = m.shape
    rows, edges = set(), set()
    for arg, is_max in ((arg_min, False), (arg_max, True)):
        if arg is None:
            continue
        has = arg < E
        cols = torch.arange(F).expand(n, F)
        best = torch.zeros((n, F), dtype=m.dtype)
        best[has] = m[arg[has], cols[has]]
        second = _second_best(m, index, n, arg, is_max)
        gap = (best - second).abs()
        bad = has & torch.isfinite(second) & (gap < tol) & (gap > 0)
        for r, c in torch.nonzero(bad).tolist():
            rows.add(r)
            edges.add(int(arg[r, c]))
            # every edge of the row within tol of the best is a candidate
            in_row = torch.nonzero(index == r).view(-1)
            near = in_row[(m[in_row, c] - best[r, c]).abs() < tol]
            edges.update(near.tolist())
    return sorted(rows), sorted(edges)
