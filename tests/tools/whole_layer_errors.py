"""GPU box, one process (the last 35 s of round 1's GPU budget): smoke(), the error figures of the whole-layer test at the
benchmark configuration, and the API / folded-layer tests of the graph-regression layer.  Uses the oracle (test tool).
    python tests/tools/whole_layer_errors.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()
import torch
import __graft_entry__ as g
g.smoke()
print("smoke s", time.time() - t0, flush=True)
import test_whole_layer_c4_gpu as T
import mma_b200
from oracle import restate
for mode, case in (("inject", T.INJECT_CASE), ("philox", T.PHILOX_CASE)):
    conv, src, dst, x, gy, keep = T.build_case(**case)
    n, E, F = case["n"], case["E"], case["F"]
    conv = conv.cuda(); conv.fold_min_rows = 16
    graph = mma_b200.Graph(src.cuda(), dst.cuda(), n, sort_rows=True)
    xg = x.cuda().requires_grad_()
    if mode == "inject":
        conv._inject_keep = keep.cuda()
    y = conv(xg, graph)
    if mode == "philox":
        keep = mma_b200.dropout_keep_scale(0.5, conv.last_seed, E, F, "cuda", graph=graph).cpu().view(E, 1, F)
    params = restate.weights_from_module(conv, clone=False).tensors()
    got = torch.autograd.grad(y, [xg] + params, gy.cuda())
    yr, ref, rows, _ = T.oracle_run(conv, src, dst, x, gy, keep)
    print(mode, "y", f"{T.rel_err(y, yr):.2e}", "grads", [f"{T.rel_err(a, b):.2e}" for a, b in zip(got, ref)], flush=True)
print("errs s", time.time() - t0, flush=True)
import pytest
rc = pytest.main(["-q", "-x", os.path.join(ROOT, "tests", "test_mmconv_gpu.py"), "-k", "api_and_errors or folded_post"])
print("pytest rc", rc, "total s", time.time() - t0, flush=True)
