"""Configs 1-3 of BASELINE.json (L2-resident, launch-latency bound): microseconds per layer call fwd+bwd on the
GPU next to the oracle port of the reference on this box's host cores.  Run on a GPU box:
    python scripts/small_configs.py
Topologies come from tests/golden/planetoid_topology.pt (the real Cora / Pubmed graphs); features and weights are
synthetic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mma_b200
from mma_b200.node_classification.layers import MMA
from oracle import restate

dev = torch.device("cuda", 0)
ORDER = ["moment_3", "sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4", "max", "max2", "max3",
         "max4", "min", "min2", "min3", "min4", "softmax", "softmin", "std", "normalized_mean"]


def gpu_time(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3          # us


def graphed(fn):
    """Replays fn as one CUDA graph (what a training loop of these launch-bound layers should do)."""
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def graph_time(fn):
    try:
        return gpu_time(graphed(fn))
    except Exception as e:                      # not capture-safe (host sync inside): report eager only
        torch.cuda.synchronize()
        return float("nan")


def cpu_time(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e6


def nc_case(name, topo, Fd, C, names, p):
    rowptr, col = topo["rowptr"], topo["col"]
    n, E = rowptr.numel() - 1, col.numel()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(n, Fd, generator=g)
    add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
    ps = {nm: torch.nn.Parameter(torch.randn(2 * Fd, Fd, generator=g).to(dev) * 0.1) for nm in ORDER}
    W = torch.nn.Parameter((torch.randn(Fd, C, generator=g) * 0.1).to(dev)); b = torch.nn.Parameter(torch.zeros(C, device=dev))
    L = MMA(add_all, "new_sigmoid", 2, Fd, C, W, b, *[ps[nm] for nm in ORDER], p, names, dev)
    adj = restate.csr_to_sparse_adj(rowptr, col, n).to(dev)
    xg = x.to(dev).requires_grad_(); gy = torch.randn(n, C, generator=g).to(dev)
    plist = [xg, W, b] + [ps[nm] for nm in names]

    def step():
        y = L(xg, adj)
        return torch.autograd.grad(y, plist, gy)
    t_eager = gpu_time(step)
    t_graph = graph_time(step)
    # CPU oracle port, same shapes
    masks = {nm: ps[nm].detach().cpu().requires_grad_() for nm in names}
    Wc, bc = W.detach().cpu().requires_grad_(), b.detach().cpu().requires_grad_()
    xc = x.clone().requires_grad_(); adjc = adj.cpu(); gyc = gy.cpu()
    torch.set_num_threads(os.cpu_count() or 1)

    def cpu_step():
        y = restate.nc_forward(xc, adjc, rowptr, col, masks, Wc, bc, names, "new_sigmoid", p)
        torch.autograd.grad(y, [xc, Wc, bc] + [masks[nm] for nm in names], gyc)
    t_cpu = cpu_time(cpu_step)
    A = len(names)
    print(f"{name}: N={n} E={E} F={Fd} A={A}: GPU {t_eager:9.1f} us/layer fwd+bwd eager, {t_graph:7.1f} us as a CUDA graph ({A * E / min(t_eager, t_graph if t_graph == t_graph else t_eager):8.2f} M edge-aggregations/s)"
          f" | CPU port ({torch.get_num_threads()} threads) {t_cpu:11.1f} us  -> x{t_cpu / t_eager:.0f}")


def zinc_case():
    ei, _ = restate.zinc_like_batch(128, seed=42)
    n, E = int(ei.max()) + 1, ei.shape[1]
    deg = restate.degree_histogram(ei, n)
    torch.manual_seed(42)
    conv = mma_b200.MMAConv(75, 75, ["min", "max"], ["identity", "amplification", "linear"], deg, edge_dim=50, towers=5).to(dev)
    x = torch.randn(n, 75); ea = torch.randn(E, 50); gy = torch.randn(n, 75)
    xg, eag, gyg, eig = x.to(dev).requires_grad_(), ea.to(dev), gy.to(dev), ei.to(dev)
    params = list(conv.parameters()) + conv.mask_parameters()

    def step():
        y = conv(xg, eig, eag)
        return torch.autograd.grad(y, [xg] + params, gyg)
    t_eager = gpu_time(step)
    t_graph = graph_time(step)
    w = restate.weights_from_module(conv)
    for t in w.tensors():
        t.requires_grad_()
    xc = x.clone().requires_grad_()
    torch.set_num_threads(os.cpu_count() or 1)

    def cpu_step():
        y = restate.mmaconv_forward(w, xc, ei, ea, None)
        torch.autograd.grad(y, [xc] + w.tensors(), gy)
    t_cpu = cpu_time(cpu_step)
    print(f"c2 ZINC-shaped batch: N={n} E={E} towers=5 F_in=75 edge_dim=50: GPU {t_eager:9.1f} us/layer fwd+bwd eager, {t_graph:7.1f} us as a CUDA graph "
          f"({E / min(t_eager, t_graph if t_graph == t_graph else t_eager):6.2f} M edges/s) | CPU port ({torch.get_num_threads()} threads) {t_cpu:11.1f} us  -> x{t_cpu / t_eager:.0f}")


topo = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden",
                               "planetoid_topology.pt"))
nc_case("c1 Cora   MMA layer (mean,mean2; hidden 64 -> 7; dropout 0.75)", topo["cora"], 64, 7, ["mean", "mean2"], 0.75)
zinc_case()
nc_case("c3 Pubmed MMA layer (min,min2,min3,min4; hidden 16 -> 3)", topo["pubmed"], 16, 3, ["min", "min2", "min3", "min4"], 0.5)
