"""BASELINE config 1 end to end on the GPU: the reference's Cora node-classification run
(`node_classification/train.py:48-116`, README.md:70: `--aggregators mean,mean2 --dataset cora --lr=0.001
--epochs=200 --weight_decay=3e-4 --hidden=64 --dropout=0.75`) through the drop-in modules
(`mma_b200.node_classification.models.MMAConv` = GraphConvolution -> ReLU -> dropout -> MMA -> log_softmax).

    python scripts/train_cora.py [--epochs 200] [--dropout 0.75] [--graph]

The dataset comes from the committed data fixture tests/golden/cora_dataset.pt (what `utils.load_data("cora")`
returns; written by oracle/make_cora_fixture.py in the build container).  `--graph` replays the whole train step
(forward, loss, backward, Adam) as ONE CUDA graph: the layers' dropout seeds then live on the device and are advanced
inside the graph (`MMA.device_seed`), so every replay draws fresh masks like an eager epoch.  No CPU path: the layers
raise without CUDA.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "cora_dataset.pt")


def load_fixture(device):
    """-> add_all, adj (sparse COO), features [N, nfeat], labels, idx_train, idx_val, idx_test  (utils.py:119)."""
    d = torch.load(FIXTURE, weights_only=False)
    n = d["n"]
    x = torch.zeros(n, d["nfeat"])
    x[d["feat_row"].long(), d["feat_col"].long()] = 1.0
    rowptr, col = d["rowptr"].long(), d["col"].long()
    add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    adj = torch.sparse_coo_tensor(torch.stack([row, col]), torch.ones(col.numel()), (n, n)).coalesce()
    to = lambda t: t.long().to(device)
    return add_all, adj.to(device), x.to(device), to(d["labels"]), to(d["idx_train"]), to(d["idx_val"]), to(d["idx_test"])


def accuracy(output, labels):                                   # utils.py:131-135
    return (output.max(1)[1] == labels).double().mean().item()


def build_model(add_all, nfeat, nclass, args, device, init=None):
    from mma_b200.node_classification.models import MMAConv
    model = MMAConv(add_all, args.activation, args.k, nfeat, args.hidden, nclass, args.dropout,
                    args.aggregators.split(","), device)
    if init is not None:                                        # start from given parameters (parity runs)
        with torch.no_grad():
            for k, v in init.items():
                getattr(model, k).copy_(v.to(device))
    return model


def run(args, init=None, verbose=True):
    """train.py:69-116.  Returns (history [epochs, 4], (test loss, test accuracy), seconds per epoch)."""
    device = torch.device("cuda", 0)
    add_all, adj, x, labels, itr, iva, ite = load_fixture(device)
    torch.manual_seed(args.seed)
    model = build_model(add_all, x.shape[1], int(labels.max()) + 1, args, device, init)
    use_graph = bool(getattr(args, "graph", False))
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay,     # train.py:66-67
                           capturable=use_graph)
    hist = []

    def train_step():
        model.train(); opt.zero_grad()
        out = model(x, adj)
        loss = F.nll_loss(out[itr], labels[itr])
        loss.backward(); opt.step()
        return out, loss

    def eval_step():
        model.eval()
        with torch.no_grad():
            return model(x, adj)

    graph = None
    if use_graph:
        model.gc2.device_seed = True
        snap = [p.detach().clone() for p in model.parameters()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                          # warm-up outside the capture (handles, caches, Adam state)
            for _ in range(3):
                train_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                                  # back to the initial parameters and a fresh optimiser state
            for p, q in zip(model.parameters(), snap):
                p.copy_(q)
            for st in opt.state.values():
                st["step"].zero_(); st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        opt.zero_grad(set_to_none=False)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            model.train()
            for p in model.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            g_out = model(x, adj)
            g_loss = F.nll_loss(g_out[itr], labels[itr])
            g_loss.backward(); opt.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ep in range(args.epochs):
        if graph is not None:
            graph.replay(); out, loss = g_out, g_loss
        else:
            out, loss = train_step()
        acc = accuracy(out[itr], labels[itr])
        ev = eval_step()                                       # train.py:79-83: validation in eval mode
        hist.append((loss.item(), acc, F.nll_loss(ev[iva], labels[iva]).item(), accuracy(ev[iva], labels[iva])))
        if verbose and (ep % 20 == 19 or ep == 0):
            print(f"Epoch: {ep + 1:04d} loss_train: {hist[-1][0]:.4f} acc_train: {acc:.4f} "
                  f"loss_val: {hist[-1][2]:.4f} acc_val: {hist[-1][3]:.4f}", flush=True)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / max(args.epochs, 1)
    ev = eval_step()
    test = (F.nll_loss(ev[ite], labels[ite]).item(), accuracy(ev[ite], labels[ite]))
    if verbose:
        print(f"Test set results: loss= {test[0]:.4f} accuracy= {test[1]:.4f}   ({sec * 1e3:.2f} ms per epoch "
              f"incl. the evaluation pass{', train step as one CUDA graph' if graph is not None else ''})")
    return torch.tensor(hist, dtype=torch.float64), test, sec


def parser():
    p = argparse.ArgumentParser()                               # train.py:19-35, README.md:70 values as defaults
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--epochs", type=int, default=200)
    p.add_argument("--lr", type=float, default=0.001)
    p.add_argument("--weight_decay", type=float, default=3e-4)
    p.add_argument("--hidden", type=int, default=64)
    p.add_argument("--dropout", type=float, default=0.75)
    p.add_argument("--aggregators", type=str, default="mean,mean2")
    p.add_argument("--activation", type=str, default="new_sigmoid")
    p.add_argument("--k", type=int, default=2)
    p.add_argument("--graph", action="store_true", help="replay the train step as one CUDA graph (device-resident seeds)")
    return p


if __name__ == "__main__":
    run(parser().parse_args())
