"""TEST INFRASTRUCTURE ONLY.  BASELINE config 1 end to end: the real Cora dataset as a compact data
fixture, plus the training trajectory of the VERBATIM reference layers on it.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_cora_fixture

Writes tests/golden/cora_dataset.pt
  * `feat_row/feat_col` -- the non-zeros of the [2708, 1433] binary bag-of-words matrix, `labels`,
    `idx_train/idx_val/idx_test`, `rowptr/col`: what `utils.load_data("cora")` returns
    (node_classification/utils.py:33-119, restated for current scipy / networkx: `adj[i].nonzero()[1]`
    -> CSR slices, `np.bool` -> bool; the values are data, not source);
and tests/golden/cora_train_ref.pt
  * the initial parameters (seed 42, the layers' own `reset_parameters`) and, per epoch, `loss_train,
    acc_train, loss_val, acc_val` of `train.py:69-96`'s loop run with the reference's
    `GraphConvolution` / `MMA` classes loaded verbatim (oracle/ref_shims.py) in the model of
    `models.py:64-68`, README.md:70 hyper-parameters (hidden 64, mean,mean2, lr 1e-3, wd 3e-4) except
    dropout = 0 so that the run is deterministic and a CUDA run can be compared epoch by epoch.
"""
from __future__ import annotations

import os
import pickle
import sys
import time
import warnings

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

from . import ref_shims
from .make_golden import NC_PARAM_ORDER, OUT, _save, planetoid_csr

DATA = os.path.join(ref_shims.REF_ROOT, "node_classification", "data")


def load_cora():
    """utils.py:33-119 for dataset == 'cora'."""
    objs = []
    for nm in ["x", "y", "tx", "ty", "allx", "ally"]:
        with open(os.path.join(DATA, f"ind.cora.{nm}"), "rb") as fh:
            objs.append(pickle.load(fh, encoding="latin1"))
    x, y, tx, ty, allx, ally = objs
    test_idx_reorder = [int(l.strip()) for l in open(os.path.join(DATA, "ind.cora.test.index"))]   # utils.py:15-20
    test_idx_range = np.sort(test_idx_reorder)
    features = sp.vstack((allx, tx)).tolil()                                                     # :62
    features[test_idx_reorder, :] = features[test_idx_range, :]                                  # :66
    labels = np.vstack((ally, ty))                                                               # :69
    labels[test_idx_reorder, :] = labels[test_idx_range, :]                                      # :71
    idx_test = test_idx_range.tolist()                                                           # :74
    idx_train = list(range(len(y) + 1068))                                                       # :78
    idx_val = list(range(len(y) + 1068, len(y) + 1068 + 500))                                    # :79
    feats = features.tocoo()
    assert np.all(feats.data == 1.0), "Cora's bag-of-words features are binary"
    rowptr, col = planetoid_csr("cora")                                                          # :67, :98-100
    return {"n": features.shape[0], "nfeat": features.shape[1],
            "feat_row": torch.from_numpy(feats.row.astype(np.int16)), "feat_col": torch.from_numpy(feats.col.astype(np.int16)),
            "labels": torch.from_numpy(np.where(labels)[1].astype(np.int8)),                      # :113
            "idx_train": torch.tensor(idx_train, dtype=torch.int16), "idx_val": torch.tensor(idx_val, dtype=torch.int16),
            "idx_test": torch.tensor(idx_test, dtype=torch.int16),
            "rowptr": rowptr.to(torch.int32), "col": col.to(torch.int32)}


def dense_features(d):
    x = torch.zeros(d["n"], d["nfeat"])
    x[d["feat_row"].long(), d["feat_col"].long()] = 1.0
    return x


class RefNet(torch.nn.Module):
    """models.py:10-68 with the reference's own layer classes; parameters via torch.empty (the
    `torch.cuda.FloatTensor` constructors of models.py:17-43 need a GPU)."""

    def __init__(self, layers, add_all, activation, k, nfeat, nhid, nclass, dropout, aggregator_list):
        super().__init__()
        new = lambda *s: torch.nn.Parameter(torch.empty(*s))
        self.weight0, self.bias0, self.weight1, self.bias1 = new(nfeat, nhid), new(nhid), new(nhid, nclass), new(nclass)
        for nm in NC_PARAM_ORDER:
            setattr(self, "weight_" + nm, new(2 * nhid, nhid))
        self.gc1 = layers.GraphConvolution(nfeat, nhid, self.weight0, self.bias0, "cpu")
        self.gc2 = layers.MMA(add_all, activation, k, nhid, nclass, self.weight1, self.bias1,
                              *[getattr(self, "weight_" + nm) for nm in NC_PARAM_ORDER], dropout, aggregator_list, "cpu")
        self.dropout = dropout

    def forward(self, x, adj):
        x = F.relu(self.gc1(x, adj))
        x = F.dropout(x, self.dropout, training=self.training)
        return F.log_softmax(self.gc2(x, adj), dim=1)


def accuracy(output, labels):                                                                    # utils.py:131-135
    return (output.max(1)[1] == labels).double().mean().item()


def main(epochs: int = 12, dropout: float = 0.0, out_name: str = "cora_train_ref.pt"):
    from . import restate
    assert ref_shims.reference_available(), "needs /root/reference"
    d = load_cora()
    if dropout == 0.0:
        _save("cora_dataset.pt", d)
    layers, _ = ref_shims.load_node_classification("cpu")
    rowptr, col = d["rowptr"].long(), d["col"].long()
    n = d["n"]
    add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
    adj = restate.csr_to_sparse_adj(rowptr, col, n)
    x, labels = dense_features(d), d["labels"].long()
    itr, iva = d["idx_train"].long(), d["idx_val"].long()
    np.random.seed(42); torch.manual_seed(42)                                                    # train.py:48-49
    names = ["mean", "mean2"]
    model = RefNet(layers, add_all, "new_sigmoid", 2, d["nfeat"], 64, 7, dropout, names)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()
            if k in ("weight0", "bias0", "weight1", "bias1") or k in ["weight_" + nm for nm in names]}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=3e-4)                       # train.py:66-67
    hist, secs = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for ep in range(epochs):
            t0 = time.time()
            model.train(); opt.zero_grad()
            out = model(x, adj)
            loss = F.nll_loss(out[itr], labels[itr]); acc = accuracy(out[itr], labels[itr])
            loss.backward(); opt.step()
            model.eval()
            with torch.no_grad():
                out = model(x, adj)
            hist.append((loss.item(), acc, F.nll_loss(out[iva], labels[iva]).item(), accuracy(out[iva], labels[iva])))
            secs.append(time.time() - t0)
            print(f"epoch {ep + 1:3d} loss_train {hist[-1][0]:.6f} acc_train {acc:.4f} loss_val {hist[-1][2]:.6f} "
                  f"acc_val {hist[-1][3]:.4f}  {time.time() - t0:.1f}s", flush=True)
    ite = d["idx_test"].long()
    model.eval()                                                                                 # train.py:98-106
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        out = model(x, adj)
    test = (F.nll_loss(out[ite], labels[ite]).item(), accuracy(out[ite], labels[ite]))
    print(f"test loss {test[0]:.4f} accuracy {test[1]:.4f}")
    rec = {"names": names, "activation": "new_sigmoid", "k": 2, "hidden": 64, "lr": 1e-3, "weight_decay": 3e-4,
           "dropout": dropout, "history": torch.tensor(hist, dtype=torch.float64), "test": test,
           "sec_per_epoch": sum(secs) / len(secs), "threads": torch.get_num_threads()}
    if dropout == 0.0:
        rec["init"] = init          # the deterministic run is compared epoch by epoch from identical parameters
    _save(out_name, rec)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        # README.md:70 as is (dropout 0.75, 200 epochs): the always-on mask dropout makes the run stochastic, so only
        # its accuracy band is comparable
        main(200, 0.75, "cora_train_ref_full.pt")
    else:
        main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
