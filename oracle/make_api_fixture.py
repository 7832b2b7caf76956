"""TEST INFRASTRUCTURE (never imported by mma_b200/): records the public API of the reference's hot-path modules
-- SURVEY.md 8(b), the drop-in boundary -- as a small JSON fixture, so that the signature parity test also runs
where /root/reference does not exist (the GPU box).

The reference files are only PARSED (`ast`), never imported or copied: for every class / function on the path
the fixture keeps the parameter names in order, the source text of their defaults, and the *args / **kwargs
names.  Regenerate with `python oracle/make_api_fixture.py` (writes tests/golden/reference_api.json).

    graph_regression/mma_conv.py      class MMAConv            (:20-199)
    graph_regression/mask_aggr.py     class MaskAggregateLinear (:7-68)
    node_classification/layers.py     class GraphConvolution (:12-51), class MMA (:54-873)
    node_classification/scalers.py    avg_d_log/avg_d_exp/scale_* (:10-62), SCALERS (:64)
    node_classification/models.py     class MMAConv            (:10-68)
"""
from __future__ import annotations

import ast
import json
import os
import sys

REF = os.environ.get("MMA_REFERENCE", "/root/reference")
FILES = ["graph_regression/mma_conv.py", "graph_regression/mask_aggr.py", "node_classification/layers.py",
         "node_classification/scalers.py", "node_classification/models.py"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_api.json")


def _sig(fn: ast.FunctionDef) -> dict:
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return {"line": fn.lineno,
            "params": [[n, d] for n, d in zip(pos, defaults)],
            "kwonly": [[x.arg, None if d is None else ast.unparse(d)] for x, d in zip(a.kwonlyargs, a.kw_defaults)],
            "vararg": a.vararg.arg if a.vararg else None, "kwarg": a.kwarg.arg if a.kwarg else None}


def _dict_keys(node: ast.AST):
    if isinstance(node, ast.Dict):
        return [k.value for k in node.keys if isinstance(k, ast.Constant)]
    return None


def collect(ref: str = REF) -> dict:
    api = {}
    for rel in FILES:
        tree = ast.parse(open(os.path.join(ref, rel)).read())
        entry = {"classes": {}, "functions": {}, "dicts": {}}
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                entry["classes"][node.name] = {
                    "line": node.lineno, "bases": [ast.unparse(b) for b in node.bases],
                    "methods": {m.name: _sig(m) for m in node.body if isinstance(m, ast.FunctionDef)}}
            elif isinstance(node, ast.FunctionDef):
                entry["functions"][node.name] = _sig(node)
            elif isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                keys = _dict_keys(node.value)
                if keys is not None:
                    entry["dicts"][node.targets[0].id] = {"line": node.lineno, "keys": keys}
        # module-level dicts built inside __init__ that the callers index by name (layers.py:80-106 AGGREGATORS)
        for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
            for m in cls.body:
                if isinstance(m, ast.FunctionDef) and m.name == "__init__":
                    for st in ast.walk(m):
                        if (isinstance(st, ast.Assign) and len(st.targets) == 1
                                and isinstance(st.targets[0], ast.Attribute) and _dict_keys(st.value)):
                            entry["dicts"][f"{cls.name}.{st.targets[0].attr}"] = {"line": st.lineno,
                                                                                 "keys": _dict_keys(st.value)}
        api[rel] = entry
    return api


if __name__ == "__main__":
    api = collect()
    with open(OUT, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
        f.write("\n")
    n = sum(len(c["methods"]) for e in api.values() for c in e["classes"].values())
    print(f"wrote {os.path.normpath(OUT)}: {n} methods", file=sys.stderr)
