"""Registers / spills of every kernel (nvcc -Xptxas -v per translation unit) and SASS mnemonic counts of the built
library, as a markdown report for profiles/ -- needs no GPU:  python scripts/ptxas_sass_report.py profiles/X.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mma_b200 import build as B  # noqa: E402


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"<.*", "", re.sub(r"^void ", "", n)).split("(")[0] for n in out]


def main(path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"] + ["-Xptxas", "-v"]
    rows = []
    for src in B.SOURCES:
        log = subprocess.run([nvcc] + flags + ["-c", os.path.join(B.CSRC, src), "-o", "/dev/null"], capture_output=True,
                             text=True).stderr
        cur = None
        per = collections.defaultdict(lambda: {"n": 0, "regs": [], "spill": 0, "stack": 0})
        for ln in log.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                cur = m.group(1)
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", ln)
            if m and cur:
                per[cur]["stack"], per[cur]["spill"] = int(m.group(1)), int(m.group(2))
            m = re.search(r"Used (\d+) registers", ln)
            if m and cur:
                per[cur]["regs"].append(int(m.group(1)))
        names = list(per)
        agg = collections.OrderedDict()
        for n, d in zip(demangle(names), names):
            a = agg.setdefault(n, {"n": 0, "regs": [], "spill": 0, "stack": 0})
            a["n"] += 1
            a["regs"] += per[d]["regs"]
            a["spill"] = max(a["spill"], per[d]["spill"])
            a["stack"] = max(a["stack"], per[d]["stack"])
        for n, a in agg.items():
            r = a["regs"]
            rows.append((src, n, a["n"], f"{min(r)}..{max(r)}" if min(r) != max(r) else str(r[0]), a["spill"], a["stack"]))
    sass = subprocess.run(["cuobjdump", "-sass", B.LIB], capture_output=True, text=True).stdout
    mn = collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            mn[m.group(1)] += 1
    keys = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "LDGSTS", "LDGDEPBAR", "SYNCS",
            "FFMA2", "ATOMG", "RED", "ATOMS", "LDL", "STL"]
    with open(path, "w") as f:
        f.write("ptxas resource usage and SASS evidence of the end-of-round-2 library (sm_100a, nvcc 12.9, -O3 -lineinfo)\n"
                "generated here (no GPU needed) by scripts/ptxas_sass_report.py: nvcc -Xptxas -v per translation unit; "
                "cuobjdump -sass of libmma_b200.so\n\n## registers / spills per kernel (template instances folded: min..max)\n\n"
                "| source | kernel | instances | registers | max spill stores (B) | max stack (B) |\n|---|---|---:|---:|---:|---:|\n")
        for r in rows:
            f.write(f"| {r[0]} | `{r[1]}` | {r[2]} | {r[3]} | {r[4]} | {r[5]} |\n")
        f.write("\n## SASS mnemonics (whole library)\n\n| mnemonic | count | meaning |\n|---|---:|---|\n")
        what = {"UTCHMMA": "tcgen05.mma", "UTCBAR": "tcgen05.commit -> mbarrier", "LDTM": "tcgen05.ld (TMEM -> registers)",
                "STTM": "tcgen05.st (registers -> TMEM)", "UTMALDG": "TMA tile load (cp.async.bulk.tensor)",
                "UTMASTG": "TMA tile store", "UTMAPF": "TMA descriptor prefetch", "LDGSTS": "cp.async (global -> shared)",
                "LDGDEPBAR": "cp.async.commit_group", "SYNCS": "mbarrier ops", "ATOMG": "global atomics", "RED": "global reductions",
                "ATOMS": "shared atomics", "LDL": "local loads (spills / indexed arrays)", "STL": "local stores",
                "FFMA2": "packed fp32 FMA", "UTCATOMSWS": "tcgen05.alloc / dealloc"}
        for k in keys:
            f.write(f"| {k} | {mn.get(k, 0)} | {what.get(k, '')} |\n")
        f.write(f"\ntotal SASS instructions: {sum(mn.values())}\n")
    print(path, len(rows), "kernels")


if __name__ == "__main__":
    main(sys.argv[1])
