"""What the built library contains, read from its SASS (no GPU needed): the dense projections are tcgen05 / TMA code,
their single-thread roles sit behind elect.sync, and no aggregation / gradient / GEMM kernel uses global atomics."""
import os
import re
import shutil
import subprocess

import pytest

from mma_b200 import _lib, build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
GEMM_KERNELS = {
    "_ZN3mma4gemm14gemm_nt_kernelENS0_8NtParamsE": "streaming 3xTF32 GEMM",
    "_ZN3mma4gemm19gemm_nt_ares_kernelENS0_8NtParamsE": "short-K GEMM, activation tile resident in tensor memory",
    "_ZN3mma4gemm17gemm_wgrad_kernelENS0_8WgParamsE": "weight-gradient GEMM",
}


def _sass(fun):
    if not os.path.isfile(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    build.build()
    out = subprocess.run([CUOBJDUMP, "-sass", "-fun", fun, _lib.LIB_PATH], capture_output=True, text=True).stdout
    ops = []
    for ln in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops.append(m.group(1))
    assert ops, f"{fun} not found in {_lib.LIB_PATH}"
    return ops


@pytest.mark.parametrize("fun", sorted(GEMM_KERNELS))
def test_gemm_kernels_are_tcgen05_tma_code_with_elected_issuers(fun):
    ops = _sass(fun)
    n_mma = sum(o == "UTCHMMA" for o in ops)
    assert n_mma >= 8, (GEMM_KERNELS[fun], n_mma)                       # tcgen05.mma
    assert any(o.startswith("UTMALDG") for o in ops)                      # TMA tile loads
    assert any(o.startswith("UTCBAR") for o in ops)                       # tcgen05.commit -> mbarrier
    assert any(o.startswith("LDTM") for o in ops)                         # tcgen05.ld in the epilogue
    # Behind `lane == 0` ptxas wraps EVERY UTCHMMA / UTMALDG in an ELECT ... BRA.U.ANY loop (it cannot prove that a
    # single thread is active): 5 instructions and a branch per MMA on the issuing thread, 4-6 % of every GEMM
    # (profiles/r2K_gemm_elect_ab.log).  Behind elect.sync (tc_common.cuh: elect_one) the loops are gone.
    n_loops = sum(o.startswith("BRA.U.ANY") for o in ops)
    assert n_loops <= 2, f"{GEMM_KERNELS[fun]}: {n_loops} ELECT/BRA.U.ANY loops around {n_mma} UTCHMMA"
    assert not any(o.split(".")[0] in ("ATOMG", "RED", "ATOM") for o in ops)   # deterministic: no global atomics


def test_streaming_and_wgrad_kernels_carry_the_n256_instruction():
    # both kernels hold two issue sequences (N = 256 merge on / off, MMA_GEMM_N256 / MMA_GEMM_WG_N256): the merged one
    # issues 2 MMAs per K step instead of 3, so a kernel holds 4 * (2 + 3) = 20 UTCHMMA (the streaming kernel 4 more for
    # its separate plain-TF32 sequence)
    for fun, want in (("_ZN3mma4gemm14gemm_nt_kernelENS0_8NtParamsE", 24), ("_ZN3mma4gemm17gemm_wgrad_kernelENS0_8WgParamsE", 20)):
        n_mma = sum(o == "UTCHMMA" for o in _sass(fun))
        assert n_mma == want, (fun, n_mma)
