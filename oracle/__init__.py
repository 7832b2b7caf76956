"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the multi-mask aggregation hot path.

Nothing under mma_b200/ imports this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs do (as the checker or the
timed CPU baseline, never as a fallback).  See oracle/restate.py for the parity
status ("unpinned" by the reference's own tests; pinned against outputs of the
verbatim reference, tests/golden/).
"""
