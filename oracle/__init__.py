"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's hot path (SparseCADGCN forward + DetectionLoss + backward) used as
the parity checker.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product package never does.

Parity status: the reference ships NO tests, golden vectors or fixtures for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself: `oracle/make_golden.py`
imports the UNMODIFIED reference files from /root/reference (through the stand-ins in
`oracle/shims/` for the third-party packages that cannot be installed here), runs them on seeded
inputs, checks `oracle/restatement.py` against them, and commits the vectors under `tests/golden/`.
"""
