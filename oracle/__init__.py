"""CPU oracle for the APG reweight -> resample hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement (numpy / torch-CPU) of the reference algorithms
in hao-w/apgs for the path named in BASELINE.json `north_star`.  It exists so that
the CUDA product path in `apgs_b200/` can be checked for parity.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
    legs of `bench.py` may import anything from here;
  * nothing under `apgs_b200/` imports it, and `apgs_b200` has no CPU fallback.

Pinning: the reference ships NO tests, golden vectors or fixtures, so the
oracle is pinned against outputs of the *reference itself* run in the build container:
`tests/golden/make_golden.py` imports `/root/reference` (with two arithmetic-free import
stubs) under fixed seeds and writes the fixtures in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every oracle function against them.

Third-party arithmetic on the path: `torch.distributions` / ATen (torch 2.11.0 here; the
reference asks for "PyTorch >= 1.6", unpinned).  The published formulas are restated in
`oracle/dists.py`.
"""
