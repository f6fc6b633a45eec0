"""CPU oracle for the AO-ADMM hot path (TEST INFRASTRUCTURE - NOT PRODUCT CODE).

This package is a NumPy float64 restatement of the reference's algorithm
(`functions/cmtf_fun_AOADMM.m` Frobenius paths and the operators it calls).  It
exists only so that the CUDA engine can be checked against it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; nothing under `matlab-code_b200/` does.

PARITY UNPINNED: the reference is pure MATLAB with three un-vendored third
party packages (Tensor Toolbox v3.1, Proximity Operator Repository,
TV_Condat_v2) and ships no tests, golden vectors or expected outputs
(SURVEY.md section 8c).  Neither MATLAB nor Octave exists in the build
container, so the reference cannot be executed to pin this oracle.  The oracle
is instead validated by known-answer checks (tests/test_oracle_*.py): MTTKRP vs
einsum, the shortcut objective vs the explicit residual, every prox vs its
variational definition (brute force / KKT), noise-free recovery, and recovery of the
ground-truth factors that the reference ships for its own example_script11 dataset
(tests/golden/script11_tparafac2.npz) - the only fixture with a known answer.
"""
