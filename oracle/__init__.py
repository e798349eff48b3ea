"""CPU oracle for the BA-layer LM hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, line by line, the maths of frobelbest/BANet's BA layer
(`bundlenet.py`, `utils.cu`, `legacy/ba.py`, `legacy/utils_python.py`) in plain
torch-CPU (float64 by default, float32 for the timed CPU baseline).  It is the
checker that the CUDA path in `banet_b200/` is compared against.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import it;
  * nothing under `banet_b200/` may import, call or fall back to it.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and
cannot execute in this environment (TensorFlow-1.x / Python-2, absent; the
custom op needs TF headers).  The oracle is therefore normative; it is
cross-checked three ways (tests/test_oracle_*.py):
  1. the materialised reference form (J,G,d tensors + einsum) against a
     structured block form that never forms J;
  2. `equation_construction` / `_grad` against a literal emulation of the
     column-major cuBLAS GEMM chain of utils.cu:331-414 / :625-690;
  3. planted-solution synthetic scenes must converge.
"""
