"""GPU: the reference's OWN native op — EquationConstruction / EquationConstructionGrad compiled unmodified from utils.cu
(oracle/_ref, built by oracle/Makefile against the TensorFlow stand-in headers) — against (a) the float64 oracle and its literal
cuBLAS-chain replay and (b) the B200 kernels banet_eqc_fwd / banet_eqc_bwd on the same inputs.  This pins SURVEY §8 rows a8-a10."""
import os

import numpy as np
import pytest
import torch

from helpers import O, rel_fro, GOLDEN_DIR

pytestmark = pytest.mark.gpu
CASES = [(2, 70, 12, 22), (1, 33, 5, 6), (2, 257, 128, 134), (1, 64, 8, 38)]


def _inputs(nb, N, C, P):
    g = torch.Generator().manual_seed(nb * 1000 + N + P)
    mk = lambda *s: torch.randn(*s, generator=g)
    return mk(nb, N, 2, P), mk(nb, N, C, 2), mk(nb, N, C, 1), mk(nb, P, P), mk(nb, P, 1)


def _ref():
    from oracle import ref_lib
    assert ref_lib.available(), "oracle/_ref/libbanet_ref_eqc.so missing: run `make -C oracle` in the build container (it travels with the repo)"
    return ref_lib


@pytest.mark.parametrize("nb,N,C,P", CASES)
def test_reference_op_vs_oracle_and_b200_kernels(nb, N, C, P):
    from banet_b200 import ops
    ref = _ref()
    J, G, d, lg, rg = _inputs(nb, N, C, P)
    rA, rb = ref.equation_construction(J.cuda(), G.cuda(), d.cuda())
    oA, ob = O.equation_construction(J.double(), G.double(), d.double())
    # the reference sums N per-pixel fp32 matrices serially in fp32 (utils.cu:181-198): ~1e-6 per element
    assert rel_fro(rA, oA) < 2e-5 and rel_fro(rb, ob) < 2e-5
    A, b = ops.equation_construction(J.cuda(), G.cuda(), d.cuda())
    assert rel_fro(A, rA) < 2e-5 and rel_fro(b, rb) < 2e-5
    rJ, rG, rd = ref.equation_construction_grad(J.cuda(), G.cuda(), d.cuda(), lg.cuda(), rg.cuda())
    oJ, oG, od = O.equation_construction_grad(J.double(), G.double(), d.double(), lg.double(), rg.double())      # the 2*A*Ghat form, utils.cu:648
    assert rel_fro(rJ, oJ) < 2e-5 and rel_fro(rG, oG) < 2e-5 and rel_fro(rd, od) < 2e-5
    dJ, dG, dd = ops.equation_construction_grad(J.cuda(), G.cuda(), d.cuda(), lg.cuda(), rg.cuda(), exact_sym=False)
    assert rel_fro(dJ, rJ) < 2e-5 and rel_fro(dG, rG) < 2e-5 and rel_fro(dd, rd) < 2e-5
    print(f"nb={nb} N={N} C={C} P={P}: ref vs oracle AtA {rel_fro(rA, oA):.1e}; b200 vs ref AtA {rel_fro(A, rA):.1e} dJ {rel_fro(dJ, rJ):.1e}")


def test_reference_op_reproduces_committed_golden():
    """tests/golden/ref_eqc.npz was written by tests/golden/gen_ref_eqc_golden.py from this same compiled reference kernel; the CPU suite
    holds the oracle to it (tests/test_oracle_pinned_eqc.py).  cuBLAS may pick other kernels on another driver: tolerance, not bit equality."""
    path = os.path.join(GOLDEN_DIR, "ref_eqc.npz")
    if not os.path.exists(path):
        pytest.skip("ref_eqc.npz not generated yet")
    ref = _ref()
    z = np.load(path)
    J, G, d, lg, rg = [torch.tensor(z[k]).cuda() for k in ("in_J", "in_G", "in_d", "in_left_grad", "in_right_grad")]
    A, b = ref.equation_construction(J, G, d)
    dJ, dG, dd = ref.equation_construction_grad(J, G, d, lg, rg)
    for got, key in ((A, "out_AtA"), (b, "out_Atb"), (dJ, "out_dJ"), (dG, "out_dG"), (dd, "out_dd")):
        assert rel_fro(got, z[key]) < 1e-5, key
