"""CPU: the oracle's `equation_construction(_grad)` and the literal cuBLAS-chain replay (oracle/gemm_chain.py) against outputs of
the reference's own compiled kernel (tests/golden/ref_eqc.npz, generated on a B200 by tests/golden/gen_ref_eqc_golden.py from
utils.cu built unmodified, oracle/Makefile).  fp32 cuBLAS + a serial fp32 column sum vs float64: tolerance 2e-5."""
import os

import numpy as np
import pytest
import torch

from helpers import O, rel_fro, GOLDEN_DIR
from oracle import gemm_chain

PATH = os.path.join(GOLDEN_DIR, "ref_eqc.npz")


@pytest.mark.skipif(not os.path.exists(PATH), reason="ref_eqc.npz not generated yet (needs one GPU session)")
def test_oracle_equals_compiled_reference_kernel():
    z = np.load(PATH)
    J, G, d, lg, rg = [torch.tensor(z[k], dtype=torch.float64) for k in ("in_J", "in_G", "in_d", "in_left_grad", "in_right_grad")]
    A, b = O.equation_construction(J, G, d)
    dJ, dG, dd = O.equation_construction_grad(J, G, d, lg, rg)
    for got, key in ((A, "out_AtA"), (b, "out_Atb"), (dJ, "out_dJ"), (dG, "out_dG"), (dd, "out_dd")):
        assert rel_fro(got, z[key]) < 2e-5, key
    for bi in range(J.shape[0]):
        l, r = gemm_chain.equation_construction_chain(z["in_J"][bi].astype(np.float64), z["in_G"][bi].astype(np.float64), z["in_d"][bi].astype(np.float64))
        assert rel_fro(l, z["out_AtA"][bi]) < 2e-5 and rel_fro(r, z["out_Atb"][bi]) < 2e-5
