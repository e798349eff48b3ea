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


def test_reference_op_timed_beside_the_b200_kernels():
    """The reference's own CUDA op (cuBLAS batched SGEMM chain + its serial column reduction, utils.cu:331-414 / :613-690) and the B200
    kernels behind the same op boundary, timed on the same GPU at the reference's own scale (nb=2, 4096 sampled points, legacy/seq_example.py:12)
    and at a 160x120 level.  Also the fused layer-level build (never materialises J, G, d) at the same shape.  Report only (written to
    gpurun_out/ref_op_timing.json when that directory exists); the single assertion is that the replacement is not slower."""
    import json, time
    from banet_b200 import ops, synth
    ref = _ref()
    report = []
    for nb, gh, gw in ((2, 64, 64), (4, 120, 160)):
        N, C, K = gh * gw, 128, 128; P = K + 6
        g = torch.Generator().manual_seed(7)
        J = torch.randn(nb, N, 2, P, generator=g).cuda(); G = torch.randn(nb, N, C, 2, generator=g).cuda(); d = torch.randn(nb, N, C, 1, generator=g).cuda()
        lg = torch.randn(nb, P, P, generator=g).cuda(); rg = torch.randn(nb, P, 1, generator=g).cuda()

        def wall(fn, reps=5):           # the reference harness synchronises the device itself: wall clock around synchronous calls
            fn(); torch.cuda.synchronize(); ts = []
            for _ in range(reps):
                t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            return sorted(ts)[len(ts) // 2] * 1e3
        row = {"nb": nb, "N": N, "C": C, "P": P,
               "reference_fwd_ms": wall(lambda: ref.equation_construction(J, G, d)),
               "reference_bwd_ms": wall(lambda: ref.equation_construction_grad(J, G, d, lg, rg)),
               "b200_eqc_fwd_ms": wall(lambda: ops.equation_construction(J, G, d)),
               "b200_eqc_bwd_ms": wall(lambda: ops.equation_construction_grad(J, G, d, lg, rg))}
        del J, G, d
        sc = synth.make_scene(nb=nb, H=gh, W=gw, C=C, K=K, level_ids=(3,), seed=11, device="cuda", dtype=torch.float32)
        lv = sc.levels[0]; L = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
        row["b200_fused_build_fp32_ms"] = wall(lambda: ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=0))
        row["b200_fused_build_auto_ms"] = wall(lambda: ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=-1))
        report.append(row)
        print(row)
        assert row["b200_eqc_fwd_ms"] < row["reference_fwd_ms"] and row["b200_eqc_bwd_ms"] < row["reference_bwd_ms"]
    out = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out")
    if os.path.isdir(out):
        json.dump(report, open(os.path.join(out, "ref_op_timing.json"), "w"), indent=1)
