"""CPU tests that pin the oracle against itself three ways (the reference ships no golden vectors —
PARITY UNPINNED, see oracle/__init__.py) and against the committed golden fixtures."""
import numpy as np
import pytest
import torch

from helpers import O, scene_case, oracle_level_inputs, mlp_for, rel_fro, GOLDEN_DIR
from oracle import gemm_chain


def test_gemm_chain_matches_equation_construction():
    rng = np.random.default_rng(0)
    N, C, P = 6, 5, 11
    J = rng.standard_normal((N, 2, P)); G = rng.standard_normal((N, C, 2)); d = rng.standard_normal((N, C, 1))
    left, right = gemm_chain.equation_construction_chain(J, G, d)
    tl, tr = O.equation_construction(torch.tensor(J)[None], torch.tensor(G)[None], torch.tensor(d)[None])
    assert np.abs(left - tl[0].numpy()).max() < 1e-12
    assert np.abs(right - tr[0].numpy()).max() < 1e-12


def test_gemm_chain_matches_equation_construction_grad():
    rng = np.random.default_rng(1)
    N, C, P = 4, 6, 9
    J = rng.standard_normal((N, 2, P)); G = rng.standard_normal((N, C, 2)); d = rng.standard_normal((N, C, 1))
    lg = rng.standard_normal((P, P)); rg = rng.standard_normal((P, 1))         # deliberately NON-symmetric
    a = gemm_chain.equation_construction_grad_chain(J, G, d, lg, rg)
    b = O.equation_construction_grad(*[torch.tensor(x)[None] for x in (J, G, d, lg, rg)])
    for x, y in zip(a, b):
        assert np.abs(x - y[0].numpy()).max() < 1e-12


def test_reference_grad_equals_autodiff_for_symmetric_upstream():
    torch.manual_seed(0)
    J = torch.randn(1, 5, 2, 7, dtype=torch.float64, requires_grad=True)
    G = torch.randn(1, 5, 4, 2, dtype=torch.float64, requires_grad=True)
    d = torch.randn(1, 5, 4, 1, dtype=torch.float64, requires_grad=True)
    S = torch.randn(1, 7, 7, dtype=torch.float64); S = S + S.transpose(1, 2)
    r = torch.randn(1, 7, 1, dtype=torch.float64)
    AtA, Atb = O.equation_construction(J, G, d)
    ((AtA * S).sum() + (Atb * r).sum()).backward()
    dJ, dG, dd = O.equation_construction_grad(J.detach(), G.detach(), d.detach(), S, r)
    assert rel_fro(dJ, J.grad) < 1e-12 and rel_fro(dG, G.grad) < 1e-12 and rel_fro(dd, d.grad) < 1e-12


@pytest.mark.parametrize("K", [0, 5])
def test_structured_equals_materialised(K):
    sc = scene_case(nb=2, C=6, K=K, level_ids=(3,), seed=5)
    lv = sc.levels[0]; a = oracle_level_inputs(lv)
    mlp = mlp_for(6, 3)
    if K:
        _, _, _, aux = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"],
                                          sc.R0, sc.T0, sc.W0 + 0.01, mlp, return_aux=True)
        H, g, rbar, nv = O.normal_equations_structured(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"],
                                                       a["D"], a["B"], sc.R0, sc.T0, sc.W0 + 0.01)
    else:
        _, _, aux = O.camera_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"],
                                       sc.R0, sc.T0, mlp, return_aux=True)
        H, g, rbar, nv = O.normal_equations_structured(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"],
                                                       a["D"], None, sc.R0, sc.T0, None)
    assert rel_fro(H, aux["AtA"]) < 1e-12 and rel_fro(g, aux["Atb"]) < 1e-12
    assert rel_fro(rbar, aux["rbar"]) < 1e-14 and torch.equal(nv, aux["nvalid"])


def test_zero_motion_depth_jacobian_vanishes():
    """At R=I, T=0 the depth Jacobian is identically 0 (bundlenet.py:69-70) — why scenes do not start there."""
    sc = scene_case(nb=1, C=4, K=3, level_ids=(3,), seed=2)
    a = oracle_level_inputs(sc.levels[0])
    H, g, _, _ = O.normal_equations_structured(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"],
                                               sc.R0, torch.zeros_like(sc.T0), sc.W0)
    scale = H[:, :6, :6].abs().max()
    assert H[:, 6:, :].abs().max() < 1e-12 * scale and g[:, 6:].abs().max() < 1e-12 * g[:, :6].abs().max()


def test_planted_solution_converges():
    sc = scene_case(nb=2, H=96, W=128, C=8, K=4, level_ids=(1, 2, 3), seed=3)
    opts = O.IterOptions(lambda_override=torch.full((2,), 1e-2, dtype=torch.float64))
    R, T, W = sc.R0, sc.T0, sc.W0
    for lv in sc.levels:
        a = oracle_level_inputs(lv)
        for _ in range(5):
            R, T, W = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"],
                                         R, T, W, None, opts)
    assert (R - sc.R_true).norm() < 1e-6 and (T - sc.T_true).norm() < 1e-6 and (W - sc.W_true).norm() < 1e-5


def test_vmatrix_scramble_is_identity_for_single_pair():
    w = torch.randn(1, 3, 1, 1, dtype=torch.float64) * 0.1
    a = O.v_matrix(w[:, 0], w[:, 1], w[:, 2], batch_scramble=False)
    b = O.v_matrix(w[:, 0], w[:, 1], w[:, 2], batch_scramble=True)
    assert torch.equal(a, b)
    w = torch.randn(3, 3, 1, 1, dtype=torch.float64) * 0.1
    a = O.v_matrix(w[:, 0], w[:, 1], w[:, 2], batch_scramble=False)
    b = O.v_matrix(w[:, 0], w[:, 1], w[:, 2], batch_scramble=True)
    assert not torch.allclose(a, b)


def test_resampler_matches_legacy_gather_sampler_in_bounds():
    torch.manual_seed(1)
    img = torch.randn(2, 9, 11, 3, dtype=torch.float64)
    x = torch.rand(2, 40, dtype=torch.float64) * 10.0; y = torch.rand(2, 40, dtype=torch.float64) * 8.0
    a = O.resampler(img, torch.stack([x, y], -1))
    b, m = O.interpolate2d(img, x, y)
    assert torch.all(m == 1) and rel_fro(a, b) < 1e-14


def test_golden_fixtures_match_oracle():
    """The committed fixtures (tests/golden/gen_golden.py) must still be what the oracle produces."""
    import gen_golden
    for name, builder in gen_golden.CASES.items():
        ref = np.load(f"{GOLDEN_DIR}/{name}.npz")
        out = builder()
        for k in ref.files:
            if k.startswith("out_"):
                assert np.allclose(out[k], ref[k], rtol=1e-10, atol=1e-12), (name, k)


@pytest.mark.parametrize("K", [0, 5])
def test_structured_chunked_solve_equals_materialised_solve(K):
    """The bounded-memory whole-solve used as the checker at BASELINE sizes is the same maths as the reference-faithful one."""
    sc = scene_case(nb=2, C=6, K=K, level_ids=(2, 3), seed=21)
    olv = []
    for lv in sc.levels:
        a = oracle_level_inputs(lv)
        olv.append(O.LevelInputs(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], mlp_for(6, lv.level)))
    W0 = None if K == 0 else sc.W0
    R1, T1, W1 = O.lm_solve(olv, 2, sc.R0, sc.T0, W0)
    R2, T2, W2 = O.lm_solve_structured(olv, 2, sc.R0, sc.T0, W0, chunk=500)
    assert rel_fro(R2, R1) < 1e-11 and rel_fro(T2, T1) < 1e-10
    if K:
        assert rel_fro(W2, W1) < 1e-9
