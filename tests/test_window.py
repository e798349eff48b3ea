"""The joint keyframe-window solve: an EXTENSION, the reference's BA layer is 2-view.  The oracle's statement of it
(oracle.window_iteration) is checked against first principles on the CPU — the materialised joint Jacobian, the 2-view iteration it must
reduce to, convergence on a planted window — and the CUDA path (banet_lm_window_run) against that oracle on the GPU."""
import pytest
import torch

from helpers import O, scene_case, oracle_level_inputs, mlp_for, rel_fro, to_cuda32


def _window_case(nf=3, C=8, K=5, seed=31, level_ids=(2, 3), **kw):
    sc = scene_case(nb=nf, C=C, K=K, seed=seed, level_ids=level_ids, shared_depth=True, **kw)
    lvs = []
    for l in sc.levels:
        a = oracle_level_inputs(l)
        lvs.append(O.LevelInputs(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], mlp_for(C, l.level)))
    return sc, lvs


def test_assembled_system_equals_the_materialised_joint_jacobian():
    sc, lvs = _window_case()
    lv = lvs[-1]
    nf, N, C = lv.conv1.shape
    K = lv.B.shape[-1]
    W = sc.W_true[0] * 0.5
    H, g, _, _ = O.normal_equations_structured(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B, sc.R0, sc.T0, W.expand(nf, K, 1))
    Hj, gj = O.window_assemble(H, g)
    # first principles: stack the residuals of all frames; frame f's rows of the joint Jacobian are [0 .. Jc_f .. 0 | Jd_f]
    Pj = 6 * nf + K
    Href = torch.zeros(Pj, Pj, dtype=torch.float64); gref = torch.zeros(Pj, 1, dtype=torch.float64)
    for f in range(nf):
        s = slice(f, f + 1)
        _, _, _, aux = O.bundle_iteration(lv.conv1[s], lv.conv2[s], lv.fx[s], lv.fy[s], lv.ox[s], lv.oy[s], lv.p[s], lv.D[s], lv.B[s], sc.R0[s], sc.T0[s],
                                          W.reshape(1, K, 1), lv.mlp, return_aux=True)
        A = aux["grad"][0] @ aux["J"][0]                                 # [N,C,P]: residual rows of frame f
        Aj = torch.zeros(N, C, Pj, dtype=torch.float64)
        Aj[..., 6 * f:6 * f + 6] = A[..., :6]; Aj[..., 6 * nf:] = A[..., 6:]
        Href += torch.einsum("ncp,ncq->pq", Aj, Aj); gref += torch.einsum("ncp,nc->p", Aj, aux["diff"][0, :, :, 0]).unsqueeze(-1)
    assert rel_fro(Hj, Href) < 1e-12 and rel_fro(gj, gref) < 1e-12


def test_a_window_of_one_frame_is_the_two_view_iteration():
    sc, lvs = _window_case(nf=1)
    lv = lvs[-1]
    a = O.window_iteration(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B, sc.R0, sc.T0, sc.W0[0], lv.mlp)
    b = O.iteration_structured(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B, sc.R0, sc.T0, sc.W0, lv.mlp)
    assert rel_fro(a[0], b[0]) < 1e-13 and rel_fro(a[1], b[1]) < 1e-13 and rel_fro(a[2], b[2][0]) < 1e-12


def test_planted_window_converges_to_the_shared_depth_and_every_pose():
    sc, lvs = _window_case(nf=4, C=8, K=5, seed=33)
    opts = O.IterOptions(lambda_override=torch.tensor([1e-3]))
    e0 = rel_fro(sc.W0[0], sc.W_true[0] + 1e-30) if float(sc.W_true.abs().max()) > 0 else 0
    R, T, W = O.window_solve(lvs, 8, sc.R0, sc.T0, sc.W0[0], opts)
    assert rel_fro(W, sc.W_true[0]) < 1e-3 < e0
    assert rel_fro(R, sc.R_true) < 1e-5 and rel_fro(T, sc.T_true) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("nf,C,K,prec", [(3, 16, 16, 0), (4, 128, 128, 0), (4, 128, 128, -1), (1, 128, 128, -1)])
def test_window_run_matches_oracle(nf, C, K, prec):
    from banet_b200 import ops
    sc, lvs = _window_case(nf=nf, C=C, K=K, seed=41, H=96, W=128)
    oR, oT, oW = O.window_solve(lvs, 3, sc.R0, sc.T0, sc.W0[0])
    levels = [ops.Level(to_cuda32(l.conv1), to_cuda32(l.conv2), to_cuda32(l.intr), to_cuda32(l.p), to_cuda32(l.D), to_cuda32(l.B), grid=l.grid) for l in sc.levels]
    packed = [ops.pack_mlp([(w.float(), b.float()) for w, b in lv.mlp]).cuda() for lv in lvs]
    R, T, W, st = ops.lm_window_run(levels, 3, to_cuda32(sc.R0), to_cuda32(sc.T0), to_cuda32(sc.W0[0]), mlp_packed=packed, l2_regularizer_base=1000.0, precision=prec)
    assert int(st.abs().max()) == 0
    e = (rel_fro(R, oR), rel_fro(T, oT), rel_fro(W, oW))
    print(f"window nf={nf} C={C} K={K} prec={prec}: R {e[0]:.1e} T {e[1]:.1e} W {e[2]:.1e}")
    assert e[0] < 1e-5 and e[1] < 1e-4 and e[2] < (2e-4 if prec == 0 else 1e-3)
    if nf == 1:                                  # one frame: the assembled system IS the pair's system -> bit-identical to banet_lm_run
        R2, T2, W2, _ = ops.lm_run(levels, 3, to_cuda32(sc.R0), to_cuda32(sc.T0), to_cuda32(sc.W0), mlp_packed=packed, l2_regularizer_base=1000.0, precision=prec)
        assert torch.equal(R, R2) and torch.equal(T, T2) and torch.equal(W, W2[0])
