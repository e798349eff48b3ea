"""GPU: the differentiable BundleIteration / CameraIteration (torch graph + native equation_construction fwd/bwd)
against float64 autograd through the CPU oracle: same outputs, same gradients w.r.t. every float input."""
import pytest
import torch

from helpers import O, scene_case, oracle_level_inputs, mlp_for, rel_fro, to_cuda32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,exact", [(6, False), (0, False), (6, True)])
def test_iteration_gradients_match_oracle_autograd(K, exact):
    """exact=False: the reference's registered op gradient (2*A*Ghat, utils.cu:648) on both sides;
    exact=True: true autodiff on the oracle side, exact_sym on ours."""
    from banet_b200.bundlenet import BundleNet
    from banet_b200 import _lib
    _lib.require_device()
    C = 8
    sc = scene_case(nb=2, C=C, K=K, level_ids=(3,), seed=61, n_points=400, dtype=torch.float32)
    lv = sc.levels[0]
    mlp = mlp_for(C, 3)
    # ---------------- oracle (float64, exact autodiff incl. the symmetric-exact op gradient)
    a = oracle_level_inputs(lv)
    names = ["conv1", "conv2", "D"] + (["B"] if K else [])
    for n in names:
        a[n] = a[n].clone().requires_grad_()
    R = sc.R0.double().clone().requires_grad_(); T = sc.T0.double().clone().requires_grad_()
    W = (sc.W0.double() + 0.01).clone().requires_grad_() if K else None
    mlp64 = [(w.clone().requires_grad_(), b.clone().requires_grad_()) for w, b in mlp]
    g = torch.Generator().manual_seed(5)
    cR, cT = torch.randn(2, 3, 3, generator=g, dtype=torch.float64), torch.randn(2, 3, 1, generator=g, dtype=torch.float64)
    cW = torch.randn(2, K, 1, generator=g, dtype=torch.float64) if K else None
    if K:
        oR, oT, oW = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], R, T, W, mlp64,
                                        O.IterOptions(l2_regularizer_base=1000.0, guard_nonfinite=True, reference_op_grad=not exact))
        loss = (oR * cR).sum() + (oT * cT).sum() + (oW * cW).sum()
    else:
        oR, oT = O.camera_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], R, T, mlp64,
                                    O.IterOptions(guard_nonfinite=True, reference_op_grad=not exact))
        loss = (oR * cR).sum() + (oT * cT).sum()
    loss.backward()
    # ---------------- ours (float32, CUDA)
    net = BundleNet(C, levels=("3",), exact_sym_grad=exact).cuda()
    for i, (w, b) in enumerate(mlp):
        getattr(net, f"lambda_3_{i + 1}_filters").data.copy_(w); getattr(net, f"lambda_3_{i + 1}_biases").data.copy_(b)
    t = {n: to_cuda32(a[n].detach()).requires_grad_() for n in names}
    Rg = to_cuda32(sc.R0).requires_grad_(); Tg = to_cuda32(sc.T0).requires_grad_()
    Wg = to_cuda32(sc.W0 + 0.01).requires_grad_() if K else None
    fx, fy, ox, oy = [to_cuda32(x) for x in lv.intr_tiled()]
    if K:
        gR, gT, gW = net.BundleIteration(t["conv1"], t["conv2"], fx, fy, ox, oy, to_cuda32(lv.p), t["D"], t["B"], Rg, Tg, Wg, 1000.0, "3")
        lossg = (gR * cR.float().cuda()).sum() + (gT * cT.float().cuda()).sum() + (gW * cW.float().cuda()).sum()
        assert rel_fro(gW, oW) < 1e-4
    else:
        gR, gT = net.CameraIteration(t["conv1"], t["conv2"], fx, fy, ox, oy, to_cuda32(lv.p), t["D"], Rg, Tg, 1.0, "3")
        lossg = (gR * cR.float().cuda()).sum() + (gT * cT.float().cuda()).sum()
    assert rel_fro(gR, oR) < 1e-5 and rel_fro(gT, oT) < 1e-4
    lossg.backward()
    tol = 2e-3
    for n in names:
        e = rel_fro(t[n].grad, a[n].grad)
        print(f"grad {n}: {e:.2e}")
        assert e < tol, n
    assert rel_fro(Rg.grad, R.grad) < tol and rel_fro(Tg.grad, T.grad) < tol
    if K:
        assert rel_fro(Wg.grad, W.grad) < tol
    for i, (w64, b64) in enumerate(mlp64):
        assert rel_fro(getattr(net, f"lambda_3_{i + 1}_filters").grad, w64.grad) < tol
        assert rel_fro(getattr(net, f"lambda_3_{i + 1}_biases").grad, b64.grad) < 5 * tol
