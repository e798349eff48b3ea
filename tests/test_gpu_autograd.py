"""GPU: the differentiable BundleIteration / CameraIteration / BundleResize against float64 autograd through the CPU oracle: same
outputs, same gradients w.r.t. every float input.  Two training paths: "fused" (banet_lm_build_bwd / banet_lm_solve_update_bwd, nothing
per-pixel materialised) and "reference_split" (torch graph + native equation_construction fwd/bwd, the reference's own split)."""
import pytest
import torch

from helpers import O, scene_case, oracle_level_inputs, mlp_for, rel_fro, to_cuda32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", ["fused", "reference_split"])
@pytest.mark.parametrize("K,exact", [(6, False), (0, False), (6, True)])
def test_iteration_gradients_match_oracle_autograd(K, exact, path):
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
    net = BundleNet(C, levels=("3",), exact_sym_grad=exact, training_path=path, precision=_lib.PREC_FP32_SIMT, strict_status=True).cuda()
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


def test_fused_backward_at_reference_training_scale():
    """N = 4096 sampled points, K = 128, C = 64 (the reference's training regime, legacy/seq_example.py:12): the fused backward against
    float64 autograd of the oracle with the reference's op gradient.  The reference-split path would materialise J [nb,4096,2,134] here."""
    from banet_b200 import autograd as ag, _lib
    _lib.require_device()
    C, K, nb = 64, 128, 2
    sc = scene_case(nb=nb, H=120, W=160, C=C, K=K, level_ids=(3,), seed=77, n_points=4096, dtype=torch.float32)
    lv = sc.levels[0]
    a = oracle_level_inputs(lv)
    names = ["conv1", "conv2", "D", "B"]
    for n in names:
        a[n] = a[n].clone().requires_grad_()
    R = sc.R0.double().clone().requires_grad_(); T = sc.T0.double().clone().requires_grad_()
    W = (sc.W0.double() + 0.01 * torch.randn(sc.W0.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)).clone().requires_grad_()
    g = torch.Generator().manual_seed(6)
    cR, cT, cW = torch.randn(nb, 3, 3, generator=g, dtype=torch.float64), torch.randn(nb, 3, 1, generator=g, dtype=torch.float64), torch.randn(nb, K, 1, generator=g, dtype=torch.float64)
    lam = torch.tensor([0.3, 0.7], dtype=torch.float64)
    oR, oT, oW = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], R, T, W, [],
                                    O.IterOptions(guard_nonfinite=True, reference_op_grad=True, lambda_override=lam))
    ((oR * cR).sum() + (oT * cT).sum() + (oW * cW).sum()).backward()
    t = {n: to_cuda32(a[n].detach()).requires_grad_() for n in names}
    Rg = to_cuda32(R.detach()).requires_grad_(); Tg = to_cuda32(T.detach()).requires_grad_(); Wg = to_cuda32(W.detach()).requires_grad_()
    gR, gT, gW, status = ag.iteration_fused(t["conv1"], t["conv2"], to_cuda32(lv.intr), to_cuda32(lv.p), t["D"], t["B"], Rg, Tg, Wg, [], None,
                                            lambda_override=lam.float().cuda(), return_status=True)
    assert int(status.abs().max()) == 0
    assert rel_fro(gR, oR) < 1e-5 and rel_fro(gT, oT) < 1e-4 and rel_fro(gW, oW) < 1e-3
    ((gR * cR.float().cuda()).sum() + (gT * cT.float().cuda()).sum() + (gW * cW.float().cuda()).sum()).backward()
    for n in names:
        e = rel_fro(t[n].grad, a[n].grad)
        print(f"grad {n}: {e:.2e}")
        assert e < 2e-3, n
    for nm, x, y in (("R", Rg, R), ("T", Tg, T), ("W", Wg, W)):
        e = rel_fro(x.grad, y.grad)
        print(f"grad {nm}: {e:.2e}")
        assert e < 2e-3, nm


def test_prestep_backward_ops_match_autograd():
    from banet_b200 import autograd as ag, _lib
    _lib.require_device()
    g = torch.Generator().manual_seed(9)
    nb, h, w, C, N, K = 2, 9, 11, 5, 60, 7
    F = torch.randn(nb, h, w, C, generator=g, dtype=torch.float64)
    for swap in (False, True):
        Fo = F.clone().requires_grad_()
        Fs = torch.cat([Fo[nb // 2:], Fo[:nb // 2]], 0) if swap else Fo
        ref = torch.cat([Fs, O.grad_fixed(Fs)], -1)
        c = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        (ref * c).sum().backward()
        Fg = F.float().cuda().requires_grad_()
        (ag.grad_fixed_concat(Fg, swap) * c.float().cuda()).sum().backward()
        assert rel_fro(Fg.grad, Fo.grad) < 1e-5
    pts = torch.rand(nb, N, 2, generator=g, dtype=torch.float64) * torch.tensor([w + 2.0, h + 2.0]) - 1.0
    Fo = F.clone().requires_grad_()
    ref = O.resampler(Fo, pts * 0.5)
    c = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * c).sum().backward()
    Fg = F.float().cuda().requires_grad_()
    (ag.resample(Fg, pts.float().cuda(), 0.5) * c.float().cuda()).sum().backward()
    assert rel_fro(Fg.grad, Fo.grad) < 1e-5
    basis = torch.randn(nb, 50, K, generator=g, dtype=torch.float64); Wt = torch.randn(nb, K, 1, generator=g, dtype=torch.float64); d0 = torch.randn(nb, 50, generator=g, dtype=torch.float64)
    bo, wo, do = basis.clone().requires_grad_(), Wt.clone().requires_grad_(), d0.clone().requires_grad_()
    c = torch.randn(nb, 50, generator=g, dtype=torch.float64)
    ((do + (bo @ wo).squeeze(-1)) * c).sum().backward()
    bg, wg, dg = basis.float().cuda().requires_grad_(), Wt.float().cuda().requires_grad_(), d0.float().cuda().requires_grad_()
    (ag.depth_compose(dg, bg, wg) * c.float().cuda()).sum().backward()
    assert rel_fro(bg.grad, bo.grad) < 1e-5 and rel_fro(wg.grad, wo.grad) < 1e-5 and rel_fro(dg.grad, do.grad) < 1e-6


def test_bundle_resize_is_differentiable():
    """The reference's training entry point (bundlenet.py:332-399) through the fused kernels: gradients w.r.t. the feature pyramid, the
    basis and the lambda-MLP parameters against float64 autograd of oracle.bundle_resize (exact op gradient on both sides)."""
    import gen_golden as GG
    from banet_b200.bundlenet import BundleNet
    from banet_b200 import _lib
    _lib.require_device()
    x = GG.resize_inputs(seed=31, nb=2, C=4, K=3, N=300)
    mlps = {str(l): [(w.clone().requires_grad_(), b.clone().requires_grad_()) for w, b in GG.mlp_for(4, l)] for l in range(4)}
    layers = [l.clone().requires_grad_() for l in x["layers"]]
    basis = x["basis"].clone().requires_grad_()
    Rs, Ts, Ds = O.bundle_resize(x["intr"], layers, x["points"], basis, x["depth"], mlps, x["R0"], x["T0"], O.IterOptions(guard_nonfinite=True))
    g = torch.Generator().manual_seed(2)
    cR, cT, cD = torch.randn(2, 3, 3, generator=g, dtype=torch.float64), torch.randn(2, 3, 1, generator=g, dtype=torch.float64), torch.randn(Ds[1].shape, generator=g, dtype=torch.float64)
    ((Rs[1] * cR).sum() + (Ts[1] * cT).sum() + 1e-2 * (Ds[1] * cD).sum()).backward()
    net = BundleNet(4, exact_sym_grad=True, precision=_lib.PREC_FP32_SIMT, strict_status=True).cuda()
    for lv in range(4):
        for i, (w, b) in enumerate(mlps[str(lv)]):
            getattr(net, f"lambda_{lv}_{i + 1}_filters").data.copy_(w.detach()); getattr(net, f"lambda_{lv}_{i + 1}_biases").data.copy_(b.detach())
    glayers = [to_cuda32(l.detach()).requires_grad_() for l in layers]
    gbasis = to_cuda32(basis.detach()).requires_grad_()
    gRs, gTs, gDs = net.BundleResize(to_cuda32(x["intr"]), glayers, to_cuda32(x["points"]), gbasis, to_cuda32(x["depth"]), to_cuda32(x["R0"]), to_cuda32(x["T0"]))
    assert rel_fro(gRs[1], Rs[1]) < 1e-5 and rel_fro(gTs[1], Ts[1]) < 1e-3 and rel_fro(gDs[1], Ds[1]) < 1e-4
    ((gRs[1] * cR.float().cuda()).sum() + (gTs[1] * cT.float().cuda()).sum() + 1e-2 * (gDs[1] * cD.float().cuda()).sum()).backward()
    for lv in (2, 3):
        e = rel_fro(glayers[lv].grad, layers[lv].grad)
        print(f"grad layers[{lv}]: {e:.2e}")
        assert e < 5e-3
    assert glayers[0].grad is None or float(glayers[0].grad.abs().max()) == 0.0          # levels 0,1 are not used by BundleResize
    e = rel_fro(gbasis.grad, basis.grad)
    print(f"grad basis: {e:.2e}")
    assert e < 5e-3
    for lv in (2, 3):
        e = rel_fro(getattr(net, f"lambda_{lv}_1_filters").grad, mlps[str(lv)][0][0].grad)
        print(f"grad lambda_{lv}_1_filters: {e:.2e}")
        assert e < 5e-3
