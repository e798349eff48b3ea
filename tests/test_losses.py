"""The mirror's training losses (banet_b200.bundlenet: stock torch, like the CNN around the layer) against the reference's own code
(tests/golden/ref_losses.npz, produced by executing bundlenet.py:6-15, 401-463 through the TF-1 shim)."""
import os

import numpy as np
import torch

from helpers import GOLDEN_DIR
import gen_ref_golden as GR


def test_mirror_losses_match_reference_code():
    from banet_b200.bundlenet import BundleNet, rotation2quaternion
    r = np.load(os.path.join(GOLDEN_DIR, "ref_losses.npz"))
    x = GR.loss_inputs()
    net = BundleNet(4, levels=())
    qp, qg = rotation2quaternion(x["Rp"]), rotation2quaternion(x["Rg"])
    assert np.allclose(qp.numpy(), r["out_qp"], rtol=1e-12, atol=1e-14)
    assert abs(float(net.lossR(qp, qg)) - float(r["out_lossR"][0])) < 1e-12
    assert abs(float(net.lossT(x["Tp"], x["Tg"])) - float(r["out_lossT"][0])) < 1e-12
    lf = net.lossF(x["intr"], x["depth"], x["mask"], x["Rp"], x["Tp"], x["Rg"], x["Tg"])
    assert abs(float(lf) - float(r["out_lossF"][0])) < 1e-10 * max(1.0, float(r["out_lossF"][0]))
    # differentiable w.r.t. the predictions
    Rp = x["Rp"].clone().requires_grad_(); Tp = x["Tp"].clone().requires_grad_()
    net.lossF(x["intr"], x["depth"], x["mask"], Rp, Tp, x["Rg"], x["Tg"]).backward()
    assert torch.isfinite(Rp.grad).all() and float(Tp.grad.abs().max()) > 0
