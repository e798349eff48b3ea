"""GPU: the legacy tracker loop on the device (banet_lm_track_legacy: accept / reject + early termination without host sync) against the
oracle's `legacy_track` / `legacy_camera_iteration2`, which tests/test_oracle_pinned.py holds to the reference's own legacy/ba.py."""
import numpy as np
import pytest
import torch

from helpers import O, rel_fro, to_cuda32
import gen_golden as GG
import gen_ref_golden as GR

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _tracker(C, mlps):
    from banet_b200 import legacy
    tr = legacy.Tracker(C).cuda()
    for lv, params in mlps.items():
        for i, (w, b) in enumerate(params):
            getattr(tr, f"lambda_{lv}_{i + 1}_filters").data.copy_(w); getattr(tr, f"lambda_{lv}_{i + 1}_biases").data.copy_(b)
    return tr, legacy


@pytest.mark.parametrize("early", [True, False])
def test_tracker_loop_matches_oracle(early):
    x = GR.track_inputs()
    mlps = {str(l): GG.mlp_for(4, l) for l in (1, 2, 3)}
    iters = [3, 5, 7] if early else [2, 2, 2]
    tr, legacy = _tracker(4, mlps)
    legacy.early_termination = early
    try:
        out = tr.trackTF(to_cuda32(x["intr"]), [to_cuda32(l) for l in x["layers"]], to_cuda32(x["points"]), to_cuda32(x["d"]), to_cuda32(x["R0"]), to_cuda32(x["T0"]), iters)
    finally:
        legacy.early_termination = True
    ref = O.legacy_track(x["intr"], x["layers"], x["points"], x["d"], x["R0"], x["T0"], iters, mlps, early_termination=early)
    if early:
        R, T, ratio = out; oR, oT, oratio = ref
        assert int(tr.last_status.abs().max()) == 0
        print("iterations per level:", tr.last_iters_done.tolist(), " |T - oT|:", float((T.cpu().double() - oT).abs().max()))
        assert rel_fro(R, oR) < 1e-5 and float((T.cpu().double() - oT).abs().max()) < 2e-5 and abs(float(ratio[0]) - float(oratio)) < 1e-4
    else:
        Rs, Ts, ratio = out; oRs, oTs, oratio = ref
        assert len(Rs) == len(oRs) == 6
        for a, b in zip(Rs, oRs):
            assert rel_fro(a, b) < 1e-5
        for a, b in zip(Ts, oTs):
            assert float((a.cpu().double() - b).abs().max()) < 2e-5
        assert abs(float(ratio[0]) - float(oratio)) < 1e-5


def test_accept_reject_per_pair():
    """Two pairs in one batch, one far from / one near the optimum, residual_ratio chosen so that one step is kept and the other rejected
    (legacy/ba.py:343): per-pair decisions, iteration counts and poses against oracle.legacy_camera_iteration2 run pair by pair."""
    from banet_b200 import ops
    sc = GG._scene(2, 48, 64, 6, 0, (3,), 25, n_points=400)
    lv = sc.levels[0]
    a = GG._lv64(lv)
    mlp = GG.mlp_for(6, 3)
    R0 = sc.R0.to(F64).clone(); T0 = sc.T0.to(F64).clone()
    T0[1] = T0[1] + 0.05                                   # pair 1 starts far away
    level = ops.Level(to_cuda32(lv.conv1), to_cuda32(lv.conv2), to_cuda32(lv.intr), to_cuda32(lv.p), to_cuda32(lv.D), None)
    for ratio in (1.0, 0.7, 1e-3):
        R, T, done, vr, st = ops.lm_track_legacy([level], [1], to_cuda32(R0), to_cuda32(T0), [ops.pack_mlp(mlp).cuda()], True, 0.0, 0.0, ratio)
        for b in range(2):
            sl = lambda t: t[b:b + 1]
            oR, oT, uw, ut, nv = O.legacy_camera_iteration2(sl(a["conv1"]), sl(a["conv2"]), sl(a["fx"]), sl(a["fy"]), sl(a["ox"]), sl(a["oy"]), sl(a["p"]), sl(a["D"]),
                                                            sl(R0), sl(T0), mlp, residual_ratio=ratio)
            kept = float(uw) > 0
            print(f"ratio={ratio} pair {b}: kept={kept}")
            assert rel_fro(R[b:b + 1], oR) < 1e-5 and rel_fro(T[b:b + 1], oT) < 1e-4
            assert abs(float(vr[b]) - float(nv)) < 1e-4 and int(done[0, b]) == 1
