"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (/root/reference/bundlenet.py, legacy/ba.py,
legacy/utils_python.py) on the seeded cases of gen_golden.py, through the torch-backed TF-1 API shim oracle/tf1_shim.py
(TensorFlow is not installable here; the shim's docstring lists exactly what is third-party restatement).  These fixtures pin the
oracle: tests/test_oracle_pinned.py holds every oracle function to what the reference code computed.

Run (in the build container, where /root/reference exists):  python tests/golden/gen_ref_golden.py
The only text edits applied to the reference before execution are the Python-2 -> 3 fixes listed in tf1_shim.PY2_FIXES.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import gen_golden as GG                      # noqa: E402  (same seeded input builders as the oracle's fixtures)
from oracle import tf1_shim as shim          # noqa: E402

REF = os.environ.get("BANET_REFERENCE", "/root/reference")
F64 = torch.float64
_np = GG._np


def load_modules():
    bn = shim.load_reference(REF, "bundlenet.py", "ref_bundlenet")
    up = shim.load_reference(REF, "legacy/utils_python.py", "utils_python")
    # legacy/ba.py imports `feat` (the CNN feature extractor, out of scope) and `utils_python`
    ba = shim.load_reference(REF, "legacy/ba.py", "ref_legacy_ba", extra_modules={"feat": types.ModuleType("feat"), "utils_python": up})
    return bn, ba, up


def set_mlp(C, levels):
    shim.VARIABLES.clear()
    for lv in levels:
        for i, (w, b) in enumerate(GG.mlp_for(C, lv)):
            shim.VARIABLES[f"lambda_{lv}_{i + 1}_filters"] = w.unsqueeze(0)       # TF conv1d filter [1,cin,cout] (bundlenet.py:105)
            shim.VARIABLES[f"lambda_{lv}_{i + 1}_biases"] = b


def case_primitives(bn, ba, up):
    g = torch.Generator().manual_seed(301)
    nb, N = 3, 17
    w = torch.randn(nb, 3, 1, generator=g, dtype=F64) * 0.3          # wx,wy,wz [nb,1,1] as tf.split of the solution gives them (bundlenet.py:269)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    x, y = torch.randn(nb, N, generator=g, dtype=F64), torch.randn(nb, N, generator=g, dtype=F64)
    Z = 1.0 + torch.rand(nb, N, generator=g, dtype=F64)
    fx = torch.full((nb, N), 31.0, dtype=F64); fy = torch.full((nb, N), 29.0, dtype=F64)
    r = torch.randn(nb, 3, N, generator=g, dtype=F64)
    img = torch.randn(nb, 7, 9, 4, generator=g, dtype=F64)
    pts = torch.rand(nb, N, 2, generator=g, dtype=F64) * torch.tensor([12.0, 10.0], dtype=F64) - 1.5      # some outside the 9x7 map
    net = bn.BundleNet()
    tr = ba.Tracker.__new__(ba.Tracker)          # __init__ builds the CNN + TF session (out of scope); the BA methods need no state
    R = GG.synth.rodrigues(torch.randn(nb, 3, generator=g) * 0.2).to(F64)
    ox = torch.full((nb, N), 4.0, dtype=F64); oy = torch.full((nb, N), 3.0, dtype=F64)
    s2d, m2d = up.interpolate2d(img, pts[..., 0], pts[..., 1])
    return dict(in_w=_np(w), in_x=_np(x), in_y=_np(y), in_Z=_np(Z), in_fx=_np(fx), in_fy=_np(fy), in_r=_np(r), in_img=_np(img), in_pts=_np(pts),
                in_R=_np(R), in_ox=_np(ox), in_oy=_np(oy),
                out_rotation=_np(bn.AngleaAxisRotation(wx, wy, wz)), out_vmatrix=_np(bn.VMatrix(wx, wy, wz)),
                out_vmatrix_nb1=_np(bn.VMatrix(wx[:1], wy[:1], wz[:1])),
                out_camera_jacobian=_np(bn.CameraJacobianMatrix(x, y, Z, fx, fy)),
                out_depth_jacobian=_np(bn.DepthJacobianMatrix(r[:, 0:1], r[:, 1:2], r[:, 2:3], x, y, Z, fx, fy)),
                out_grad_fixed=_np(net.grad_fixed(img)), out_coordinates=_np(net.computeCoordinates(pts, fx, fy, ox, oy)),
                out_resampler=_np(shim.tf.contrib.resampler.resampler(img, pts)),
                out_quaternion=_np(bn.rotation2quaternion(R)),
                out_legacy_camera_jacobian=_np(tr.CameraJacobianMatrix(x, y, Z, fx, fy)),
                out_legacy_coordinates=_np(tr.computeCoordinates(pts, fx, fy, ox, oy)),
                out_legacy_rotation=_np(tr.AngleaAxisRotation(wx, wy, wz)),
                out_interpolate2d=_np(s2d), out_interpolate2d_mask=_np(m2d))


def case_bundle_iteration(bn, nb):
    """BundleNet.BundleIteration (bundlenet.py:193-278) on gen_golden's bundle_iteration inputs (first nb pairs)."""
    sc = GG._scene(2, 48, 64, 8, 4, (3,), 21)
    a = GG._lv64(sc.levels[0])
    a = {k: (None if v is None else v[:nb]) for k, v in a.items()}
    set_mlp(8, ["3"])
    R, T, W = sc.R0.to(F64)[:nb], sc.T0.to(F64)[:nb], GG._f32(sc.W0.to(F64) + 0.01)[:nb]
    net = bn.BundleNet()
    Rn, Tn, Wn = net.BundleIteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], R, T, W, 1000.0, "3")
    return dict(in_checksum=GG._checksum(a["conv1"], a["conv2"], a["p"], a["D"], a["B"], R, T, W), out_R=_np(Rn), out_T=_np(Tn), out_W=_np(Wn))


def case_camera_iteration(bn, nb):
    sc = GG._scene(2, 48, 64, 6, 0, (3,), 22, n_points=300)
    a = GG._lv64(sc.levels[0])
    a = {k: (None if v is None else v[:nb]) for k, v in a.items()}
    set_mlp(6, ["3"])
    R, T = sc.R0.to(F64)[:nb], sc.T0.to(F64)[:nb]
    Rn, Tn = bn.BundleNet().CameraIteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], R, T, 1.0, "3")
    return dict(in_checksum=GG._checksum(a["conv1"], a["conv2"], a["p"], a["D"], R, T), out_R=_np(Rn), out_T=_np(Tn))


def case_resize(bn):
    """BundleNet.BundleResize (:332-399) and CameraResize (:280-329) on gen_golden's resize inputs."""
    x = GG.resize_inputs()
    set_mlp(4, ["0", "1", "2", "3"])
    net = bn.BundleNet()
    Rs, Ts, Ds = net.BundleResize(x["intr"], x["layers"], x["points"], x["basis"], x["depth"], x["R0"], x["T0"])
    rot, tr = net.CameraResize(x["intr"], x["layers"], x["points"], x["depth"])
    out = dict(in_checksum=GG._checksum(*x["layers"], x["basis"], x["depth"], x["points"]))
    for i in range(2):
        out[f"out_R{i}"] = _np(Rs[i]); out[f"out_T{i}"] = _np(Ts[i]); out[f"out_D{i}"] = _np(Ds[i])
    for i in range(4):
        out[f"out_camR{i}"] = _np(rot[i]); out[f"out_camT{i}"] = _np(tr[i])
    return out


def loss_inputs(seed=27, nb=2, h=12, w=16):
    g = torch.Generator().manual_seed(seed)
    Rp = GG.synth.rodrigues(torch.randn(nb, 3, generator=g) * 0.05).to(F64); Rg = GG.synth.rodrigues(torch.randn(nb, 3, generator=g) * 0.05).to(F64)
    Tp = torch.randn(nb, 3, generator=g, dtype=F64) * 0.1; Tg = torch.randn(nb, 3, generator=g, dtype=F64) * 0.1
    depth = 1.0 + 2.0 * torch.rand(nb, h, w, 1, generator=g, dtype=F64)
    mask = (torch.rand(nb, h, w, 1, generator=g) > 0.3).to(F64)
    intr = torch.tensor([[[20.0], [21.0], [8.0], [6.0]]], dtype=F64).repeat(nb, 1, 1)
    return dict(Rp=Rp, Rg=Rg, Tp=Tp, Tg=Tg, depth=depth, mask=mask, intr=intr)


def case_losses(bn):
    """rotation2quaternion (bundlenet.py:6-15), lossR / lossT / lossF (:401-463)."""
    x = loss_inputs()
    net = bn.BundleNet()
    qp, qg = bn.rotation2quaternion(x["Rp"]), bn.rotation2quaternion(x["Rg"])
    return dict(out_qp=_np(qp), out_qg=_np(qg), out_lossR=np.array([float(net.lossR(qp, qg))]), out_lossT=np.array([float(net.lossT(x["Tp"], x["Tg"]))]),
                out_lossF=np.array([float(net.lossF(x["intr"], x["depth"], x["mask"], x["Rp"], x["Tp"], x["Rg"], x["Tg"]))]))


def legacy_inputs():
    """One keyframe->frame pair for the legacy pose-only tracker (legacy/ba.py; nb = 1 there): sparse points, C=6."""
    sc = GG._scene(1, 48, 64, 6, 0, (3,), 25, n_points=400)
    a = GG._lv64(sc.levels[0])
    return sc, a


def case_legacy(ba):
    """Tracker.CameraIteration2 (legacy/ba.py:226-345: lambda-MLP step + accept/reject re-evaluation) from two starting poses, and
    Tracker.CameraIteration (:147-214)."""
    sc, a = legacy_inputs()
    set_mlp(6, ["3"])
    tr = ba.Tracker.__new__(ba.Tracker)
    out = dict(in_checksum=GG._checksum(a["conv1"], a["conv2"], a["p"], a["D"]))
    starts = {"near": (sc.R0.to(F64), sc.T0.to(F64)),
              "far": (GG.synth.rodrigues(torch.tensor([[0.02, -0.03, 0.01]])).to(F64) @ sc.R0.to(F64), sc.T0.to(F64) + 0.05)}
    for name, (R, T) in starts.items():
        for ratio in (1.0, 1e-3):                 # 1e-3: the residual must drop 1000x for the step to be kept -> rejected
            ba.residual_ratio = ratio
            Rn, Tn, uw, ut, nv = tr.CameraIteration2(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], R, T, "3")
            tag = f"{name}_ratio{ratio:g}"
            out[f"out_{tag}_R"] = _np(Rn); out[f"out_{tag}_T"] = _np(Tn)
            out[f"out_{tag}_update"] = np.array([float(uw), float(ut), float(nv)])
        ba.residual_ratio = 1.0
        out[f"in_{name}_R"] = _np(R); out[f"in_{name}_T"] = _np(T)
    R, T = starts["near"]
    Rn, Tn, valid = tr.CameraIteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], R, T)
    out["out_plain_R"] = _np(Rn); out["out_plain_T"] = _np(Tn); out["out_plain_valid"] = np.array([float(valid)])
    return out


def track_inputs(seed=26, C=4, N=300):
    """Keyframe -> frame tracking problem of legacy/seq_example.py in miniature: 3 pyramid levels of a 2-image batch
    (index 0 = keyframe, 1 = current frame, legacy/ba.py:112-113), sparse keyframe points with depth, planted relative pose."""
    sc = GG.synth.make_scene(nb=1, H=48, W=64, C=C, K=0, level_ids=(1, 2, 3), seed=seed, n_points=N, dtype=torch.float32, device="cpu")
    # one shared point set (finest-level pixel coordinates) and depth: rebuild conv1 as features of a keyframe image is not needed --
    # trackTF samples layers[level-1][0:1] itself, so make the keyframe map = the map whose samples are conv1: use F2 warped back is
    # overkill; instead plant the solution by sampling the SAME map at the keyframe points (identity motion is the optimum of each level)
    layers = []
    for lv in sc.levels:
        f2 = lv.conv2[..., :C].to(F64)
        layers.append(torch.cat([f2, f2], dim=0))                      # keyframe map == frame map: optimum at R = I, T = 0
    fin = sc.levels[-1]
    pts = fin.points.to(F64)
    intr = torch.tensor(GG.synth.TUM_INTRINSICS, dtype=F64).mul(64.0 / 640.0).reshape(1, 4, 1)
    d = fin.D.to(F64)
    R0 = GG.synth.rodrigues(torch.tensor([[0.004, -0.006, 0.003]])).to(F64)
    T0 = torch.tensor([[[0.004], [-0.003], [0.002]]], dtype=F64)
    return dict(intr=intr, layers=layers, points=pts, d=d, R0=R0, T0=T0)


def case_track(ba):
    """Tracker.trackTF (legacy/ba.py:96-145): levels 1..3, early-terminated while_loop of CameraIteration2, and the fixed-count variant."""
    x = track_inputs()
    set_mlp(4, ["1", "2", "3"])
    tr = ba.Tracker.__new__(ba.Tracker)
    out = dict(in_checksum=GG._checksum(*x["layers"], x["points"], x["d"]))
    ba.early_termination = True
    R, T, ratio = tr.trackTF(x["intr"], x["layers"], x["points"], x["d"], x["R0"], x["T0"], [3, 5, 7])
    out["out_early_R"] = _np(R); out["out_early_T"] = _np(T); out["out_early_ratio"] = np.array([float(ratio)])
    ba.early_termination = False
    Rs, Ts, ratio = tr.trackTF(x["intr"], x["layers"], x["points"], x["d"], x["R0"], x["T0"], [2, 2, 2])
    out["out_fixed_R"] = _np(torch.stack(Rs)); out["out_fixed_T"] = _np(torch.stack(Ts)); out["out_fixed_ratio"] = np.array([float(ratio)])
    ba.early_termination = True
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    bn, ba, up = load_modules()
    cases = {"ref_primitives": case_primitives(bn, ba, up),
             "ref_bundle_iteration_nb1": case_bundle_iteration(bn, 1), "ref_bundle_iteration_nb2": case_bundle_iteration(bn, 2),
             "ref_camera_iteration_nb1": case_camera_iteration(bn, 1), "ref_camera_iteration_nb2": case_camera_iteration(bn, 2),
             "ref_resize": case_resize(bn), "ref_legacy": case_legacy(ba), "ref_track": case_track(ba), "ref_losses": case_losses(bn)}
    for name, out in cases.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
