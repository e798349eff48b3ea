"""Generates the golden fixtures tests/golden/*.npz from the CPU oracle (float64 arithmetic on
float32-representable inputs).  Run:  python tests/golden/gen_golden.py

PARITY UNPINNED: the reference cannot execute here (TF-1.x/py2 absent) and ships no vectors, so these
fixtures freeze the ORACLE's outputs, not the reference's.  They exist so that (a) the oracle cannot
drift silently and (b) the GPU box (which has no /root/reference) compares against committed numbers.
Inputs are regenerated from seeds by the same builders (torch CPU generator); an input checksum is
stored with every case.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

from banet_b200 import synth                 # noqa: E402
from oracle import ba_oracle as O            # noqa: E402

F64 = torch.float64


def _f32(t):
    """Round to float32 (what the CUDA path sees), keep as float64 for the oracle."""
    return t.to(torch.float32).to(F64)


def _np(t):
    return t.detach().cpu().numpy()


def _checksum(*ts):
    return np.array([float(sum(t.double().abs().sum() for t in ts if t is not None))])


def case_eqc():
    g = torch.Generator().manual_seed(101)
    nb, N, C, P = 2, 70, 12, 22
    J = _f32(torch.randn(nb, N, 2, P, generator=g, dtype=F64))
    G = _f32(torch.randn(nb, N, C, 2, generator=g, dtype=F64))
    d = _f32(torch.randn(nb, N, C, 1, generator=g, dtype=F64))
    lg = _f32(torch.randn(nb, P, P, generator=g, dtype=F64)); rg = _f32(torch.randn(nb, P, 1, generator=g, dtype=F64))
    AtA, Atb = O.equation_construction(J, G, d)
    dJ, dG, dd = O.equation_construction_grad(J, G, d, lg, rg)
    return dict(in_J=_np(J), in_G=_np(G), in_d=_np(d), in_left_grad=_np(lg), in_right_grad=_np(rg),
                out_AtA=_np(AtA), out_Atb=_np(Atb), out_dJ=_np(dJ), out_dG=_np(dG), out_dd=_np(dd))


def _scene(nb, H, W, C, K, level_ids, seed, n_points=None):
    sc = synth.make_scene(nb=nb, H=H, W=W, C=C, K=K, level_ids=level_ids, seed=seed, n_points=n_points,
                          dtype=torch.float32, device="cpu")
    return sc


def _lv64(lv):
    fx, fy, ox, oy = [t.to(F64) for t in lv.intr_tiled()]
    return dict(conv1=lv.conv1.to(F64), conv2=lv.conv2.to(F64), fx=fx, fy=fy, ox=ox, oy=oy, p=lv.p.to(F64), D=lv.D.to(F64),
                B=None if lv.B is None else lv.B.to(F64))


def mlp_for(C, level):
    return [(_f32(w), _f32(b)) for w, b in O.init_lambda_mlp(C, seed=100 + int(level))]


def case_bundle_iteration():
    """One BundleIteration (bundlenet.py:193-278) at level 3, nb=2, 48x64 dense, C=8, K=4, MLP lambda."""
    sc = _scene(2, 48, 64, 8, 4, (3,), 21)
    a = _lv64(sc.levels[0]); mlp = mlp_for(8, 3)
    R, T, W = sc.R0.to(F64), sc.T0.to(F64), _f32(sc.W0.to(F64) + 0.01)
    Rn, Tn, Wn, aux = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"],
                                         R, T, W, mlp, O.IterOptions(l2_regularizer_base=1000.0), return_aux=True)
    return dict(in_checksum=_checksum(a["conv1"], a["conv2"], a["p"], a["D"], a["B"], R, T, W),
                out_AtA=_np(aux["AtA"]), out_Atb=_np(aux["Atb"]), out_rbar=_np(aux["rbar"]), out_lam=_np(aux["lam"]),
                out_solution=_np(aux["solution"]), out_nvalid=_np(aux["nvalid"]), out_R=_np(Rn), out_T=_np(Tn), out_W=_np(Wn))


def case_camera_iteration():
    """One CameraIteration (bundlenet.py:122-191), sparse sub-pixel points (N=300), C=6."""
    sc = _scene(2, 48, 64, 6, 0, (3,), 22, n_points=300)
    a = _lv64(sc.levels[0]); mlp = mlp_for(6, 3)
    R, T = sc.R0.to(F64), sc.T0.to(F64)
    Rn, Tn, aux = O.camera_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], R, T, mlp,
                                     return_aux=True)
    return dict(in_checksum=_checksum(a["conv1"], a["conv2"], a["p"], a["D"], R, T),
                out_AtA=_np(aux["AtA"]), out_Atb=_np(aux["Atb"]), out_rbar=_np(aux["rbar"]), out_lam=_np(aux["lam"]),
                out_solution=_np(aux["solution"]), out_R=_np(Rn), out_T=_np(Tn))


def case_lm_solve():
    """BASELINE cfg1 shape in miniature: 2 levels x 3 iterations, fixed lambda, nb=2, C=8, K=16."""
    sc = _scene(2, 48, 64, 8, 16, (2, 3), 23)
    R, T, W = sc.R0.to(F64), sc.T0.to(F64), sc.W0.to(F64)
    opts = O.IterOptions(lambda_override=torch.full((2,), 0.05, dtype=F64))
    levels = []
    for lv in sc.levels:
        a = _lv64(lv)
        levels.append(O.LevelInputs(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], []))
    Rn, Tn, Wn = O.lm_solve(levels, 3, R, T, W, opts)
    return dict(in_checksum=_checksum(*[x for lv in sc.levels for x in (lv.conv1, lv.conv2, lv.B)]),
                out_R=_np(Rn), out_T=_np(Tn), out_W=_np(Wn),
                out_err=np.array([float((Rn - sc.R_true.to(F64)).norm()), float((Tn - sc.T_true.to(F64)).norm()),
                                  float((Wn - sc.W_true.to(F64)).norm())]))


def resize_inputs(seed=24, nb=2, C=4, K=3, N=500):
    """Inputs of BundleResize/CameraResize at the reference's hard-coded geometry (320x256 crop, bundlenet.py:338-357)."""
    g = torch.Generator().manual_seed(seed)
    Hf, Wf = 256, 320
    layers = []
    for level in range(4):
        s = 2 ** (3 - level)
        f = synth.gaussian_blur_nchw(torch.randn(nb, C, Hf // s, Wf // s, generator=g), 1.5)
        layers.append(_f32((f / f.flatten(2).std(dim=2).view(nb, C, 1, 1)).permute(0, 2, 3, 1).contiguous().to(F64)))
    basis = synth.gaussian_blur_nchw(torch.randn(nb, K, Hf // 2, Wf // 2, generator=g), 4.0)
    basis = _f32((basis * torch.rsqrt(basis.flatten(2).var(dim=2) + 1e-3).view(nb, K, 1, 1)).permute(0, 2, 3, 1).contiguous().to(F64))
    depth = _f32((1.0 + 2.0 * torch.rand(nb, Hf // 2, Wf // 2, 1, generator=g)).to(F64))
    points = _f32((torch.rand(nb, N, 2, generator=g) * torch.tensor([300.0, 220.0]) + 10.0).to(F64))
    intr = _f32(torch.tensor([[[280.0], [285.0], [160.0], [120.0]]], dtype=F64).repeat(nb, 1, 1))
    w = torch.randn(nb, 3, generator=g) * 0.01
    R0 = _f32(synth.rodrigues(w).to(F64)); T0 = _f32((torch.randn(nb, 3, 1, generator=g) * 0.02).to(F64))
    return dict(intr=intr, layers=layers, points=points, basis=basis, depth=depth, R0=R0, T0=T0)


def case_bundle_resize():
    x = resize_inputs()
    mlps = {str(l): mlp_for(4, l) for l in range(4)}
    Rs, Ts, Ds = O.bundle_resize(x["intr"], x["layers"], x["points"], x["basis"], x["depth"], mlps, x["R0"], x["T0"])
    rot, tr = O.camera_resize(x["intr"], x["layers"], x["points"], x["depth"], mlps)
    out = dict(in_checksum=_checksum(*x["layers"], x["basis"], x["depth"], x["points"]))
    for i in range(2):
        out[f"out_R{i}"] = _np(Rs[i]); out[f"out_T{i}"] = _np(Ts[i]); out[f"out_D{i}"] = _np(Ds[i])
    for i in range(4):
        out[f"out_camR{i}"] = _np(rot[i]); out[f"out_camT{i}"] = _np(tr[i])
    return out


CASES = {"eqc": case_eqc, "bundle_iteration": case_bundle_iteration, "camera_iteration": case_camera_iteration,
         "lm_solve": case_lm_solve, "bundle_resize": case_bundle_resize}

if __name__ == "__main__":
    for name, fn in CASES.items():
        out = fn()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
