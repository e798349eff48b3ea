"""Differentiable BundleIteration / CameraIteration with the reference's own split of labour:

    reference:  TF graph ops (warp, resampler, Jacobians, damping, solve, update; TF autodiff)  +  native op
                `equation_construction` with its registered native gradient (bundlenet.py:76-82, 263)
    here:       the same graph in stock torch CUDA ops (torch autograd)                          +  native op
                `ops.equation_construction` = banet_eqc_fwd / banet_eqc_bwd (sm_100a)

This is the TRAINING path: it materialises J[nb,N,2,P], G[nb,N,C,2], d[nb,N,C,1] exactly like the reference does
(so it is meant for the reference's training regime, N <= a few thousand sampled points).  The fused kernels
(`ops.lm_build` ...) are the inference / no-grad path; a fused analytic backward is a later-round item.
Gradient signature = every float input: conv1, conv2, D, B, R, T, W and the lambda-MLP parameters.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops

Tensor = torch.Tensor
_SELU_ALPHA, _SELU_SCALE = 1.6732632423543772, 1.0507009873554805


def _resampler(data: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """bilinear, zero outside (tf.contrib.resampler semantics); differentiable w.r.t. data and coordinates."""
    nb, h, w, C = data.shape
    x0f, y0f = torch.floor(x), torch.floor(y)
    dx, dy = (x - x0f).unsqueeze(-1), (y - y0f).unsqueeze(-1)
    x0, y0 = x0f.long(), y0f.long()
    flat = data.reshape(nb, h * w, C)
    out = 0
    for xi, yi, wg in ((x0, y0, (1 - dx) * (1 - dy)), (x0 + 1, y0, dx * (1 - dy)), (x0, y0 + 1, (1 - dx) * dy), (x0 + 1, y0 + 1, dx * dy)):
        ok = ((xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)).unsqueeze(-1).to(data.dtype)
        idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).unsqueeze(-1).expand(-1, -1, C)
        out = out + torch.gather(flat, 1, idx) * (wg * ok)
    return out


def _rodrigues(w: Tensor) -> Tensor:
    """bundlenet.py:17-37 (theta clamped to 1e-6).  w [nb,3] -> [nb,3,3]"""
    th = torch.sqrt((w * w).sum(1)).clamp_min(1e-6)
    k = w / th.unsqueeze(1)
    c, s = torch.cos(th), torch.sin(th)
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    oc = 1 - c
    rows = [c + kx * kx * oc, kx * ky * oc - kz * s, ky * s + kx * kz * oc,
            kz * s + kx * ky * oc, c + ky * ky * oc, -kx * s + ky * kz * oc,
            -ky * s + kx * kz * oc, kx * s + ky * kz * oc, c + kz * kz * oc]
    return torch.stack(rows, 1).reshape(-1, 3, 3)


def _vmatrix(w: Tensor) -> Tensor:
    """bundlenet.py:39-46, per pair (series below 1e-4 like the CUDA path)."""
    th2 = (w * w).sum(1)
    th = torch.sqrt(th2.clamp_min(1e-30))
    small = th < 1e-4
    ths = torch.where(small, torch.ones_like(th), th)
    ca = torch.where(small, 0.5 - th2 / 24, (1 - torch.cos(ths)) / (ths * ths))
    cb = torch.where(small, 1.0 / 6 - th2 / 120, (ths - torch.sin(ths)) / (ths * ths * ths))
    z = torch.zeros_like(th)
    K = torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], 1).reshape(-1, 3, 3)
    eye = torch.eye(3, device=w.device, dtype=w.dtype).unsqueeze(0)
    return eye + ca.view(-1, 1, 1) * K + cb.view(-1, 1, 1) * (K @ K)


def lambda_mlp(avg_residual: Tensor, params: Sequence[Tuple[Tensor, Tensor]]) -> Tensor:
    h = avg_residual
    for i, (Wt, b) in enumerate(params):
        h = h @ Wt + b
        h = torch.tanh(h) if i == len(params) - 1 else _SELU_SCALE * torch.where(h > 0, h, _SELU_ALPHA * torch.expm1(h))
    return h


def iteration(conv1, conv2, intr, p, D, B, R, T, W, mlp_params, l2_regularizer_base: Optional[float],
              damping_eps: float = 1e-5, exact_sym: bool = False, lambda_override: Optional[Tensor] = None):
    """One differentiable LM iteration.  B/W None -> CameraIteration (bundlenet.py:122-191), else BundleIteration (:193-278).
    intr [nb,4].  Returns (R', T', W')."""
    nb, N, C = conv1.shape
    h, w = conv2.shape[1], conv2.shape[2]
    fx, fy, ox, oy = [intr[:, i:i + 1] for i in range(4)]
    bundle = B is not None
    Dt = (D + B @ W) if bundle else D                                             # :208
    Rp = R @ p                                                                    # :209
    X = Rp * Dt.transpose(1, 2) + T                                               # :211-214
    Z = X[:, 2]; x = X[:, 0] / Z; y = X[:, 1] / Z                                 # :216-221
    px, py = fx * x + ox, fy * y + oy                                             # :223-224
    ok = (px >= 0) & (px <= w - 1) & (py >= 0) & (py <= h - 1) & torch.isfinite(px) & torch.isfinite(py)   # :231 (+ finite guard)
    pxs, pys = torch.where(ok, px, torch.zeros_like(px)), torch.where(ok, py, torch.zeros_like(py))
    s = _resampler(conv2, pxs, pys)                                               # :230
    m = ok.to(conv1.dtype).unsqueeze(-1)
    diff = ((conv1 - s[..., :C]) * m).unsqueeze(-1)                               # :234,238
    grad = torch.stack([s[..., C:2 * C] * m, s[..., 2 * C:3 * C] * m], dim=-1)    # :235-239
    avg = diff.squeeze(-1).abs().mean(dim=1, keepdim=True)                        # :243
    if lambda_override is not None:
        lam = lambda_override.reshape(nb, 1, 1)
    else:
        lam = torch.pow(torch.linalg.norm(avg, dim=-1, keepdim=True), 2.0 + lambda_mlp(avg, mlp_params))   # :244-249
        if bundle and l2_regularizer_base is not None:
            lam = l2_regularizer_base * lam                                       # :252-253
    iZ = 1.0 / Z
    zeros = torch.zeros_like(x)
    Jx = -fx.unsqueeze(-1) * torch.stack([x * y, -1 - x * x, y, -iZ, zeros, x * iZ], dim=2)      # :49-61
    Jy = -fy.unsqueeze(-1) * torch.stack([1 + y * y, -x * y, -x, zeros, -iZ, y * iZ], dim=2)
    J = torch.stack([Jx, Jy], dim=2)                                              # [nb,N,2,6]
    if bundle:
        jd = torch.stack([fx * ((Rp[:, 0] - Rp[:, 2] * x) * iZ), fy * ((Rp[:, 1] - Rp[:, 2] * y) * iZ)], dim=2)   # :63-74
        J = torch.cat([J, jd.unsqueeze(-1) * B.unsqueeze(-2)], dim=-1)            # :260-261
    AtA, Atb = ops.equation_construction(J.contiguous(), grad.contiguous(), diff.contiguous(), exact_sym)   # :263 (native fwd + bwd)
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)
    if bundle:
        dvec = torch.cat([diag[:, :-1] + damping_eps, torch.zeros(nb, 1, device=diag.device, dtype=diag.dtype)], dim=-1)   # :266
    else:
        dvec = diag + damping_eps                                                 # :182
    sol = torch.linalg.solve(AtA + torch.diag_embed(dvec * lam.reshape(nb, 1)), Atb)               # :267 / :183
    wv, tv = sol[:, 0:3, 0], sol[:, 3:6, :]
    dr = _rodrigues(wv)
    Rn = dr @ R                                                                   # :274
    Tn = _vmatrix(wv) @ tv + dr @ T                                               # :275
    Wn = (W + sol[:, 6:, :]) if bundle else None                                  # :276
    return Rn, Tn, Wn
