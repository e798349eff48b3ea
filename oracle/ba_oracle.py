"""CPU restatement of the reference BA layer (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Tensor layouts are the reference's:

    conv1 [nb,N,C]      source-frame features sampled at the points
    conv2 [nb,h,w,3C]   target-frame map  [F2 | gradx | grady]   (NHWC)
    fx,fy,ox,oy [nb,N]  level-scaled intrinsics (constant along N)
    p  [nb,3,N]         unit rays            D [nb,N,1]   depth (range along the ray)
    B  [nb,N,K]         depth basis          W [nb,K,1]   basis coefficients
    R  [nb,3,3]  T [nb,3,1]

Everything is torch-CPU and differentiable, so the same code is also the
gradient oracle (float64).  Pinned to the reference's own code: see oracle/__init__.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


# --------------------------------------------------------------------------- SE(3) helpers
def angle_axis_rotation(wx: Tensor, wy: Tensor, wz: Tensor, clamp: bool = True) -> Tensor:
    """bundlenet.py:17-37 `AngleaAxisRotation`.  Inputs [nb,1,1] -> [nb,3,3].

    theta is clamped to >= 1e-6 (:20); the nine entries are stacked on the last
    axis, reshaped to [-1,3,3] and TRANSPOSED (:37), which yields the standard
    Rodrigues matrix  I + sin(t)[k]x + (1-cos t)[k]x^2.
    legacy/ba.py:60-80 is the same without the clamp (`clamp=False`).
    """
    ones = torch.ones_like(wx)
    theta = torch.sqrt(wx * wx + wy * wy + wz * wz)
    if clamp:
        theta = torch.clamp(theta, min=1e-6)
    wx, wy, wz = wx / theta, wy / theta, wz / theta
    c, s = torch.cos(theta), torch.sin(theta)
    e = torch.stack([c + wx * wx * (ones - c),
                     wz * s + wx * wy * (ones - c),
                     -wy * s + wx * wz * (ones - c),
                     wx * wy * (ones - c) - wz * s,
                     c + wy * wy * (ones - c),
                     wx * s + wy * wz * (ones - c),
                     wy * s + wx * wz * (ones - c),
                     -wx * s + wy * wz * (ones - c),
                     c + wz * wz * (ones - c)], dim=-1)
    return e.reshape(-1, 3, 3).transpose(1, 2)


def v_matrix(wx: Tensor, wy: Tensor, wz: Tensor, batch_scramble: bool = False) -> Tensor:
    """bundlenet.py:39-46 `VMatrix` (legacy twin: legacy/ba.py:51-58).  [nb,1,1] -> [nb,3,3].

    V = I + (1-cos t)/t^2 [w]x + (t - sin t)/t^3 [w]x^2  with UNCLAMPED t (0/0 at w = 0).

    Reference quirk (:45): the nine skew entries are `tf.stack`ed on axis 0 and then
    reshaped to [-1,3,3]; for nb > 1 that interleaves batch entries.  `batch_scramble=True`
    reproduces that literally; the default is the per-pair matrix, which is what the
    reference computes for nb == 1.
    """
    theta = torch.sqrt(wx * wx + wy * wy + wz * wz)
    c, s = torch.cos(theta), torch.sin(theta)
    zero = torch.zeros_like(wx)
    ents = [zero, -wz, wy, wz, zero, -wx, -wy, wx, zero]
    if batch_scramble:
        skew = torch.stack(ents, dim=0).reshape(-1, 3, 3)       # literal :45
    else:
        skew = torch.stack(ents, dim=-1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=wx.dtype).unsqueeze(0)
    return eye + ((1 - c) / (theta * theta)) * skew + ((theta - s) / theta.pow(3)) * (skew @ skew)


def camera_jacobian_matrix(x, y, Z, fx, fy, negate: bool = True) -> Tensor:
    """bundlenet.py:49-61 `CameraJacobianMatrix` -> [nb,N,2,6] (negated, :60).

    legacy/ba.py:36-48 is the same without the minus sign (`negate=False`).
    """
    xy = x * y
    xx = -1.0 - x * x
    x_z = x / Z
    yy = 1.0 + y * y
    y_z = y / Z
    iZ = 1.0 / Z
    zeros = torch.zeros_like(xy)
    dx = fx.unsqueeze(-1) * torch.stack([xy, xx, y, -iZ, zeros, x_z], dim=2)
    dy = fy.unsqueeze(-1) * torch.stack([yy, -xy, -x, zeros, -iZ, y_z], dim=2)
    J = torch.stack([dx, dy], dim=2)
    return -J if negate else J


def depth_jacobian_matrix(rx, ry, rz, x, y, Z, fx, fy) -> Tensor:
    """bundlenet.py:63-74 `DepthJacobianMatrix`.  rx,ry,rz [nb,1,N]; rest [nb,N] -> [nb,N,2]."""
    rx, ry, rz = rx.squeeze(1), ry.squeeze(1), rz.squeeze(1)
    dx = fx * ((rx - rz * x) / Z)
    dy = fy * ((ry - rz * y) / Z)
    return torch.stack([dx, dy], dim=2)


# --------------------------------------------------------------------------- image helpers
def grad_fixed(inp: Tensor) -> Tensor:
    """bundlenet.py:92-100 `grad_fixed`.  [nb,h,w,C] -> [nb,h,w,2C] = [gradx | grady].

    REFLECT pad by one (edge texel not repeated), central difference * 0.5.
    """
    nb, h, w, C = inp.shape
    pad = torch.nn.functional.pad(inp.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
    gradx = 0.5 * (pad[:, 1:h + 1, 2:w + 2, :] - pad[:, 1:h + 1, 0:w, :])
    grady = 0.5 * (pad[:, 2:h + 2, 1:w + 1, :] - pad[:, 0:h, 1:w + 1, :])
    return torch.cat([gradx, grady], dim=-1)


def compute_coordinates(points2d: Tensor, fx, fy, ox, oy, normalize: bool = True) -> Tensor:
    """bundlenet.py:112-120 `computeCoordinates` -> p [nb,3,N], L2-normalised (:119).

    legacy/ba.py:27-34 is the un-normalised variant (`normalize=False`).
    """
    x = ((points2d[:, :, 0] - ox) / fx).unsqueeze(1)
    y = ((points2d[:, :, 1] - oy) / fy).unsqueeze(1)
    ones = torch.ones_like(x)
    p = torch.cat([x, y, ones], dim=1)
    if normalize:
        p = p / torch.sqrt(torch.clamp((p * p).sum(dim=1, keepdim=True), min=1e-12))   # tf.nn.l2_normalize eps
    return p


def resampler(data: Tensor, warp: Tensor) -> Tensor:
    """`tf.contrib.resampler.resampler` (third-party, TF<=1.15, not vendored): bilinear
    sampling of NHWC `data` [nb,h,w,C] at `warp` [nb,N,2]=(x,y), texels outside the map read as 0.
    Call sites bundlenet.py:154,230,290,320,343,344,385.  Wherever the in-bounds mask is 1 this
    is identical to the in-repo gather sampler legacy/utils_python.py:61-117 (clamped indices).
    """
    nb, h, w, C = data.shape
    x, y = warp[..., 0], warp[..., 1]
    x0f, y0f = torch.floor(x), torch.floor(y)
    dx, dy = x - x0f, y - y0f
    x0, y0 = x0f.long(), y0f.long()
    flat = data.reshape(nb, h * w, C)
    out = 0
    for (xi, yi, wgt) in ((x0, y0, (1 - dx) * (1 - dy)), (x0 + 1, y0, dx * (1 - dy)),
                          (x0, y0 + 1, (1 - dx) * dy), (x0 + 1, y0 + 1, dx * dy)):
        valid = ((xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)).to(data.dtype)
        idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).unsqueeze(-1).expand(-1, -1, C)
        out = out + torch.gather(flat, 1, idx) * (wgt * valid).unsqueeze(-1)
    return out


def interpolate2d(imgs: Tensor, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """legacy/utils_python.py:61-117 `interpolate2d`: gather-based bilinear with CLAMPED indices
    (:96-99) and mask = (x == clip(x,0,w-1)) & (y == clip(y,0,h-1)) (:114-116).  -> ([nb,N,C],[nb,N,1])
    """
    nb, h, w, C = imgs.shape
    x0f, y0f = torch.floor(x), torch.floor(y)
    dx, dy = x - x0f, y - y0f
    x0 = x0f.long().clamp(0, w - 1); x1 = (x0f.long() + 1).clamp(0, w - 1)
    y0 = y0f.long().clamp(0, h - 1); y1 = (y0f.long() + 1).clamp(0, h - 1)
    flat = imgs.reshape(nb, h * w, C)
    def g(yi, xi):
        return torch.gather(flat, 1, (yi * w + xi).unsqueeze(-1).expand(-1, -1, C))
    out = (g(y0, x0) * ((1 - dx) * (1 - dy)).unsqueeze(-1) + g(y0, x1) * (dx * (1 - dy)).unsqueeze(-1)
           + g(y1, x0) * ((1 - dx) * dy).unsqueeze(-1) + g(y1, x1) * (dx * dy).unsqueeze(-1))
    mask = ((x == x.clamp(0.0, w - 1.0)) & (y == y.clamp(0.0, h - 1.0))).to(imgs.dtype).unsqueeze(-1)
    return out, mask


# --------------------------------------------------------------------------- lambda MLP
def selu(x: Tensor) -> Tensor:
    return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * (torch.exp(x) - 1))


def lambda_mlp(avg_residual: Tensor, params: Sequence[Tuple[Tensor, Tensor]]) -> Tensor:
    """bundlenet.py:244-248 (pose-only :168-172): five `conv1d` layers with kernel width 1
    (:102-110, i.e. dense layers) C->2C->4C->2C->C->1, selu x4 then tanh.
    params[i] = (filters[cin,cout], biases[cout])  (TF filter shape [1,cin,cout]).
    avg_residual [nb,1,C] -> [nb,1,1]
    """
    h = avg_residual
    for i, (Wt, b) in enumerate(params):
        h = h @ Wt + b
        h = torch.tanh(h) if i == len(params) - 1 else selu(h)
    return h


def init_lambda_mlp(C: int, seed: int = 7, dtype=torch.float64) -> List[Tuple[Tensor, Tensor]]:
    """he_normal filters / zero biases as in bundlenet.py:105-106 (seeded for tests)."""
    g = torch.Generator().manual_seed(seed)
    dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
    out = []
    for i in range(5):
        std = math.sqrt(2.0 / dims[i])
        # tf.keras he_normal draws from a truncated normal; the init is not part of parity.
        Wt = (torch.randn(dims[i], dims[i + 1], generator=g, dtype=torch.float64) * std).to(dtype)
        out.append((Wt, torch.zeros(dims[i + 1], dtype=dtype)))
    return out


# --------------------------------------------------------------------------- the native op
def equation_construction(jacobian: Tensor, gradient: Tensor, difference: Tensor) -> Tuple[Tensor, Tensor]:
    """utils.cu:150-171 (op) / :219-417 (kernel): AtA[b] = sum_n J^T G^T G J, Atb[b] = sum_n J^T G^T d.

    jacobian [nb,N,2,P], gradient [nb,N,C,2], difference [nb,N,C,1] -> left [nb,P,P], right [nb,P,1].
    Written as the reference's GEMM chain: M=G^T G (:331-340), J^T M (:344-353), J^T M J (:356-365),
    column-reduce over N (:380); d^T G (:382-391), d^T G J (:393-402), reduce (:414).
    In-repo pure-TF twin: legacy/ba.py:197-198.
    """
    M = gradient.transpose(-1, -2) @ gradient                       # [nb,N,2,2]
    JtM = jacobian.transpose(-1, -2) @ M                            # [nb,N,P,2]
    left = (JtM @ jacobian).sum(dim=1)                              # [nb,P,P]
    dtG = difference.transpose(-1, -2) @ gradient                   # [nb,N,1,2]
    right = (dtG @ jacobian).sum(dim=1).transpose(-1, -2)           # [nb,P,1]
    return left, right


def equation_construction_grad(jacobian, gradient, difference, left_grad, right_grad):
    """utils.cu:420-428 (op) / :465-694 (kernel), registered bundlenet.py:79-82.

    With A = G J (:625-634):  dA = 2 A Ghat + d ghat^T (:648-668 — NOT A(Ghat+Ghat^T); exact
    only for a symmetric upstream gradient), dd = A ghat (:636-645), dJ = G^T dA (:670-679),
    dG = dA J^T (:681-690).  Ghat/ghat are first tiled to every pixel (tile_kernel :442-463).
    """
    A = gradient @ jacobian                                          # [nb,N,C,P]
    Gh = left_grad.unsqueeze(1)                                      # [nb,1,P,P]
    gh = right_grad.unsqueeze(1)                                     # [nb,1,P,1]
    dd = A @ gh                                                      # [nb,N,C,1]
    dA = 2.0 * (A @ Gh) + difference @ gh.transpose(-1, -2)          # [nb,N,C,P]
    dJ = gradient.transpose(-1, -2) @ dA                             # [nb,N,2,P]
    dG = dA @ jacobian.transpose(-1, -2)                             # [nb,N,C,2]
    return dJ, dG, dd


class _EquationConstructionRefGrad(torch.autograd.Function):
    """equation_construction whose backward is the reference's registered gradient
    (bundlenet.py:79-82) instead of exact autodiff."""

    @staticmethod
    def forward(ctx, J, G, d):
        ctx.save_for_backward(J, G, d)
        return equation_construction(J, G, d)

    @staticmethod
    def backward(ctx, gl, gr):
        J, G, d = ctx.saved_tensors
        return equation_construction_grad(J, G, d, gl, gr)


# --------------------------------------------------------------------------- one LM iteration
@dataclass
class IterOptions:
    l2_regularizer_base: Optional[float] = 1000.0   # bundlenet.py:393 (BundleIteration); 1.0/unused for CameraIteration :326
    damping_eps: float = 1e-5                       # bundlenet.py:182,266
    undamped_last: bool = True                      # bundlenet.py:266 (last depth coefficient gets 0)
    vmatrix_batch_scramble: bool = False            # bundlenet.py:45 quirk
    guard_nonfinite: bool = False                   # CUDA path masks non-finite projections; reference would emit NaN
    reference_op_grad: bool = False                 # use utils.cu's 2*A*Ghat gradient for the op
    lambda_override: Optional[Tensor] = None        # [nb] — bypass the MLP (op-level tests)


def _warp(p, Dt, R, T, fx, fy, ox, oy):
    """bundlenet.py:209-224 (pose-only :136-148)."""
    npix = p.shape[2]
    Rp = R @ p                                                       # [nb,3,N]
    RP = Rp * Dt.transpose(1, 2).expand(-1, 3, -1)
    RPT = RP + T.expand(-1, -1, npix)
    X, Y, Z = RPT[:, 0, :], RPT[:, 1, :], RPT[:, 2, :]
    x, y = X / Z, Y / Z
    px, py = fx * x + ox, fy * y + oy
    return Rp, x, y, Z, px, py


def _sample_diff_grad(conv1, conv2, px, py, guard_nonfinite=False):
    """bundlenet.py:226-239 (pose-only :150-163): sample, in-bounds mask, diff and grad."""
    nb, h, w, C3 = conv2.shape
    C = conv1.shape[2]
    if guard_nonfinite:
        finite = torch.isfinite(px) & torch.isfinite(py)
        px = torch.where(finite, px, torch.full_like(px, -1.0))
        py = torch.where(finite, py, torch.full_like(py, -1.0))
    s = resampler(conv2, torch.stack([px, py], dim=-1))
    oob = (px < 0) | (px > float(w - 1)) | (py < 0) | (py > float(h - 1))
    m = (~oob).to(conv1.dtype).unsqueeze(-1).unsqueeze(-1)           # [nb,N,1,1]
    _diff = (conv1 - s[:, :, 0:C]).unsqueeze(-1)
    _gx = s[:, :, C:2 * C].unsqueeze(-1)
    _gy = s[:, :, 2 * C:3 * C].unsqueeze(-1)
    diff = _diff @ m                                                 # [nb,N,C,1]
    grad = torch.cat([_gx @ m, _gy @ m], dim=-1)                     # [nb,N,C,2]
    return diff, grad, m


def _lambda(diff, mlp_params, opts: IterOptions, scale_by_base: bool):
    """bundlenet.py:241-253 (pose-only :165-173)."""
    avg_residual = diff.squeeze(-1).abs().mean(dim=1, keepdim=True)  # [nb,1,C]  (divides by N, :243)
    if opts.lambda_override is not None:
        lam = opts.lambda_override.reshape(-1, 1, 1).to(diff.dtype)
        return lam, avg_residual
    h5 = lambda_mlp(avg_residual, mlp_params)
    lam = torch.pow(torch.linalg.norm(avg_residual, dim=-1, keepdim=True), 2.0 + h5)
    if scale_by_base and opts.l2_regularizer_base is not None:
        lam = opts.l2_regularizer_base * lam
    return lam, avg_residual


def _update(solution, R, T, opts: IterOptions):
    """bundlenet.py:269-275 (pose-only :184-190)."""
    wx, wy, wz, tx, ty, tz = [solution[:, i:i + 1, :] for i in range(6)]
    dr = angle_axis_rotation(wx, wy, wz)
    dv = v_matrix(wx, wy, wz, opts.vmatrix_batch_scramble)
    dt = torch.cat([tx, ty, tz], dim=1)
    return dr @ R, dv @ dt + dr @ T


def bundle_iteration(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W, mlp_params,
                     opts: IterOptions = IterOptions(), return_aux: bool = False):
    """bundlenet.py:193-278 `BundleNet.BundleIteration` -> (R',T',W')."""
    nb = conv1.shape[0]
    Dt = D + B @ W                                                   # :208
    Rp, x, y, Z, px, py = _warp(p, Dt, R, T, fx, fy, ox, oy)
    rx, ry, rz = Rp[:, 0:1, :], Rp[:, 1:2, :], Rp[:, 2:3, :]          # :210
    diff, grad, m = _sample_diff_grad(conv1, conv2, px, py, opts.guard_nonfinite)
    lam, rbar = _lambda(diff, mlp_params, opts, scale_by_base=True)
    Jc = camera_jacobian_matrix(x, y, Z, fx, fy)                     # :259
    Jd = depth_jacobian_matrix(rx, ry, rz, x, y, Z, fx, fy).unsqueeze(-1) @ B.unsqueeze(-2)   # :260
    J = torch.cat([Jc, Jd], dim=-1)                                  # :261  [nb,N,2,6+K]
    eqc = _EquationConstructionRefGrad.apply if opts.reference_op_grad else equation_construction
    AtA, Atb = eqc(J, grad, diff)                                    # :263
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)                     # :264
    if opts.undamped_last:
        dvec = torch.cat([diag[:, :-1] + opts.damping_eps, torch.zeros(nb, 1, dtype=diag.dtype)], dim=-1)
    else:
        dvec = diag + opts.damping_eps
    AtA_d = AtA + torch.diag_embed(dvec * lam.squeeze(-1))           # :266
    solution = torch.linalg.solve(AtA_d, Atb)                        # :267 tf.matrix_solve (LU, partial pivoting)
    Rn, Tn = _update(solution[:, :6, :], R, T, opts)
    Wn = W + solution[:, 6:, :]                                      # :276
    if return_aux:
        return Rn, Tn, Wn, dict(AtA=AtA, Atb=Atb, lam=lam, rbar=rbar, solution=solution,
                                nvalid=m.reshape(nb, -1).sum(dim=1), J=J, grad=grad, diff=diff, px=px, py=py)
    return Rn, Tn, Wn


def camera_iteration(conv1, conv2, fx, fy, ox, oy, p, D, R, T, mlp_params,
                     opts: IterOptions = IterOptions(), return_aux: bool = False):
    """bundlenet.py:122-191 `BundleNet.CameraIteration` -> (R',T').  P = 6, every diagonal
    damped (:181-182), lambda NOT multiplied by l2_regularizer_base (accepted but unused)."""
    nb = conv1.shape[0]
    Rp, x, y, Z, px, py = _warp(p, D, R, T, fx, fy, ox, oy)
    diff, grad, m = _sample_diff_grad(conv1, conv2, px, py, opts.guard_nonfinite)
    lam, rbar = _lambda(diff, mlp_params, opts, scale_by_base=False)
    J = camera_jacobian_matrix(x, y, Z, fx, fy)
    eqc = _EquationConstructionRefGrad.apply if opts.reference_op_grad else equation_construction
    AtA, Atb = eqc(J, grad, diff)
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)
    AtA_d = AtA + torch.diag_embed(((diag.unsqueeze(-1) + opts.damping_eps) @ lam).squeeze(-1))   # :182
    motion = torch.linalg.solve(AtA_d, Atb)                          # :183
    Rn, Tn = _update(motion, R, T, opts)
    if return_aux:
        return Rn, Tn, dict(AtA=AtA, Atb=Atb, lam=lam, rbar=rbar, solution=motion,
                            nvalid=m.reshape(nb, -1).sum(dim=1), J=J, grad=grad, diff=diff)
    return Rn, Tn


# --------------------------------------------------------------------------- structured form
def normal_equations_structured(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W,
                                guard_nonfinite: bool = False):
    """Same AtA/Atb/rbar as bundle_iteration but WITHOUT forming J (the block decomposition
    the CUDA kernel uses):  with M=G^T G, q=G^T d, J=[Jc | jd b^T]
        H_cc = sum Jc^T M Jc      g_c = sum Jc^T q
        H_cd = sum (Jc^T M jd) b^T               g_d = sum (jd^T q) b
        H_dd = sum (jd^T M jd) b b^T
    Used to prove the decomposition exact (tests/test_oracle_consistency.py) and as the checker at BASELINE sizes.
    B is None -> pose-only (P=6).
    """
    nb, N, C = conv1.shape
    Dt = D if B is None else D + B @ W
    Rp, x, y, Z, px, py = _warp(p, Dt, R, T, fx, fy, ox, oy)
    diff, grad, m = _sample_diff_grad(conv1, conv2, px, py, guard_nonfinite)
    M = grad.transpose(-1, -2) @ grad                                # [nb,N,2,2]
    q = grad.transpose(-1, -2) @ diff                                # [nb,N,2,1]
    Jc = camera_jacobian_matrix(x, y, Z, fx, fy)                     # [nb,N,2,6]
    Hcc = (Jc.transpose(-1, -2) @ M @ Jc).sum(1)
    gc = (Jc.transpose(-1, -2) @ q).sum(1)
    rbar = diff.squeeze(-1).abs().mean(dim=1, keepdim=True)
    nvalid = m.reshape(nb, -1).sum(1)
    if B is None:
        return Hcc, gc, rbar, nvalid
    jd = depth_jacobian_matrix(Rp[:, 0:1], Rp[:, 1:2], Rp[:, 2:3], x, y, Z, fx, fy).unsqueeze(-1)  # [nb,N,2,1]
    v = (Jc.transpose(-1, -2) @ M @ jd).squeeze(-1)                  # [nb,N,6]
    s = (jd.transpose(-1, -2) @ M @ jd).reshape(nb, N)               # [nb,N]
    t = (jd.transpose(-1, -2) @ q).reshape(nb, N)
    Hcd = v.transpose(1, 2) @ B                                      # [nb,6,K]
    Hdd = B.transpose(1, 2) @ (s.unsqueeze(-1) * B)                  # [nb,K,K]
    gd = B.transpose(1, 2) @ t.unsqueeze(-1)                         # [nb,K,1]
    H = torch.cat([torch.cat([Hcc, Hcd], 2), torch.cat([Hcd.transpose(1, 2), Hdd], 2)], 1)
    g = torch.cat([gc, gd], 1)
    return H, g, rbar, nvalid


def normal_equations_structured_chunked(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W,
                                        guard_nonfinite: bool = False, chunk: int = 32768):
    """`normal_equations_structured` over point chunks (same sums, bounded memory): lets the float64 oracle run the
    BASELINE cfg2 level sizes (N = 307 200, C = K = 128) that the materialised J of bundlenet.py:259-261 cannot."""
    nb, N, C = conv1.shape
    H = g = None
    rbar = torch.zeros(nb, 1, C, dtype=conv1.dtype)
    nvalid = torch.zeros(nb, dtype=conv1.dtype)
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        Hc, gc, rb, nv = normal_equations_structured(conv1[:, a:b], conv2, fx[:, a:b], fy[:, a:b], ox[:, a:b], oy[:, a:b],
                                                     p[:, :, a:b], D[:, a:b], None if B is None else B[:, a:b], R, T, W,
                                                     guard_nonfinite)
        H = Hc if H is None else H + Hc
        g = gc if g is None else g + gc
        rbar += rb * float(b - a)                                    # structured() returns the chunk MEAN; keep the sum
        nvalid += nv
    return H, g, rbar / float(N), nvalid


def iteration_structured(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W, mlp_params,
                         opts: IterOptions = IterOptions(), chunk: int = 32768):
    """bundle_iteration / camera_iteration (bundlenet.py:193-278 / :122-191) with the normal equations taken from the
    chunked block form; identical maths (tests/test_oracle_consistency.py), usable at full BASELINE sizes."""
    nb = conv1.shape[0]
    bundle = B is not None
    AtA, Atb, avg_residual, _ = normal_equations_structured_chunked(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W,
                                                                    opts.guard_nonfinite, chunk)
    if opts.lambda_override is not None:
        lam = opts.lambda_override.reshape(-1, 1, 1).to(conv1.dtype)
    else:
        lam = torch.pow(torch.linalg.norm(avg_residual, dim=-1, keepdim=True), 2.0 + lambda_mlp(avg_residual, mlp_params))
        if bundle and opts.l2_regularizer_base is not None:
            lam = opts.l2_regularizer_base * lam                     # :252-253 (CameraIteration ignores the base)
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)
    if bundle and opts.undamped_last:
        dvec = torch.cat([diag[:, :-1] + opts.damping_eps, torch.zeros(nb, 1, dtype=diag.dtype)], dim=-1)   # :266
    else:
        dvec = diag + opts.damping_eps                               # :182
    solution = torch.linalg.solve(AtA + torch.diag_embed(dvec * lam.reshape(nb, 1)), Atb)
    Rn, Tn = _update(solution[:, :6, :], R, T, opts)
    return (Rn, Tn, W + solution[:, 6:, :]) if bundle else (Rn, Tn, None)


def lm_solve_structured(levels, iters_per_level: int, R, T, W=None, opts: IterOptions = IterOptions(), chunk: int = 32768):
    """`lm_solve` through `iteration_structured` (bounded memory)."""
    for lv in levels:
        for _ in range(iters_per_level):
            R, T, W = iteration_structured(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B, R, T, W, lv.mlp, opts, chunk)
    return R, T, W


# --------------------------------------------------------------------------- joint keyframe window (extension)
def window_assemble(H: Tensor, g: Tensor) -> Tuple[Tensor, Tensor]:
    """Per-pair normal equations of nf pairs that share one W -> the block-arrow system of the window: pose blocks and pose-depth
    couplings per frame, depth block and depth right-hand side summed over the frames.  H [nf,P,P], g [nf,P,1] -> Hj [Pj,Pj], gj [Pj,1],
    Pj = 6 nf + K."""
    nf, P, _ = H.shape
    K = P - 6
    Pj = 6 * nf + K
    Hj = torch.zeros(Pj, Pj, dtype=H.dtype); gj = torch.zeros(Pj, 1, dtype=H.dtype)
    for f in range(nf):
        s = slice(6 * f, 6 * f + 6)
        Hj[s, s] = H[f, :6, :6]
        Hj[s, 6 * nf:] = H[f, :6, 6:]
        Hj[6 * nf:, s] = H[f, 6:, :6]
        Hj[6 * nf:, 6 * nf:] += H[f, 6:, 6:]
        gj[s] = g[f, :6]
        gj[6 * nf:] += g[f, 6:]
    return Hj, gj


def window_iteration(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W, mlp_params, opts: IterOptions = IterOptions(), chunk: int = 32768):
    """One joint LM iteration of a keyframe window.  NOT in the reference (its BA layer is 2-view, bundlenet.py:193-278; SURVEY.md
    section 8f-4 lists the window as an extension): the nf pairs (keyframe -> frame f) share the keyframe depth D + B.W, W [K,1].
    Everything else follows BundleIteration: per-pair residuals / Jacobians (:206-261), lambda from the mean |residual| over all points
    of all frames (:241-253), damping of every diagonal entry but the last depth coefficient (:264-266), one solve (:267), per-frame
    SE(3) update (:269-275), shared W update (:276).  -> (R' [nf,3,3], T' [nf,3,1], W' [K,1])."""
    nf = conv1.shape[0]
    K = B.shape[-1]
    Wrep = W.reshape(1, K, 1).expand(nf, K, 1)
    H, g, rbar, _ = normal_equations_structured_chunked(conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, Wrep, opts.guard_nonfinite, chunk)
    Hj, gj = window_assemble(H, g)
    avg = rbar.mean(dim=0, keepdim=True)                              # every frame has N points: mean over all nf * N
    if opts.lambda_override is not None:
        lam = opts.lambda_override.reshape(-1)[0].to(conv1.dtype)
    else:
        lam = torch.pow(torch.linalg.norm(avg, dim=-1, keepdim=True), 2.0 + lambda_mlp(avg, mlp_params)).reshape(())
        if opts.l2_regularizer_base is not None:
            lam = opts.l2_regularizer_base * lam
    diag = torch.diagonal(Hj)
    dvec = diag + opts.damping_eps
    if opts.undamped_last:
        dvec = torch.cat([dvec[:-1], torch.zeros(1, dtype=diag.dtype)])
    sol = torch.linalg.solve(Hj + torch.diag(dvec * lam), gj)
    pose = sol[:6 * nf].reshape(nf, 6, 1)
    Rn, Tn = _update(pose, R, T, opts)
    return Rn, Tn, W.reshape(K, 1) + sol[6 * nf:]


def window_solve(levels, iters_per_level: int, R, T, W, opts: IterOptions = IterOptions(), chunk: int = 32768):
    """Coarse-to-fine loop of `window_iteration` over `LevelInputs` (nb = nf pairs per level)."""
    for lv in levels:
        for _ in range(iters_per_level):
            R, T, W = window_iteration(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B, R, T, W, lv.mlp, opts, chunk)
    return R, T, W


# --------------------------------------------------------------------------- schedulers
@dataclass
class ResizeGeometry:
    """Hard-coded crop / intrinsics fix-ups of bundlenet.py:286-287,298-302 (=:338-339,354-357)
    and the half-resolution output shape of :397."""
    sx: float = 320.0; cx: float = 4.0; dx: float = 312.0          # x = 320*(x-4)/312
    sy: float = 256.0; cy: float = 4.0; dy: float = 232.0          # y = 256*(y-4)/232
    fx_num: float = 40.0; fx_den: float = 39.0; ox_sub: float = 160.0 / 39.0
    fy_num: float = 32.0; fy_den: float = 29.0; oy_sub: float = 128.0 / 29.0
    out_hw: Tuple[int, int] = (256 // 2, 320 // 2)


def _prepare(intrisic, points, geo: ResizeGeometry):
    npix = points.shape[1]
    x = geo.sx * (points[..., 0:1] - geo.cx) / geo.dx
    y = geo.sy * (points[..., 1:2] - geo.cy) / geo.dy
    _points = torch.cat([x, y], dim=-1)
    fx = geo.fx_num * intrisic[:, 0].repeat(1, npix) / geo.fx_den
    fy = geo.fy_num * intrisic[:, 1].repeat(1, npix) / geo.fy_den
    ox = geo.fx_num * intrisic[:, 2].repeat(1, npix) / geo.fx_den - geo.ox_sub
    oy = geo.fy_num * intrisic[:, 3].repeat(1, npix) / geo.fy_den - geo.oy_sub
    return _points, fx, fy, ox, oy


def _swap_halves(x):
    nb = x.shape[0]
    return torch.cat([x[nb // 2:nb], x[0:nb // 2]], dim=0)           # bundlenet.py:321,386


def bundle_resize(intrisic, layers, points, basis, init_depth, mlp_params_by_level,
                  init_rotation=None, init_translation=None, opts: IterOptions = IterOptions(),
                  geo: ResizeGeometry = ResizeGeometry()):
    """bundlenet.py:332-399 `BundleNet.BundleResize` -> (Rs, Ts, depths), one entry per level (2,3).

    intrisic [nb,4,1]; layers: 4 x [nb,h_l,w_l,C]; points [nb,N,2]; basis [nb,h/2,w/2,K];
    init_depth [nb,h/2,w/2,1].  mlp_params_by_level: {"2": params, "3": params}.
    """
    nb = layers[-1].shape[0]
    K = basis.shape[-1]
    _points, sfx, sfy, sox, soy = _prepare(intrisic, points, geo)
    depths = init_depth.detach()                                     # :341
    d = resampler(depths, _points / 2)                               # :343
    b = resampler(basis, _points / 2)                                # :344
    p = compute_coordinates(_points, sfx, sfy, sox, soy)             # :358
    dt = layers[-1].dtype
    R = torch.eye(3, dtype=dt).repeat(nb, 1, 1) if init_rotation is None else init_rotation
    T = torch.zeros(nb, 3, 1, dtype=dt) if init_translation is None else init_translation
    W = torch.zeros(nb, K, 1, dtype=dt)
    Rs, Ts, Ds = [], [], []
    for level in range(2, 4):                                        # :376
        scale = 2 ** (3 - level)
        fx, fy, ox, oy = sfx / scale, sfy / scale, sox / scale, soy / scale
        points1 = _points / scale
        layer1 = resampler(layers[level], points1)                   # :385
        layer2 = _swap_halves(layers[level])                         # :386
        layer2 = torch.cat([layer2, grad_fixed(layer2)], dim=-1)     # :388-389
        for _ in range(1):                                           # :391
            R, T, W = bundle_iteration(layer1, layer2, fx, fy, ox, oy, p, d, b, R, T, W,
                                       mlp_params_by_level[str(level)], opts)
            Rs.append(R); Ts.append(T)
            Ds.append(init_depth + (basis.reshape(nb, -1, K) @ W).reshape(nb, geo.out_hw[0], geo.out_hw[1], 1))  # :397
    return Rs, Ts, Ds


def camera_resize(intrisic, layers, points, _depths, mlp_params_by_level,
                  opts: IterOptions = IterOptions(), geo: ResizeGeometry = ResizeGeometry()):
    """bundlenet.py:280-329 `BundleNet.CameraResize` -> (rotations, translations), levels 0..3 x 1 iter."""
    nb = layers[-1].shape[0]
    _points, sfx, sfy, sox, soy = _prepare(intrisic, points, geo)
    d = resampler(_depths.detach(), _points / 2)                     # :289-290
    p = compute_coordinates(_points, sfx, sfy, sox, soy)
    dt = layers[-1].dtype
    rotations = [torch.eye(3, dtype=dt).repeat(nb, 1, 1)]
    translations = [torch.zeros(nb, 3, 1, dtype=dt)]
    for level in range(0, 4):                                        # :309
        scale = 2 ** (3 - level)
        fx, fy, ox, oy = sfx / scale, sfy / scale, sox / scale, soy / scale
        layer1 = resampler(layers[level], _points / scale)
        layer2 = _swap_halves(layers[level])
        layer2 = torch.cat([layer2, grad_fixed(layer2)], dim=-1)
        R, T = camera_iteration(layer1, layer2, fx, fy, ox, oy, p, d, rotations[-1], translations[-1],
                                mlp_params_by_level[str(level)], opts)
        rotations.append(R); translations.append(T)
    return rotations[1:], translations[1:]


@dataclass
class LevelInputs:
    """One pyramid level of the BASELINE LM solve: its own point set."""
    conv1: Tensor; conv2: Tensor
    fx: Tensor; fy: Tensor; ox: Tensor; oy: Tensor
    p: Tensor; D: Tensor; B: Optional[Tensor]
    mlp: Sequence[Tuple[Tensor, Tensor]] = field(default_factory=list)


def lm_solve(levels: Sequence[LevelInputs], iters_per_level: int, R, T, W=None,
             opts: IterOptions = IterOptions()):
    """The BASELINE.json workload: coarse->fine over `levels`, `iters_per_level` BundleIterations
    (or CameraIterations when W is None) each — the level_iters loop of legacy/ba.py:106-121 applied to
    bundlenet.py:391-393.  Returns (R,T,W) after the last iteration.
    """
    for lv in levels:
        for _ in range(iters_per_level):
            if W is None:
                R, T = camera_iteration(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, R, T, lv.mlp, opts)
            else:
                R, T, W = bundle_iteration(lv.conv1, lv.conv2, lv.fx, lv.fy, lv.ox, lv.oy, lv.p, lv.D, lv.B,
                                           R, T, W, lv.mlp, opts)
    return R, T, W


# --------------------------------------------------------------------------- legacy LM policy
def legacy_camera_iteration2(conv1, conv2, fx, fy, ox, oy, p, D, R, T, mlp_params,
                             residual_ratio: float = 1.0, use_qr: bool = True,
                             opts: IterOptions = IterOptions()):
    """legacy/ba.py:226-345 `Tracker.CameraIteration2`: pose-only step with lambda-MLP, then the
    residual is re-evaluated at the updated pose and the step is kept only if it decreased
    (:304-345).  Legacy conventions: diff = F2w - conv1 (:263), un-negated camera Jacobian
    (ba.py:36-48), rbar rescaled by N/valid (:256,274), lambda = ||rbar||^(1+tanh) (:280),
    clamped-index sampler `interpolate2d`.  nb == 1 in the reference.
    Returns (R,T,update_w,update_t,num_valid_ratio).
    """
    C = conv1.shape[2]
    npix = conv1.shape[1]

    def residual(Rm, Tm):
        Rp, x, y, Z, px, py = _warp(p, D, Rm, Tm, fx, fy, ox, oy)
        s, mask = interpolate2d(conv2, px, py)
        num_valid = npix / mask.sum(dim=1, keepdim=True)
        return s, mask, num_valid, x, y, Z

    s, _mask, num_valid, x, y, Z = residual(R, T)
    mask = _mask.unsqueeze(-1)
    diff = (s[:, :, 0:C] - conv1).unsqueeze(-1) @ mask
    grad = torch.cat([s[:, :, C:2 * C].unsqueeze(-1) @ mask, s[:, :, 2 * C:3 * C].unsqueeze(-1) @ mask], -1)
    avg_residual = num_valid * diff.squeeze(-1).abs().mean(dim=1, keepdim=True)
    h5 = lambda_mlp(avg_residual, mlp_params)
    lam = torch.pow(torch.linalg.norm(avg_residual, dim=-1, keepdim=True), 1.0 + h5)
    avg_scalar = avg_residual.mean()
    J = camera_jacobian_matrix(x, y, Z, fx, fy, negate=False)
    AtA, Atb = equation_construction(J, grad, diff)                  # ba.py:282-283
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)
    AtA = AtA + torch.diag_embed(((diag.unsqueeze(-1) + 1e-5) @ lam).squeeze(-1))
    if use_qr:
        q, r = torch.linalg.qr(AtA, mode="complete")                 # ba.py:292-293
        motion = torch.linalg.solve(r, q.transpose(-1, -2) @ Atb)
    else:
        motion = torch.linalg.inv(AtA) @ Atb
    Rn, Tn = _legacy_update(motion, R, T, with_vmatrix=True)          # ba.py:297-302 (un-clamped AngleaAxisRotation :60-80)
    s2, mask2, num_valid2, *_ = residual(Rn, Tn)
    avg2 = (num_valid2 * (mask2 * (s2[:, :, 0:C] - conv1)).abs().mean(dim=1, keepdim=True)).mean()
    if bool(avg2 < residual_ratio * avg_scalar):                     # ba.py:343
        m = motion.reshape(-1)
        return Rn, Tn, torch.linalg.norm(m[:3]), torch.linalg.norm(m[3:]), num_valid.squeeze()
    z = torch.zeros((), dtype=conv1.dtype)
    return R, T, z, z, num_valid.squeeze()


def interpolate2d2(imgs: Tensor, p: Tensor) -> Tensor:
    """legacy/utils_python.py:177-232 `interpolate2d2`: `interpolate2d` on points [nb,N,2] without the mask."""
    return interpolate2d(imgs, p[:, :, 0], p[:, :, 1])[0]


def _legacy_update(motion, R, T, with_vmatrix: bool):
    """legacy/ba.py:208-213 (`CameraIteration`: T' = t + dr T) and :297-302 (`CameraIteration2`: T' = V t + dr T); un-clamped rotation."""
    wx, wy, wz, tx, ty, tz = [motion[:, i:i + 1, :] for i in range(6)]       # tf.unstack(axis=1) -> [nb,1] each; same values
    dr = angle_axis_rotation(wx, wy, wz, clamp=False)
    dt = torch.cat([tx, ty, tz], dim=1)
    if with_vmatrix:
        dt = v_matrix(wx, wy, wz) @ dt
    return dr @ R, dt + dr @ T


def legacy_camera_iteration(conv1, conv2, fx, fy, ox, oy, p, D, R, T, use_qr: bool = True):
    """legacy/ba.py:147-214 `Tracker.CameraIteration`: pose-only step, lambda = ||rbar||^2 (no MLP, :190), normal equations as plain
    matmuls + reduce_sum (:197-198, the in-repo statement of the custom op), QR solve (:203-206), T' = t + dr T (no V matrix, :213).
    Returns (R', T', valid_fraction)."""
    C = conv1.shape[2]; npix = conv1.shape[1]
    Rp, x, y, Z, px, py = _warp(p, D, R, T, fx, fy, ox, oy)
    s, _mask = interpolate2d(conv2, px, py)
    mask = _mask.unsqueeze(-1)
    diff = (s[:, :, 0:C] - conv1).unsqueeze(-1) @ mask
    grad = torch.cat([s[:, :, C:2 * C].unsqueeze(-1) @ mask, s[:, :, 2 * C:3 * C].unsqueeze(-1) @ mask], -1)
    avg_residual = diff.squeeze(-1).abs().mean(dim=1, keepdim=True)
    lam = torch.pow(torch.linalg.norm(avg_residual, dim=-1, keepdim=True), 2.0)
    J = camera_jacobian_matrix(x, y, Z, fx, fy, negate=False)
    AtA = (J.transpose(-1, -2) @ ((grad.transpose(-1, -2) @ grad) @ J)).sum(dim=1)        # :197
    Atb = (J.transpose(-1, -2) @ (grad.transpose(-1, -2) @ diff)).sum(dim=1)              # :198
    diag = torch.diagonal(AtA, dim1=-2, dim2=-1)
    AtA = AtA + torch.diag_embed(((diag.unsqueeze(-1) + 1e-5) @ lam).squeeze(-1))
    if use_qr:
        q, r = torch.linalg.qr(AtA, mode="complete")
        motion = torch.linalg.solve(r, q.transpose(-1, -2) @ Atb)
    else:
        motion = torch.linalg.inv(AtA) @ Atb
    Rn, Tn = _legacy_update(motion, R, T, with_vmatrix=False)
    return Rn, Tn, mask.sum() / npix


def legacy_track(intrisic, layers, points, d, initR, initT, level_iters, mlp_params_by_level, early_termination: bool = True,
                 angle_change: float = 0.002 * (3.14 / 180.0), translation_change: float = 0.0002, residual_ratio: float = 1.0):
    """legacy/ba.py:83-145 `Tracker.trackTF`: keyframe -> frame pose tracking over pyramid levels 1..3 (scale 2^(3-level)).
    layers[l] is a 2-image batch (0 = keyframe, 1 = current frame, :112-113); un-normalised rays (:27-34); per level either
    `level_iters[level-1]` plain CameraIterations (:117-121) or the early-terminated loop of CameraIteration2 (:123-141): continue
    while iters < level_iters[level-1] and update_w > angle_change and update_t > translation_change (:132-133; a rejected step
    returns 0 updates and therefore ends the level).  Returns (R, T, ratio) / (Rs, Ts, ratio) like the reference."""
    npix = points.shape[1]
    sfx, sfy = intrisic[:, 0].repeat(1, npix), intrisic[:, 1].repeat(1, npix)
    sox, soy = intrisic[:, 2].repeat(1, npix), intrisic[:, 3].repeat(1, npix)
    p = compute_coordinates(points, sfx, sfy, sox, soy, normalize=False)
    R, T = initR, initT
    rotations, translations = [initR], [initT]
    ratio = torch.tensor(1.0, dtype=points.dtype)
    for level in range(1, 4):
        scale = 2 ** (3 - level)
        fx, fy, ox, oy = sfx / scale, sfy / scale, sox / scale, soy / scale
        layer1 = interpolate2d2(layers[level - 1][0:1], points / scale)                   # :112
        layer2 = layers[level - 1][1:2]
        layer2 = torch.cat([layer2, grad_fixed(layer2)], dim=-1)                          # :113-115
        if not early_termination:
            for _ in range(level_iters[level - 1]):
                R, T, ratio = legacy_camera_iteration(layer1, layer2, fx, fy, ox, oy, p, d, rotations[-1], translations[-1])
                rotations.append(R); translations.append(T)
        else:
            iters, update_w, update_t = 0, 1.0, 1.0
            while iters < level_iters[level - 1] and angle_change < float(update_w) and translation_change < float(update_t):
                R, T, update_w, update_t, ratio = legacy_camera_iteration2(layer1, layer2, fx, fy, ox, oy, p, d, R, T,
                                                                           mlp_params_by_level[str(level)], residual_ratio)
                iters += 1
    if not early_termination:
        return rotations[1:], translations[1:], ratio
    return R, T, ratio


# --------------------------------------------------------------------------- training losses (bundlenet.py:6-15, 401-463)
def rotation2quaternion(R: Tensor) -> Tensor:
    """bundlenet.py:6-15 `rotation2quaternion`: [nb,3,3] -> unit quaternion [nb,4] (w first; assumes 1 + trace > 0)."""
    diag = 1.0 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    q0 = torch.sqrt(diag) / 2.0
    q = torch.stack([q0, (R[:, 2, 1] - R[:, 1, 2]) / (4.0 * q0), (R[:, 0, 2] - R[:, 2, 0]) / (4.0 * q0), (R[:, 1, 0] - R[:, 0, 1]) / (4.0 * q0)], dim=1)
    return q / torch.sqrt(torch.clamp((q * q).sum(dim=1, keepdim=True), min=1e-12))


def loss_r(predQ: Tensor, gtQ: Tensor) -> Tensor:
    """bundlenet.py:401-404 `lossR`: tf.losses.cosine_distance of quaternions = mean(1 - <pred, gt>)."""
    return (1.0 - (predQ * gtQ).sum(dim=1, keepdim=True)).mean()


def loss_t(predT: Tensor, gtT: Tensor) -> Tensor:
    """bundlenet.py:411-413 (the second `lossT` definition overrides the angular one at :406-409): mean |predT - gtT|."""
    return (predT - gtT).abs().mean()


def loss_f(intrisic: Tensor, depth: Tensor, mask: Tensor, predR: Tensor, predT: Tensor, gtR: Tensor, gtT: Tensor,
           geo: ResizeGeometry = ResizeGeometry()) -> Tensor:
    """bundlenet.py:415-463 `lossF`: mean masked |flow(pred) - flow(gt)| of the dense pixel grid, in units of the image width, rescaled by
    total / valid.  intrisic [nb,4,1], depth [nb,h,w,1], mask [nb,h,w,1], R [nb,3,3], T [nb,3]."""
    nb, h, w = depth.shape[0], depth.shape[1], depth.shape[2]
    npix = h * w
    predT, gtT = predT.unsqueeze(-1), gtT.unsqueeze(-1)
    m = mask.reshape(nb, npix)
    fx = geo.fx_num * intrisic[:, 0].repeat(1, npix) / geo.fx_den
    fy = geo.fy_num * intrisic[:, 1].repeat(1, npix) / geo.fy_den
    ox = geo.fx_num * intrisic[:, 2].repeat(1, npix) / geo.fx_den - geo.ox_sub
    oy = geo.fy_num * intrisic[:, 3].repeat(1, npix) / geo.fy_den - geo.oy_sub
    yy, xx = torch.meshgrid(torch.arange(h, dtype=depth.dtype), torch.arange(w, dtype=depth.dtype), indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=-1).unsqueeze(0).repeat(nb, 1, 1)
    p = compute_coordinates(pts, fx, fy, ox, oy)

    def flow(Rm, Tm):
        X = (Rm @ p) * depth.reshape(nb, 1, npix) + Tm
        return fx * (X[:, 0] / X[:, 2]) + ox, fy * (X[:, 1] / X[:, 2]) + oy

    fxp, fyp = flow(predR, predT); fxg, fyg = flow(gtR, gtT)
    return (float(npix * nb) / m.sum()) * (((fxp - fxg).abs() * m).mean() / w + ((fyp - fyg).abs() * m).mean() / w)
