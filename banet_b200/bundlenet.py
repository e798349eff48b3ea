"""Host-side mirror of the reference's BA-layer interface (reference bundlenet.py:86-399), backed by the
sm_100a kernels.  Same method names, argument order, tensor layouts and return values as the reference's
`BundleNet`; arithmetic happens in libbanet_sm100.so (no torch maths on the hot path, no CPU fallback).

Differences a reference user should know (all documented in DESIGN.md):
  * fx,fy,ox,oy may be passed as the reference does ([nb,N], constant along N) — column 0 is used;
  * lambda-MLP weights are ordinary parameters named like the TF variables
    (`lambda_{level}_{i}_filters` [cin,cout], `lambda_{level}_{i}_biases`, reference bundlenet.py:105-106);
  * `tf.matrix_solve` (LU) is replaced by a Cholesky factorisation; non-finite projections are masked
    instead of poisoning the sums with NaN; VMatrix is evaluated per pair unless
    `vmatrix_batch_scramble=True` (reference bundlenet.py:45 interleaves pairs for nb > 1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from . import autograd as _ag
from ._lib import PREC_AUTO

Tensor = torch.Tensor


def _intr_from_tiled(fx, fy, ox, oy) -> Tensor:
    """[nb,N] (reference) / [nb,1] / [nb] -> [nb,4]."""
    cols = []
    for t in (fx, fy, ox, oy):
        t = t.reshape(t.shape[0], -1)[:, 0]
        cols.append(t)
    return torch.stack(cols, dim=1).to(torch.float32).contiguous()


@dataclass
class ResizeGeometry:
    """Crop / intrinsics fix-ups hard-coded in the reference (bundlenet.py:286-287, 298-302, 397)."""
    sx: float = 320.0; cx: float = 4.0; dx: float = 312.0
    sy: float = 256.0; cy: float = 4.0; dy: float = 232.0
    fx_num: float = 40.0; fx_den: float = 39.0; ox_sub: float = 160.0 / 39.0
    fy_num: float = 32.0; fy_den: float = 29.0; oy_sub: float = 128.0 / 29.0
    out_hw: Tuple[int, int] = (256 // 2, 320 // 2)


class BundleNet(torch.nn.Module):
    """Drop-in for reference `BundleNet` (bundlenet.py:86).  `channels` = feature channels C of the pyramid."""

    def __init__(self, channels: int, levels: Sequence[str] = ("0", "1", "2", "3"), is_training: bool = True,
                 reuse_variables=None, vmatrix_batch_scramble: bool = False, precision: int = PREC_AUTO, seed: int = 7,
                 exact_sym_grad: bool = False):
        super().__init__()
        self.is_training = is_training
        self.reuse_variables = reuse_variables
        self.channels = channels
        self.vmatrix_batch_scramble = vmatrix_batch_scramble
        self.precision = precision
        self.exact_sym_grad = exact_sym_grad      # False: the reference's op gradient 2*A*Ghat (utils.cu:648); True: A(Ghat+Ghat^T)
        self.geo = ResizeGeometry()
        g = torch.Generator().manual_seed(seed)
        dims = [channels, 2 * channels, 4 * channels, 2 * channels, channels, 1]
        for lv in levels:
            for i in range(5):
                # he_normal filters, zero biases (reference bundlenet.py:105-106)
                w = torch.randn(dims[i], dims[i + 1], generator=g) * math.sqrt(2.0 / dims[i])
                self.register_parameter(f"lambda_{lv}_{i + 1}_filters", torch.nn.Parameter(w))
                self.register_parameter(f"lambda_{lv}_{i + 1}_biases", torch.nn.Parameter(torch.zeros(dims[i + 1])))

    # ---- helpers -------------------------------------------------------------------------------
    def mlp_params(self, level: str) -> List[Tuple[Tensor, Tensor]]:
        return [(getattr(self, f"lambda_{level}_{i}_filters"), getattr(self, f"lambda_{level}_{i}_biases")) for i in range(1, 6)]

    def mlp_packed(self, level: str) -> Tensor:
        return ops.pack_mlp([(w.detach(), b.detach()) for w, b in self.mlp_params(level)])

    def grad_fixed(self, input: Tensor, name=None) -> Tensor:
        """reference bundlenet.py:92-100: [nb,h,w,C] -> [nb,h,w,2C] = [gradx|grady]."""
        return ops.grad_fixed_concat(input)[..., input.shape[-1]:].contiguous()

    def computeCoordinates(self, points2d: Tensor, fx, fy, ox, oy) -> Tensor:
        """reference bundlenet.py:112-120 -> p [nb,3,N] (L2-normalised)."""
        return ops.compute_coordinates(points2d, _intr_from_tiled(fx, fy, ox, oy), normalize=True)

    # ---- one LM iteration ----------------------------------------------------------------------
    @staticmethod
    def _wants_grad(*tensors) -> bool:
        return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)

    def CameraIteration(self, conv1, conv2, fx, fy, ox, oy, p, D, R, T, l2_regularizer_base=None, level=None,
                        return_aux: bool = False, differentiable: Optional[bool] = None):
        """reference bundlenet.py:122-191 -> (updatedR, updatedT).  l2_regularizer_base accepted, unused (as there).
        differentiable: None = automatically when gradients are being recorded (training path, banet_b200/autograd.py)."""
        if differentiable or (differentiable is None and not return_aux and self._wants_grad(conv1, conv2, D, R, T, *self.parameters())):
            Rn, Tn, _ = _ag.iteration(conv1, conv2, _intr_from_tiled(fx, fy, ox, oy), p, D, None, R, T, None,
                                      self.mlp_params(str(level)), None, exact_sym=self.exact_sym_grad)
            return Rn, Tn
        lv = ops.Level(conv1, conv2, _intr_from_tiled(fx, fy, ox, oy), p, D, None)
        H, g, rbar, nvalid = ops.lm_build(lv, R, T, None, self.precision)
        lam = ops.lm_lambda(rbar, conv1.shape[1], self.mlp_packed(str(level)), 1.0)
        Rn, Tn, _, delta, status = ops.lm_solve_update(H, g, lam, R, T, None, undamped_last=False,
                                                       vmatrix_batch_scramble=self.vmatrix_batch_scramble)
        if return_aux:
            return Rn, Tn, dict(AtA=H, Atb=g, lam=lam, rbar_sum=rbar, nvalid=nvalid, solution=delta, status=status)
        return Rn, Tn

    def BundleIteration(self, conv1, conv2, fx, fy, ox, oy, p, D, B, R, T, W, l2_regularizer_base=None, level=None,
                        return_aux: bool = False, differentiable: Optional[bool] = None):
        """reference bundlenet.py:193-278 -> (updatedR, updatedT, updatedW)."""
        if differentiable or (differentiable is None and not return_aux and self._wants_grad(conv1, conv2, D, B, R, T, W, *self.parameters())):
            return _ag.iteration(conv1, conv2, _intr_from_tiled(fx, fy, ox, oy), p, D, B, R, T, W,
                                 self.mlp_params(str(level)), l2_regularizer_base, exact_sym=self.exact_sym_grad)
        lv = ops.Level(conv1, conv2, _intr_from_tiled(fx, fy, ox, oy), p, D, B)
        H, g, rbar, nvalid = ops.lm_build(lv, R, T, W, self.precision)
        base = 1.0 if l2_regularizer_base is None else float(l2_regularizer_base)      # :252-253
        lam = ops.lm_lambda(rbar, conv1.shape[1], self.mlp_packed(str(level)), base)
        Rn, Tn, Wn, delta, status = ops.lm_solve_update(H, g, lam, R, T, W, undamped_last=True,
                                                        vmatrix_batch_scramble=self.vmatrix_batch_scramble)
        if return_aux:
            return Rn, Tn, Wn, dict(AtA=H, Atb=g, lam=lam, rbar_sum=rbar, nvalid=nvalid, solution=delta, status=status)
        return Rn, Tn, Wn

    # ---- level schedulers ----------------------------------------------------------------------
    def _prepare(self, intrisic: Tensor, points: Tensor):
        geo = self.geo
        x = geo.sx * (points[..., 0:1] - geo.cx) / geo.dx
        y = geo.sy * (points[..., 1:2] - geo.cy) / geo.dy
        _points = torch.cat([x, y], dim=-1).contiguous()
        k = intrisic.reshape(intrisic.shape[0], 4)
        intr = torch.stack([geo.fx_num * k[:, 0] / geo.fx_den, geo.fy_num * k[:, 1] / geo.fy_den,
                            geo.fx_num * k[:, 2] / geo.fx_den - geo.ox_sub,
                            geo.fy_num * k[:, 3] / geo.fy_den - geo.oy_sub], dim=1).contiguous()
        return _points, intr

    def CameraResize(self, intrisic, layers, points, _depths, reuse_variables=False):
        """reference bundlenet.py:280-329 -> (rotations, translations), levels 0..3 x 1 iteration."""
        nb = layers[-1].shape[0]
        _points, intr = self._prepare(intrisic, points)
        d = ops.resample(_depths.detach(), _points, 0.5)                       # :289-290
        p = ops.compute_coordinates(_points, intr, True)
        R = torch.eye(3, device=points.device).repeat(nb, 1, 1)
        T = torch.zeros(nb, 3, 1, device=points.device)
        rotations, translations = [], []
        for level in range(0, 4):
            scale = 2 ** (3 - level)
            layer1 = ops.resample(layers[level], _points, 1.0 / scale)         # :320
            layer2 = ops.grad_fixed_concat(layers[level], swap_halves=True)    # :321-324
            lv = ops.Level(layer1, layer2, intr / scale, p, d, None)
            H, g, rbar, _ = ops.lm_build(lv, R, T, None, self.precision)
            lam = ops.lm_lambda(rbar, _points.shape[1], self.mlp_packed(str(level)), 1.0)
            R, T, _, _, _ = ops.lm_solve_update(H, g, lam, R, T, None, undamped_last=False,
                                                vmatrix_batch_scramble=self.vmatrix_batch_scramble)
            rotations.append(R); translations.append(T)
        return rotations, translations

    def BundleResize(self, intrisic, layers, points, basis, init_depth, init_rotation=None, init_translation=None,
                     reuse_variables=False):
        """reference bundlenet.py:332-399 -> (output_rotations, output_translations, output_depths), levels 2,3."""
        nb = layers[-1].shape[0]
        K = basis.shape[-1]
        _points, intr = self._prepare(intrisic, points)
        d = ops.resample(init_depth.detach(), _points, 0.5)                    # :341-343
        b = ops.resample(basis, _points, 0.5)                                  # :344
        p = ops.compute_coordinates(_points, intr, True)                       # :358
        dev = points.device
        R = torch.eye(3, device=dev).repeat(nb, 1, 1) if init_rotation is None else init_rotation
        T = torch.zeros(nb, 3, 1, device=dev) if init_translation is None else init_translation
        W = torch.zeros(nb, K, 1, device=dev)
        oh, ow = self.geo.out_hw
        Rs, Ts, Ds = [], [], []
        for level in range(2, 4):                                              # :376
            scale = 2 ** (3 - level)
            layer1 = ops.resample(layers[level], _points, 1.0 / scale)         # :385
            layer2 = ops.grad_fixed_concat(layers[level], swap_halves=True)    # :386-389
            lv = ops.Level(layer1, layer2, intr / scale, p, d, b)
            H, g, rbar, _ = ops.lm_build(lv, R, T, W, self.precision)
            lam = ops.lm_lambda(rbar, _points.shape[1], self.mlp_packed(str(level)), 1000.0)   # :393
            R, T, W, _, _ = ops.lm_solve_update(H, g, lam, R, T, W, undamped_last=True,
                                                vmatrix_batch_scramble=self.vmatrix_batch_scramble)
            Rs.append(R); Ts.append(T)
            depth = ops.depth_compose(init_depth.reshape(nb, -1), basis.reshape(nb, -1, K), W)   # :397
            Ds.append(depth.reshape(nb, oh, ow, 1))
        return Rs, Ts, Ds
