"""Host-side mirror of the reference's legacy pose-only tracker (reference legacy/ba.py:15-145): same module-level knobs, same
`Tracker.trackTF` signature and return values, backed by banet_lm_track_legacy (accept / reject and early termination on the device).

Differences a reference user should know: the CNN feature extractor (`feat.DRN`, `Pyramid`, legacy/ba.py:455-459) is out of scope, so
`Tracker` is constructed from the number of feature channels instead of a checkpoint and `trackTF` is the entry point; any batch size
works (the reference runs one keyframe/frame pair); with early termination every pair stops on its own."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops

Tensor = torch.Tensor

# legacy/ba.py:5-9
early_termination = True
angle_change = 0.002 * (3.14 / 180.0)
translation_change = 0.0002
residual_ratio = 1.0


class Tracker(torch.nn.Module):
    """Drop-in for the BA part of reference `Tracker` (legacy/ba.py:15).  `channels` = feature channels of the pyramid."""

    def __init__(self, channels: int, levels: Sequence[str] = ("1", "2", "3"), seed: int = 7):
        super().__init__()
        self.channels = channels
        g = torch.Generator().manual_seed(seed)
        dims = [channels, 2 * channels, 4 * channels, 2 * channels, channels, 1]
        for lv in levels:
            for i in range(5):      # he_normal filters, zero biases (legacy/ba.py:219-220)
                self.register_parameter(f"lambda_{lv}_{i + 1}_filters", torch.nn.Parameter(torch.randn(dims[i], dims[i + 1], generator=g) * math.sqrt(2.0 / dims[i])))
                self.register_parameter(f"lambda_{lv}_{i + 1}_biases", torch.nn.Parameter(torch.zeros(dims[i + 1])))
        self.last_iters_done: Optional[Tensor] = None
        self.last_status: Optional[Tensor] = None

    def mlp_packed(self, level: str) -> Tensor:
        return ops.pack_mlp([(getattr(self, f"lambda_{level}_{i}_filters").detach(), getattr(self, f"lambda_{level}_{i}_biases").detach()) for i in range(1, 6)])

    def grad_fixed(self, input: Tensor, name=None) -> Tensor:
        """legacy/ba.py:17-25."""
        return ops.grad_fixed_concat(input)[..., input.shape[-1]:].contiguous()

    def computeCoordinates(self, points2d: Tensor, fx, fy, ox, oy) -> Tensor:
        """legacy/ba.py:27-34 (un-normalised rays)."""
        intr = torch.stack([t.reshape(t.shape[0], -1)[:, 0] for t in (fx, fy, ox, oy)], dim=1).float().contiguous()
        return ops.compute_coordinates(points2d, intr, normalize=False)

    @torch.no_grad()
    def trackTF(self, intrisic: Tensor, layers: Sequence[Tensor], points: Tensor, d: Tensor, initR: Tensor, initT: Tensor, level_iters: Sequence[int]):
        """legacy/ba.py:83-145.  layers[l] [2*nb,h,w,C]: first half keyframes, second half current frames (the reference uses nb = 1: images 0 and 1);
        points [nb,N,2] keyframe pixels at the finest level, d [nb,N,1] their depths, intrisic [nb,4,1].
        Returns (R, T, ratio) with early termination, else (rotations, translations, ratio) per iteration like the reference."""
        nb = points.shape[0]
        k = intrisic.reshape(nb, 4).float()
        p = ops.compute_coordinates(points, k.contiguous(), normalize=False)
        levels, mlps = [], []
        for level in range(1, 4):
            scale = 2 ** (3 - level)
            lay = layers[level - 1]
            layer1 = ops.interpolate2d(lay[0:nb].contiguous(), points, 1.0 / scale)        # utils.interpolate2d2(layers[level-1][0:1], points/scale), :112
            layer2 = ops.grad_fixed_concat(lay[nb:2 * nb].contiguous())                     # :113-115
            levels.append(ops.Level(layer1, layer2, (k / scale).contiguous(), p, d, None))
            mlps.append(self.mlp_packed(str(level)))
        if early_termination:
            R, T, done, ratio, status = ops.lm_track_legacy(levels, level_iters, initR, initT, mlps, True, angle_change, translation_change, residual_ratio)
            self.last_iters_done, self.last_status = done, status
            return R, T, ratio
        rotations, translations, R, T = [], [], initR, initT
        ratio = None
        for lv, n in zip(levels, level_iters):
            for _ in range(int(n)):
                R, T, _, ratio, status = ops.lm_track_legacy([lv], [1], R, T, None, False)
                rotations.append(R); translations.append(T)
        self.last_status = status
        return rotations, translations, ratio
