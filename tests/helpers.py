"""Shared test helpers: small seeded cases in the reference's layouts, oracle <-> torch plumbing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from banet_b200 import synth            # noqa: E402
from oracle import ba_oracle as O       # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if GOLDEN_DIR not in sys.path:
    sys.path.insert(0, GOLDEN_DIR)


def rel_fro(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu(); b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def scene_case(nb=2, H=48, W=64, C=8, K=4, level_ids=(2, 3), seed=11, n_points=None, dtype=torch.float64, **kw):
    """A planted-solution scene on CPU (float64 by default, for the oracle)."""
    return synth.make_scene(nb=nb, H=H, W=W, C=C, K=K, level_ids=level_ids, seed=seed, n_points=n_points,
                            dtype=dtype, device="cpu", **kw)


def oracle_level_inputs(lv, dtype=torch.float64):
    fx, fy, ox, oy = [t.to(dtype) for t in lv.intr_tiled()]
    return dict(conv1=lv.conv1.to(dtype), conv2=lv.conv2.to(dtype), fx=fx, fy=fy, ox=ox, oy=oy, p=lv.p.to(dtype),
                D=lv.D.to(dtype), B=None if lv.B is None else lv.B.to(dtype))


def mlp_for(C, level, dtype=torch.float64):
    return O.init_lambda_mlp(C, seed=100 + int(level), dtype=dtype)


def to_cuda32(t):
    return None if t is None else t.to(device="cuda", dtype=torch.float32).contiguous()
