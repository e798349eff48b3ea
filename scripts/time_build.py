"""Timing of lm_build at 640x480, C=K=128 for every precision / layout / tiling combination (GPU)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from banet_b200 import ops, synth
nb = int(os.environ.get("BANET_NB", "8"))
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
f2 = lv.conv2[..., :128].contiguous()
precs = [int(x) for x in os.environ.get("BANET_PRECS", "0,1,2,3").split(",")]
for prec in precs:
    for fly in (False, True):
        for grid in ((None, lv.grid) if prec else (None,)):
            L = ops.Level(lv.conv1, f2 if fly else lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=grid)
            for _ in range(3): out = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): out = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            by = nb * (4 * lv.N * (2 * 128 + 128 + 4))
            print(f"prec={prec} fly={int(fly)} grid={int(grid is not None)}: {ms:7.3f} ms / {nb} pairs @640x480 -> {by/ms/1e6:6.0f} GB/s alg ({by/ms/1e6/6567.1:.3f} of peak)", flush=True)
