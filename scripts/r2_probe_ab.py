"""Interleaved A/B (the GPU's power state drifts by several percent within a session, so only adjacent measurements compare):
F2-only layout, TF32X1, generation 6 vs generation 7, at 640x480 (nb=32) and 320x240 (nb=32)."""
import os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import ops, synth, _lib
for H, W in ((480, 640), (240, 320)):
    sc = synth.make_scene(nb=32, H=H, W=W, C=128, K=128, level_ids=(3,), seed=1236, device="cuda", dtype=torch.float32)
    lv = sc.levels[0]
    Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
    L3 = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
    def t(L, gen):
        _lib.set_tuning(tc_generation=gen)
        ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=1)
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=1)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 4)
        return statistics.median(ts)
    for _ in range(3): t(L3, 6)                  # bring the board to its steady power state first
    for r in range(4):
        a, b, c = t(Lf, 6), t(Lf, 7), t(L3, 6)
        print(f"{W}x{H} round {r}: f2 gen6 {a:.3f}  f2 gen7 {b:.3f}  3c gen6 {c:.3f} ms", flush=True)
    del sc, lv, Lf, L3
    torch.cuda.empty_cache()
