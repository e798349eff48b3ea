"""Generation-6 build kernel after a change of the per-role register budgets: 640x480, nb=32, every mode on both layouts (CUDA events)."""
import os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import ops, synth, _lib
sc = synth.make_scene(nb=32, H=480, W=640, C=128, K=128, level_ids=(3,), seed=1236, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
L3 = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
if os.environ.get("BANET_ONE"):
    for _ in range(3): ops.lm_build(L3, sc.R0, sc.T0, sc.W0, precision=1)
    torch.cuda.synchronize(); sys.exit(0)
for name, L in (("3c", L3), ("f2", Lf)):
    for prec in (1, 2, 3):
        for _ in range(2): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 4)
        print(f"{name} x{prec}: min {min(ts):.3f} med {statistics.median(ts):.3f} ms", flush=True)
