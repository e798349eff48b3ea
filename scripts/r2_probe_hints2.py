"""After the evict-last policy on the partial slots: do the stream / tap L2 hints pay now?  640x480, nb=32, TF32X1 (CUDA events)."""
import os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import ops, synth, _lib
sc = synth.make_scene(nb=32, H=480, W=640, C=128, K=128, level_ids=(3,), seed=1236, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
L3 = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
def run(name, L, prec, tun):
    _lib.set_tuning(**tun)
    for _ in range(2): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 4)
    print(f"{name:24s} min {min(ts):.3f} med {statistics.median(ts):.3f} ms", flush=True)
for rep in range(2):
    for h in (1, 2, 3):
        run(f"3c x1 hints{h}", L3, 1, dict(tc6_l2_hints=h))
for h in (1, 2):
    run(f"f2 x1 gen6 hints{h}", Lf, 1, dict(tc6_l2_hints=h))
run("f2 x1 gen7", Lf, 1, dict(tc_generation=7))
run("f2 x2 gen7", Lf, 2, dict(tc_generation=7))
run("3c x3 hints1", L3, 3, dict(tc6_l2_hints=1)); run("3c x3 hints2", L3, 3, dict(tc6_l2_hints=2))
