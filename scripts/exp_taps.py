"""Experiment (GPU): how much of the build time is conv2-tap memory latency?  Compare the normal scene with one whose
intrinsics make every point sample the same texel (all tap loads hit L1), for SIMT and tensor-core paths."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from banet_b200 import ops, synth
nb = 8
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
f2 = lv.conv2[..., :128].contiguous()
intr0 = lv.intr.clone(); intr0[:, 0] = 0; intr0[:, 1] = 0          # u = ox, v = oy for every point
def run(tag, L, prec):
    for _ in range(3): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag:40s} prec={prec}: {e0.elapsed_time(e1)/5:7.3f} ms  nvalid={out[3].sum().item():.0f}", flush=True)
for prec in (1, 3):
    for fly in (0, 1):
        c2 = f2 if fly else lv.conv2
        run(f"normal fly={fly} grid", ops.Level(lv.conv1, c2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid), prec)
        run(f"same-texel fly={fly} grid", ops.Level(lv.conv1, c2, intr0, lv.p, lv.D, lv.B, grid=lv.grid), prec)
T_far = sc.T0.clone(); T_far[:, 0, 0] = 1e4
L = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
for prec in (1, 3):
    for _ in range(3): ops.lm_build(L, sc.R0, T_far, sc.W0, precision=prec)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(5): out = ops.lm_build(L, sc.R0, T_far, sc.W0, precision=prec)
    e1.record(); torch.cuda.synchronize()
    print(f"all-masked (no taps at all)              prec={prec}: {e0.elapsed_time(e1)/5:7.3f} ms  nvalid={out[3].sum().item():.0f}", flush=True)
