"""Profiling driver (run under ncu on the GPU box): a few lm_build launches at 640x480, C=K=128."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from banet_b200 import ops, synth
prec = int(os.environ.get("BANET_PREC", "2")); fly = int(os.environ.get("BANET_FLY", "0")); nb = int(os.environ.get("BANET_NB", "8"))
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
conv2 = lv.conv2[..., :128].contiguous() if fly else lv.conv2
L = ops.Level(lv.conv1, conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid if int(os.environ.get("BANET_GRID", "1")) else None)
for _ in range(int(os.environ.get("BANET_REPS", "3"))):
    ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
torch.cuda.synchronize()
