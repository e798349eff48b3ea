"""Timing experiments on the generation-7 build kernel: variants built into gpurun_variants/lib_*.so (BANET_LIB_PATH picks one per process)."""
import os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import ops, synth, _lib
nb = int(os.environ.get("BANET_NB", "32"))
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
tag = os.environ.get("BANET_LIB_PATH", "default").split("/")[-1]
cases = [("gen6 x1", 1, dict(tc_generation=6)), ("gen7 x1 band4", 1, dict(tc_generation=7))]
if tag == "default":
    cases += [("gen7 x1 band1", 1, dict(tc_generation=7, tc7_band_rows=1)), ("gen7 x1 band2", 1, dict(tc_generation=7, tc7_band_rows=2)),
              ("gen7 x1 band8", 1, dict(tc_generation=7, tc7_band_rows=8)), ("gen7 x2 band4", 2, dict(tc_generation=7))]
for name, prec, tun in cases:
    _lib.set_tuning(**tun)
    for _ in range(2): ops.lm_build(Lf, sc.R0, sc.T0, sc.W0, precision=prec)
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): ops.lm_build(Lf, sc.R0, sc.T0, sc.W0, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 4)
    print(f"[{tag}] nb={nb} {name:16s} min {min(ts):7.3f} med {statistics.median(ts):7.3f} ms", flush=True)
