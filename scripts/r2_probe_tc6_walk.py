"""Generation-6 build kernel, 640x480, nb=32: tile walk (bands) x L2 eviction hints.  Times with CUDA events; BANET_ONE=band,hints runs
that single configuration three times (for an ncu dram-bytes pass)."""
import os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import ops, synth, _lib
nb = int(os.environ.get("BANET_NB", "32"))
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=1236, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
L3 = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
one = os.environ.get("BANET_ONE")
if one:
    band, hints, pf, f2 = (int(x) for x in one.split(","))
    _lib.set_tuning(tc6_band_rows=band, tc6_l2_hints=hints, tc6_tap_prefetch=pf)
    if f2: L3 = Lf
    for _ in range(3): ops.lm_build(L3, sc.R0, sc.T0, sc.W0, precision=1)
    torch.cuda.synchronize(); sys.exit(0)
def run(name, L, prec, tun, ref=None):
    _lib.set_tuning(**tun)
    for _ in range(2): out = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): out = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 4)
    H = out[0].double()
    rel = float((H - ref).norm() / ref.norm()) if ref is not None else 0.0
    print(f"nb={nb} {name:28s} min {min(ts):7.3f} med {statistics.median(ts):7.3f} ms  relH-vs-first {rel:.2e}", flush=True)
    return H
T = lambda pf, hints=1, band=1: dict(tc6_band_rows=band, tc6_l2_hints=hints, tc6_tap_prefetch=pf)
ref = run("3c x1 pf-off", L3, 1, T(1))
run("3c x1 pf-edge", L3, 1, T(2), ref)
run("3c x1 pf-edge+row", L3, 1, T(3), ref)
run("3c x1 pf-edge hints2", L3, 1, T(2, 2), ref)
run("3c x1 pf-edge hints3", L3, 1, T(2, 3), ref)
run("3c x1 pf-edge+row hints2", L3, 1, T(3, 2), ref)
run("3c x1 pf-off (again)", L3, 1, T(1), ref)
ref2 = run("3c x2 pf-off", L3, 2, T(1))
run("3c x2 pf-edge", L3, 2, T(2), ref2)
ref3 = run("3c x3 pf-off", L3, 3, T(1))
run("3c x3 pf-edge", L3, 3, T(2), ref3)
reff = run("f2 x1 pf-off", Lf, 1, T(1))
run("f2 x1 pf-edge", Lf, 1, T(2), reff)
run("f2 x1 pf-edge+row", Lf, 1, T(3), reff)
run("f2 x1 pf-edge hints2", Lf, 1, T(2, 2), reff)
reff2 = run("f2 x2 pf-off", Lf, 2, T(1))
run("f2 x2 pf-edge", Lf, 2, T(2), reff2)
