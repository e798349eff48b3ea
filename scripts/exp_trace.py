"""Debug (GPU): per-tile event timeline of two gather warps of CTA 1 (set via BANET_TC_TRACE_PTR)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
buf = torch.zeros(2 * 32 * 8, dtype=torch.int64, device="cuda")
os.environ["BANET_TC_TRACE_PTR"] = str(buf.data_ptr())
from banet_b200 import ops, synth
nb = 8
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
prec = int(os.environ.get("BANET_PREC", "1"))
mode = os.environ.get("BANET_CASE", "normal")
T = sc.T0.clone()
if mode == "masked": T[:, 0, 0] = 1e4
L = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
for _ in range(3): ops.lm_build(L, sc.R0, T, sc.W0, precision=prec)
torch.cuda.synchronize()
tr = buf.cpu().reshape(2, 32, 8)
for wsel in (0, 1):
    print(f"--- gather warp {'first' if wsel == 0 else 'last'} ({mode}, prec={prec}): per-tile phase durations in ns")
    print("tile  " + "  ".join(f"{n:>11s}" for n in ["pre", "wait fullB", "dots+geom", "gather", "S3", "wait rfree", "scale+arr", "tile total"]))
    for i in range(4, 24):
        e = tr[wsel, i].tolist(); nxt = tr[wsel, i + 1, 0].item()
        d = [e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6] - e[5], e[7] - e[6], nxt - e[0]]
        print(f"{i+16:4d}  " + "  ".join(f"{x:11d}" for x in d))
