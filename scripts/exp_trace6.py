"""Debug (GPU): per-tile timeline of gather warp 0 and helper warp 0 of CTA 1 in the generation-6 build kernel."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
buf = torch.zeros(2 * 32 * 12, dtype=torch.int64, device="cuda")
os.environ["BANET_TC_TRACE_PTR"] = str(buf.data_ptr())
from banet_b200 import ops, synth
nb = 8
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
prec = int(os.environ.get("BANET_PREC", "2"))
L = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
for _ in range(3): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
torch.cuda.synchronize()
tr = buf.cpu().reshape(2, 32, 12)
print(f"--- gather warp 0 (prec={prec}): ns")
print("tile   wait recs      units   tile total")
for i in range(2, 30):
    e = tr[0, i].tolist(); nxt = tr[0, i + 1, 0].item()
    print(f"{i+16:4d} {e[1]-e[0]:10d} {e[2]-e[1]:10d} {nxt-e[0]:10d}")
print("--- helper warp 0: ns  (geometry of tile j+1, then algebra / R rows of tile j)")
print("tile  " + " ".join(f"{n:>10s}" for n in ["wait fullB", "dots", "geom", "drain?", "wait gath", "S3", "wait rfree", "scale", "fence", "iter total"]))
for i in range(2, 30):
    e = tr[1, i].tolist(); nxt = tr[1, i + 1, 0].item()
    d = [e[1]-e[0], e[2]-e[1], e[3]-e[2], e[4]-e[3], e[5]-e[4], e[6]-e[5], e[7]-e[6], e[8]-e[7], e[9]-e[8], nxt-e[0]]
    # scale = e7 - e6 includes the rfree wait; split not traced separately
    print(f"{i+16:4d}  " + " ".join(f"{x:10d}" for x in d))
