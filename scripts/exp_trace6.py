"""Debug (GPU): per-tile timeline of gather warp 0, algebra warp 0 and geometry warp 0 of CTA 1 in the generation-6 build
kernel (needs a build with `make EXTRA=-DBANET_TC6_TRACE_ON`)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
buf = torch.zeros(3 * 32 * 12, dtype=torch.int64, device="cuda")
os.environ["BANET_TC_TRACE_PTR"] = str(buf.data_ptr())
from banet_b200 import ops, synth
nb = int(os.environ.get("BANET_NB", "8"))
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
prec = int(os.environ.get("BANET_PREC", "2"))
fly = int(os.environ.get("BANET_FLY", "0"))
L = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous() if fly else lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
for _ in range(3): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
torch.cuda.synchronize()
tr = buf.cpu().reshape(3, 32, 12)
t0 = tr[0, 2, 0].item()
print(f"--- prec={prec} fly={fly}: ns; columns are phase durations, 'start' is relative to gather tile 18")
print("GATHER  tile    start  wait recs      units   tile total")
for i in range(2, 30):
    e = tr[0, i].tolist(); nxt = tr[0, i + 1, 0].item()
    print(f"       {i+16:4d} {e[0]-t0:8d} {e[1]-e[0]:10d} {e[2]-e[1]:10d} {nxt-e[0]:10d}")
print("ALGEBRA tile    start  wait gath        S3  wait rfree     scale   bar+issue  iter total")
for i in range(2, 30):
    e = tr[1, i].tolist(); nxt = tr[1, i + 1, 0].item()
    print(f"       {i+16:4d} {e[0]-t0:8d} {e[1]-e[0]:10d} {e[2]-e[1]:10d} {e[3]-e[2]:10d} {e[4]-e[3]:10d} {e[5]-e[4]:10d} {nxt-e[0]:10d}")
print("GEOM    tile    start  wait(recfree,fullB)   dots   geom   iter total")
for i in range(2, 30):
    e = tr[2, i].tolist(); nxt = tr[2, i + 1, 0].item()
    print(f"       {i+16:4d} {e[0]-t0:8d} {e[1]-e[0]:10d} {e[2]-e[1]:10d} {e[3]-e[2]:10d} {nxt-e[0]:10d}")
