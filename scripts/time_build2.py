"""Interleaved, repeated timing of lm_build at 640x480, C=K=128 (GPU): for each case min / median ms over R rounds.
cases: BANET_CASES="prec:fly:grid,..." (default 1:0:1,2:0:1,2:1:1,3:0:1); BANET_NB pairs; BANET_ROUNDS rounds."""
import os, sys, statistics, subprocess, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from banet_b200 import ops, synth
nb = int(os.environ.get("BANET_NB", "32")); rounds = int(os.environ.get("BANET_ROUNDS", "4"))
cases = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("BANET_CASES", "1:0:1,2:0:1,2:1:1,3:0:1").split(",")]
sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
f2 = lv.conv2[..., :128].contiguous()
levels = {}
for prec, fly, grid in cases:
    levels[(prec, fly, grid)] = ops.Level(lv.conv1, f2 if fly else lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid if grid else None)
res = {c: [] for c in cases}
for c in cases:
    for _ in range(2): ops.lm_build(levels[c], sc.R0, sc.T0, sc.W0, precision=c[0])
torch.cuda.synchronize()
for r in range(rounds):
    for c in cases:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ops.lm_build(levels[c], sc.R0, sc.T0, sc.W0, precision=c[0])
        e0.record()
        for _ in range(4): ops.lm_build(levels[c], sc.R0, sc.T0, sc.W0, precision=c[0])
        e1.record(); torch.cuda.synchronize()
        res[c].append(e0.elapsed_time(e1) / 4)
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
except Exception as e:
    q = str(e)
print(f"# gen={os.environ.get('BANET_TC_GEN', 'default')} nb={nb} rounds={rounds} smi(after): {q}")
for c in cases:
    v = res[c]
    print(f"prec={c[0]} fly={c[1]} grid={c[2]}: min {min(v):7.3f}  med {statistics.median(v):7.3f} ms   all {' '.join(f'{x:.2f}' for x in v)}")
