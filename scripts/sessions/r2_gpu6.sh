#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g6_*
cat > /tmp/one.py <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
from banet_b200 import ops, synth, _lib
gen = int(os.environ.get("GEN", "7"))
sc = synth.make_scene(nb=4, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
_lib.set_tuning(tc_generation=gen)
for _ in range(3): ops.lm_build(Lf, sc.R0, sc.T0, sc.W0, precision=1)
torch.cuda.synchronize()
PY
GEN=7 timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:lm_build_tc7 --launch-skip 2 --launch-count 1 -f -o gpurun_out/g6_tc7_x1 python /tmp/one.py > gpurun_out/g6_ncu7.log 2>&1
GEN=6 timeout -s KILL 400 ncu --section SpeedOfLight --section InstructionStats --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis --clock-control none -k regex:lm_build_tc6 --launch-skip 2 --launch-count 1 -f -o gpurun_out/g6_tc6_x1 python /tmp/one.py > gpurun_out/g6_ncu6.log 2>&1
ls -la gpurun_out/g6_*
