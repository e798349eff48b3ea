#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g3_*
for v in "" notma nogather nogather_notma nwb2 boxy4 boxy1; do
  if [ -z "$v" ]; then unset BANET_LIB_PATH; else export BANET_LIB_PATH=$PWD/gpurun_variants/lib_$v.so; fi
  timeout -s KILL 120 python scripts/r2_probe_tc7.py >> gpurun_out/g3_variants.log 2>&1
done
unset BANET_LIB_PATH
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from banet_b200 import ops, synth, _lib
sc = synth.make_scene(nb=4, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
Lf = ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
_lib.set_tuning(tc_generation=7)
for _ in range(3): ops.lm_build(Lf, sc.R0, sc.T0, sc.W0, precision=1)
torch.cuda.synchronize()
PY
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:lm_build_tc7 --launch-skip 2 --launch-count 1 -f -o gpurun_out/g3_tc7_x1 python /tmp/one.py > gpurun_out/g3_ncu.log 2>&1
ls -la gpurun_out/g3_*; cat gpurun_out/g3_variants.log
