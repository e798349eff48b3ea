#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g34_*
( time timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 400 --durations=6 ) > gpurun_out/g34_all.log 2>&1; echo "all rc=$?" >> gpurun_out/g34_rc.txt
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g34_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g34_rc.txt
timeout -s KILL 400 python bench.py > gpurun_out/g34_bench.json 2> gpurun_out/g34_bench.err; echo "bench rc=$?" >> gpurun_out/g34_rc.txt
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from banet_b200 import ops, synth, _lib
sc = synth.make_scene(nb=32, H=480, W=640, C=128, K=128, level_ids=(3,), seed=1236, device="cuda", dtype=torch.float32)
lv = sc.levels[0]
L = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
for _ in range(3): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=_lib.PREC_AUTO)
torch.cuda.synchronize()
PY
timeout -s KILL 200 ncu --set full --import-source on --clock-control none -k regex:lm_build_tc6 --launch-skip 2 --launch-count 1 -f -o gpurun_out/g34_tc6_x1_3c python /tmp/one.py > gpurun_out/g34_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/g34_rc.txt
cat gpurun_out/g34_rc.txt; grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/g34_all.log | tail -8; tail -2 gpurun_out/g34_smoke.log; tail -c 300 gpurun_out/g34_bench.err
