#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g18_*
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/g18_bench.json 2> gpurun_out/g18_bench.err; echo "bench rc=$?" >> gpurun_out/g18_rc.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lm_ -c 400 --csv --log-file gpurun_out/g18_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-check > gpurun_out/g18_b.log 2>&1; echo "launches rc=$?" >> gpurun_out/g18_rc.txt
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
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:lm_build_tc6 --launch-skip 2 --launch-count 1 -f -o gpurun_out/g18_tc6_x1_3c python /tmp/one.py > gpurun_out/g18_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/g18_rc.txt
timeout -s KILL 400 python bench.py --config cfg5 --steps 3 > gpurun_out/g18_cfg5.json 2> gpurun_out/g18_cfg5.err; echo "cfg5 rc=$?" >> gpurun_out/g18_rc.txt
timeout -s KILL 300 python bench.py --config cfg4 --steps 5 > gpurun_out/g18_cfg4.json 2> gpurun_out/g18_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/g18_rc.txt
cat gpurun_out/g18_rc.txt; ls -la gpurun_out/g18_*; tail -c 600 gpurun_out/g18_cfg5.err; tail -c 600 gpurun_out/g18_cfg4.err
