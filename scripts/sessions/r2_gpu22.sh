#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g22_*
( time timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 400 --durations=10 -rA ) > gpurun_out/g22_all.log 2>&1; echo "all rc=$?" >> gpurun_out/g22_rc.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g22_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g22_rc.txt
timeout -s KILL 200 python scripts/r2_time_training_step.py > gpurun_out/g22_train.log 2>&1; echo "train rc=$?" >> gpurun_out/g22_rc.txt
cat gpurun_out/g22_rc.txt; grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/g22_all.log | tail -12; tail -3 gpurun_out/g22_smoke.log; tail -6 gpurun_out/g22_train.log
