#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g30_*
timeout -s KILL 200 python -m pytest tests/test_gpu_tc7.py -x -q -m gpu --timeout 150 > gpurun_out/g30_tc7.log 2>&1; echo "tc7 rc=$?" >> gpurun_out/g30_rc.txt
timeout -s KILL 240 python scripts/r2_probe_hints2.py > gpurun_out/g30_hints.log 2>&1; echo "hints rc=$?" >> gpurun_out/g30_rc.txt
cat gpurun_out/g30_rc.txt; tail -2 gpurun_out/g30_tc7.log; cat gpurun_out/g30_hints.log
