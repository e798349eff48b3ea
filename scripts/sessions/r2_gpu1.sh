#!/bin/bash
# round-2 GPU session 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/g1_smi.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_tc7.py -q -m gpu -s --timeout 150 > gpurun_out/g1_tc7.log 2>&1; echo "tc7 rc=$?" >> gpurun_out/g1_rc.txt
timeout -s KILL 300 python scripts/check_small_k.py > gpurun_out/g1_smallk.log 2>&1; echo "smallk rc=$?" >> gpurun_out/g1_rc.txt
PROBE=time timeout -s KILL 600 python scripts/r2_probe1.py > gpurun_out/g1_probe_time.log 2>&1; echo "probe_time rc=$?" >> gpurun_out/g1_rc.txt
PROBE=acc PROBE_TAG=r2_probe1_acc timeout -s KILL 900 python scripts/r2_probe1.py > gpurun_out/g1_probe_acc.log 2>&1; echo "probe_acc rc=$?" >> gpurun_out/g1_rc.txt
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_tc7.py > gpurun_out/g1_all.log 2>&1; echo "all rc=$?" >> gpurun_out/g1_rc.txt
cat gpurun_out/g1_rc.txt; tail -5 gpurun_out/g1_tc7.log; tail -15 gpurun_out/g1_probe_time.log
