#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g31_*
timeout -s KILL 240 python scripts/r2_probe_ab.py > gpurun_out/g31_ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/g31_rc.txt
cat gpurun_out/g31_rc.txt; cat gpurun_out/g31_ab.log
