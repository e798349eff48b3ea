#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g32_*
timeout -s KILL 400 python -m pytest tests/test_gpu_default_precision.py tests/test_gpu_determinism.py -q -m gpu --timeout 300 -rA > gpurun_out/g32_prec.log 2>&1; echo "prec rc=$?" >> gpurun_out/g32_rc.txt
timeout -s KILL 200 python bench.py --layout f2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g32_f2.json 2> gpurun_out/g32_f2.err; echo "f2 rc=$?" >> gpurun_out/g32_rc.txt
cat gpurun_out/g32_rc.txt; grep -n "passed\|failed\|auto:\|levelwise:\|tf32x1:" gpurun_out/g32_prec.log | head; tail -c 300 gpurun_out/g32_f2.err
