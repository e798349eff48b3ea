#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g12_*
timeout -s KILL 300 python -m pytest "tests/test_gpu_parity.py" tests/test_gpu_tc7.py -q -m gpu --timeout 120 > gpurun_out/g12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g12_rc.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lm_ -c 300 --csv --log-file gpurun_out/g12_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-check > gpurun_out/g12_b.log 2>&1
cat gpurun_out/g12_rc.txt; tail -3 gpurun_out/g12_tests.log
