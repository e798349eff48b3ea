#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g10_*
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lm_ -c 400 --csv --log-file gpurun_out/g10_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-check > gpurun_out/g10_b.log 2>&1
echo rc=$?; wc -l gpurun_out/g10_launches.csv
