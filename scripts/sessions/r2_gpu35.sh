#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g35_*
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lm_ -c 400 --csv --log-file gpurun_out/g35_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-check > gpurun_out/g35_b.log 2>&1; echo "launches rc=$?"
