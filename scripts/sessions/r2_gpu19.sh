#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g19_*
timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc or build" > gpurun_out/g19_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/g19_rc.txt
timeout -s KILL 240 python scripts/r2_probe_tc6_walk.py > gpurun_out/g19_walk.log 2>&1; echo "walk rc=$?" >> gpurun_out/g19_rc.txt
for one in 1,1 4,1 4,2 4,3; do
BANET_ONE=$one timeout -s KILL 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:lm_build_tc6 --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/g19_dram_${one/,/_}.csv python scripts/r2_probe_tc6_walk.py > /dev/null 2>&1; echo "ncu $one rc=$?" >> gpurun_out/g19_rc.txt
done
cat gpurun_out/g19_rc.txt; tail -5 gpurun_out/g19_parity.log; cat gpurun_out/g19_walk.log; grep -h "dram__\|lts__\|gpu__time" gpurun_out/g19_dram_*.csv | cut -d, -f5,13-
