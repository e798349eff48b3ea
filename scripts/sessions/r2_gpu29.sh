#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g29_*
timeout -s KILL 200 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_determinism.py -x -q -m gpu --timeout 150 > gpurun_out/g29_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/g29_rc.txt
timeout -s KILL 200 python scripts/r2_probe_regs.py > gpurun_out/g29_regs.log 2>&1; echo "regs rc=$?" >> gpurun_out/g29_rc.txt
BANET_ONE=1 timeout -s KILL 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:lm_build_tc6 --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/g29_dram.csv python scripts/r2_probe_regs.py > /dev/null 2>&1; echo "ncu rc=$?" >> gpurun_out/g29_rc.txt
cat gpurun_out/g29_rc.txt; tail -3 gpurun_out/g29_tc.log; cat gpurun_out/g29_regs.log; grep -h "dram__\|lts__\|gpu__time" gpurun_out/g29_dram.csv | cut -d, -f13-
