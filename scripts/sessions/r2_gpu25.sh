#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g25_*
timeout -s KILL 300 python -m pytest tests/test_gpu_determinism.py -q -m gpu --timeout 200 -rA > gpurun_out/g25_det.log 2>&1; echo "det rc=$?" >> gpurun_out/g25_rc.txt
timeout -s KILL 300 python bench.py --config cfg4 --steps 5 > gpurun_out/g25_cfg4.json 2> gpurun_out/g25_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/g25_rc.txt
timeout -s KILL 300 python bench.py --layout f2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g25_f2.json 2> gpurun_out/g25_f2.err; echo "f2 rc=$?" >> gpurun_out/g25_rc.txt
cat gpurun_out/g25_rc.txt; grep -n "passed\|failed\|^E " gpurun_out/g25_det.log | head -20; tail -c 400 gpurun_out/g25_cfg4.err; tail -c 300 gpurun_out/g25_f2.err
