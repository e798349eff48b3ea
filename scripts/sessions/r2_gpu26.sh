#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g26_*
timeout -s KILL 300 python -m pytest tests/test_window.py tests/test_gpu_determinism.py -q -m gpu --timeout 200 -rA > gpurun_out/g26_win.log 2>&1; echo "win rc=$?" >> gpurun_out/g26_rc.txt
timeout -s KILL 300 python bench.py --config cfg4 --steps 5 > gpurun_out/g26_cfg4.json 2> gpurun_out/g26_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/g26_rc.txt
cat gpurun_out/g26_rc.txt; grep -n "passed\|failed\|^E \|window nf" gpurun_out/g26_win.log | head -30; tail -c 600 gpurun_out/g26_cfg4.err
