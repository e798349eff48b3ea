#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g24_*
timeout -s KILL 500 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_parity.py tests/test_gpu_legacy.py tests/test_gpu_default_precision.py -q -m gpu --timeout 200 -rA > gpurun_out/g24_det.log 2>&1; echo "det rc=$?" >> gpurun_out/g24_rc.txt
cat gpurun_out/g24_rc.txt; grep -n "passed\|failed\|Error\|assert\|^E " gpurun_out/g24_det.log | head -40
