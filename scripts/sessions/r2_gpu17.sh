#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g17_*
timeout -s KILL 330 python -X faulthandler -m pytest tests/test_gpu_autograd.py tests/test_gpu_default_precision.py tests/test_gpu_legacy.py tests/test_gpu_parity.py -v -m gpu -o faulthandler_timeout=45 --durations=10 > gpurun_out/g17_seq.log 2>&1; echo "seq rc=$?" >> gpurun_out/g17_rc.txt
cat gpurun_out/g17_rc.txt; grep -n "PASSED\|FAILED\|ERROR" gpurun_out/g17_seq.log | tail -8; grep -n "Timeout\|File \"" gpurun_out/g17_seq.log | head -30
