#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g5_*
timeout -s KILL 300 python -m pytest tests/test_gpu_tc7.py -q -m gpu -s --timeout 120 > gpurun_out/g5_tc7.log 2>&1; echo "tc7 rc=$?" >> gpurun_out/g5_rc.txt
unset BANET_LIB_PATH
timeout -s KILL 150 python scripts/r2_probe_tc7.py >> gpurun_out/g5_variants.log 2>&1
timeout -s KILL 400 python -m pytest tests/test_gpu_legacy.py tests/test_host_pipeline.py "tests/test_gpu_parity.py::test_lm_build_matches_oracle" tests/test_gpu_tensorcore.py -q -m gpu -s --timeout 150 > gpurun_out/g5_misc.log 2>&1; echo "misc rc=$?" >> gpurun_out/g5_rc.txt
cat gpurun_out/g5_rc.txt; tail -4 gpurun_out/g5_tc7.log; cat gpurun_out/g5_variants.log; tail -12 gpurun_out/g5_misc.log
