#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g8_*
timeout -s KILL 300 python -m pytest tests/test_gpu_tc7.py -q -m gpu --timeout 120 > gpurun_out/g8_tc7.log 2>&1; echo "tc7 rc=$?" >> gpurun_out/g8_rc.txt
for v in "" nst3nwb2 nst2nwb2 nopf; do
  if [ -z "$v" ]; then unset BANET_LIB_PATH; else export BANET_LIB_PATH=$PWD/gpurun_variants/lib_$v.so; fi
  timeout -s KILL 150 python scripts/r2_probe_tc7.py >> gpurun_out/g8_variants.log 2>&1
done
cat gpurun_out/g8_rc.txt; tail -3 gpurun_out/g8_tc7.log; cat gpurun_out/g8_variants.log
