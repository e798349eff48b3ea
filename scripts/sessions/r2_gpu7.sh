#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g7_*
for v in "" nst2nwb3 nodep w12 nopf nst2; do
  if [ -z "$v" ]; then unset BANET_LIB_PATH; else export BANET_LIB_PATH=$PWD/gpurun_variants/lib_$v.so; fi
  timeout -s KILL 150 python scripts/r2_probe_tc7.py >> gpurun_out/g7_variants.log 2>&1
done
cat gpurun_out/g7_variants.log
