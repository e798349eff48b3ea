#!/bin/bash
# round-2 GPU session 2
mkdir -p gpurun_out; rm -f gpurun_out/g2_*
timeout -s KILL 300 python -m pytest tests/test_gpu_tc7.py -q -m gpu -s --timeout 120 > gpurun_out/g2_tc7.log 2>&1; echo "tc7 rc=$?" >> gpurun_out/g2_rc.txt
timeout -s KILL 420 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_reference_pin.py tests/test_host_pipeline.py tests/test_gpu_parity.py -q -m gpu -s --timeout 120 > gpurun_out/g2_bwd.log 2>&1; echo "bwd rc=$?" >> gpurun_out/g2_rc.txt
timeout -s KILL 120 python tests/golden/gen_ref_eqc_golden.py > gpurun_out/g2_refgold.log 2>&1; echo "refgold rc=$?" >> gpurun_out/g2_rc.txt
PROBE=time PROBE_TAG=r2_probe2_time timeout -s KILL 300 python scripts/r2_probe1.py > gpurun_out/g2_probe_time.log 2>&1; echo "probe_time rc=$?" >> gpurun_out/g2_rc.txt
for cfg in "levelwise concat" "levelwise f2" "tf32x2 f2" "tf32x1 f2" "tf32x1 concat"; do
  set -- $cfg
  timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --precision $1 --layout $2 > gpurun_out/g2_bench_$1_$2.json 2> gpurun_out/g2_bench_$1_$2.err; echo "bench $1 $2 rc=$?" >> gpurun_out/g2_rc.txt
done
cat gpurun_out/g2_rc.txt; tail -3 gpurun_out/g2_tc7.log; tail -8 gpurun_out/g2_bwd.log; tail -14 gpurun_out/g2_probe_time.log
