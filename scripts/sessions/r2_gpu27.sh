#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g27_*
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g27_bench.json 2> gpurun_out/g27_bench.err; echo "bench rc=$?" >> gpurun_out/g27_rc.txt
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --motion large --no-cpu-baseline --no-e2e > gpurun_out/g27_large.json 2> gpurun_out/g27_large.err; echo "large rc=$?" >> gpurun_out/g27_rc.txt
cat gpurun_out/g27_rc.txt; tail -c 500 gpurun_out/g27_bench.err; tail -c 500 gpurun_out/g27_large.err; ls /sys/devices/system/node/ | head; nvidia-smi topo -m | head -12
