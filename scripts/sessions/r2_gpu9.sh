#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g9_*
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/g9_bench_concat.json 2> gpurun_out/g9_bench_concat.err; echo "bench concat rc=$?" >> gpurun_out/g9_rc.txt
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --layout f2 --no-e2e --no-cpu-baseline > gpurun_out/g9_bench_f2.json 2> gpurun_out/g9_bench_f2.err; echo "bench f2 rc=$?" >> gpurun_out/g9_rc.txt
timeout -s KILL 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/g9_bench_ref.json 2> gpurun_out/g9_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/g9_rc.txt
cat gpurun_out/g9_rc.txt; tail -c 1500 gpurun_out/g9_bench_concat.err; head -c 600 gpurun_out/g9_bench_ref.json
