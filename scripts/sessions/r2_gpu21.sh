#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g21_*
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g21_bench2.json 2> gpurun_out/g21_bench2.err; echo "bench2 rc=$?" >> gpurun_out/g21_rc.txt
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/g21_ref2.json 2> gpurun_out/g21_ref2.err; echo "ref2 rc=$?" >> gpurun_out/g21_rc.txt
cat gpurun_out/g21_rc.txt; tail -c 1500 gpurun_out/g21_bench2.json; tail -c 800 gpurun_out/g21_bench2.err; tail -c 600 gpurun_out/g21_ref2.json; tail -c 400 gpurun_out/g21_ref2.err
