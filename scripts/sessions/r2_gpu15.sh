#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g16_*
( time timeout -s KILL 700 python -m pytest tests -q -m gpu --timeout 300 --durations=8 ) > gpurun_out/g16_all.log 2>&1; echo "all rc=$?" >> gpurun_out/g16_rc.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g16_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g16_rc.txt
cat gpurun_out/g16_rc.txt; tail -25 gpurun_out/g16_all.log; tail -3 gpurun_out/g16_smoke.log
