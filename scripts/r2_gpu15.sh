#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g15_*
( time timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 900 --durations=8 ) > gpurun_out/g15_all.log 2>&1; echo "all rc=$?" >> gpurun_out/g15_rc.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g15_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g15_rc.txt
cat gpurun_out/g15_rc.txt; tail -25 gpurun_out/g15_all.log; tail -3 gpurun_out/g15_smoke.log
