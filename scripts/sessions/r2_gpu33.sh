#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g33_*
for g in 6 7 6 7; do
timeout -s KILL 120 python bench.py --layout f2 --tc-generation $g --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-check >> gpurun_out/g33_f2_gen$g.json 2>> gpurun_out/g33_err.txt; echo "gen$g rc=$?" >> gpurun_out/g33_rc.txt
done
cat gpurun_out/g33_rc.txt; python - <<'PY'
import json
for g in (6,7):
    for l in open(f"gpurun_out/g33_f2_gen{g}.json"):
        d=json.loads(l); print(g, round(d["ms_per_step"],2), [round(p["ms"],3) for p in d["roofline"]["per_level"]])
PY
