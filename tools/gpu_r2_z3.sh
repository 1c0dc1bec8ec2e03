#!/bin/bash
# round 2, call Z3: the MC kernel built for 6 (default) / 7 / 8 resident CTAs per SM
mkdir -p gpurun_out
for M in 8 9 10 12 8 9 10 12; do
  timeout 300 python bench.py --no-secondary --no-verify --steps 40 --warmup 5 --workload h264 --tune mc_min_blocks=$M > gpurun_out/r2z3_h264_$M.json 2> gpurun_out/r2z3_h264_$M.err
  python - <<PY
import json
for l in open("gpurun_out/r2z3_h264_$M.json"):
    if l.startswith("{"):
        d = json.loads(l); print("minb=$M value", round(d["value"]), "ms", round(d["ms_per_step"], 4))
PY
done
