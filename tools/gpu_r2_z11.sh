#!/bin/bash
# round 2, call Z11: smoke + the default workload's line on the final commit
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --no-secondary --steps 40 --warmup 5 > gpurun_out/r2z11_bench.json 2> gpurun_out/r2z11_bench.err; python - <<'PY'
import json
for l in open("gpurun_out/r2z11_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d.get("verified"))
PY
tail -2 gpurun_out/r2z11_bench.err
