#!/bin/bash
# round 2, call Z1: rgb48 / yuva420p GPU tests; e2e arm with 2..12 concurrent host callers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sws_rgb48_dst.py tests/test_sws_yuva_src.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2z1_sws.log 2>&1; echo "sws rc=$?"; tail -5 gpurun_out/r2z1_sws.log | cut -c1-400
for T in 2 4 6 8 12; do
  AVB200_E2E_STREAMS=$T timeout 300 python bench.py --no-secondary --steps 30 --warmup 5 > gpurun_out/r2z1_bench_e2e_$T.json 2> gpurun_out/r2z1_bench_e2e_$T.err
  python - <<PY
import json
for l in open("gpurun_out/r2z1_bench_e2e_$T.json"):
    if l.startswith("{"):
        d = json.loads(l); print("T=$T value", round(d["value"]), "e2e", d["e2e"])
PY
done
