#!/bin/bash
# round 2, call Z4: the whole GPU suite, the default bench line, the reference arm, smoke (state after rgb48 / yuva / rgb2rgb / 4:2:2 intra / MC occupancy)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z4_gpu_tests.log 2>&1 ) 2> gpurun_out/r2z4_gpu_tests.time; echo "rc=$?" >> gpurun_out/r2z4_gpu_tests.log
grep -v "QMAT\|full chroma" gpurun_out/r2z4_gpu_tests.log | tail -12 | cut -c1-400; tail -3 gpurun_out/r2z4_gpu_tests.time
( time timeout 900 python bench.py > gpurun_out/r2z4_bench_default.json 2> gpurun_out/r2z4_bench_default.err ) 2> gpurun_out/r2z4_bench_default.time
tail -3 gpurun_out/r2z4_bench_default.time; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2z4_bench_default.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "e2e", "secondary", "cpu", "verified")}); print(d["roofline"]); print(d["cpu_baseline"])
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2z4_bench_default.err').read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z4_bench_reference.json 2> gpurun_out/r2z4_bench_reference.err; tail -c 700 gpurun_out/r2z4_bench_reference.json


python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
