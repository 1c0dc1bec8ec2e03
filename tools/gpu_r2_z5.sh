#!/bin/bash
# round 2, call Z5: the whole GPU suite again (two refusal expectations moved with the rgb2rgb take-over), the true-rescale workload with the
# word-load / dp2a horizontal pass and the unrolled 4-tap vertical pass of the tile kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z5_gpu_tests.log 2>&1 ) 2> gpurun_out/r2z5_gpu_tests.time; echo "rc=$?" >> gpurun_out/r2z5_gpu_tests.log
grep -v "QMAT\|full chroma\|swscaler" gpurun_out/r2z5_gpu_tests.log | tail -8 | cut -c1-400
for k in 1 2; do timeout 300 python bench.py --no-secondary --steps 40 --warmup 5 --workload sws_up > gpurun_out/r2z5_bench_sws_up_$k.json 2> gpurun_out/r2z5_bench_sws_up_$k.err; python - <<PY
import json
for l in open("gpurun_out/r2z5_bench_sws_up_$k.json"):
    if l.startswith("{"):
        d = json.loads(l); print("sws_up", round(d["value"]), d["roofline"]["frac"], d.get("verified"))
PY
done
tail -3 gpurun_out/r2z5_bench_sws_up_1.err
