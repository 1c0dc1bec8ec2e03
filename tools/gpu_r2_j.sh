#!/bin/bash
# round 2, call J: multi-stream end-to-end arm, two-chain H.264 step, idct10, default line
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2j_bench_%s.json' % v).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-22s %.0f Mpix/s  %.4f ms  frac %.3f  e2e %s  verified %s" % (v, d["value"], d["ms_per_step"], d["roofline"]["frac"], e.get("value"), d.get("verified")))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2j_bench_%s.err' % v).read()[-1500:])
PY
}
for t in 1 2 4 8; do
  AVB200_E2E_STREAMS=$t timeout 600 python bench.py --no-secondary --steps 30 --warmup 5 > gpurun_out/r2j_bench_sws_t$t.json 2> gpurun_out/r2j_bench_sws_t$t.err; show sws_t$t
done
timeout 600 python bench.py --no-secondary --steps 50 --warmup 5 --workload h264 > gpurun_out/r2j_bench_h264_2chains.json 2> gpurun_out/r2j_bench_h264_2chains.err; show h264_2chains
AVB200_H264_CHAINS=0 timeout 600 python bench.py --no-secondary --steps 50 --warmup 5 --workload h264 > gpurun_out/r2j_bench_h264_1chain.json 2> gpurun_out/r2j_bench_h264_1chain.err; show h264_1chain
timeout 600 python bench.py --no-secondary --steps 30 --warmup 5 --workload idct10 > gpurun_out/r2j_bench_idct10.json 2> gpurun_out/r2j_bench_idct10.err; show idct10
( time timeout 900 python bench.py > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err ) 2> gpurun_out/r2j_bench_default.time
tail -3 gpurun_out/r2j_bench_default.time; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2j_bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "e2e", "secondary", "cpu", "verified")}); print(d["cpu_baseline"])
PY
timeout 900 python -m pytest tests/test_gpu_sws.py tests/test_gpu_sws_fused.py tests/test_gpu_idct.py -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log; tail -4 gpurun_out/r2j_tests.log | cut -c1-300
