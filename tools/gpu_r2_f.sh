#!/bin/bash
# round 2, call F: MC occupancy variants, the default bench line end to end, full GPU suite
mkdir -p gpurun_out
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py --no-secondary --steps 50 --warmup 5 "$@" > gpurun_out/r2f_bench_$n.json 2> gpurun_out/r2f_bench_$n.err
  python - "$n" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2f_bench_%s.json' % v).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-22s %.0f Mpix/s  %.4f ms  frac %.3f  e2e %s  verified %s" % (v, d["value"], d["ms_per_step"], d["roofline"]["frac"], e.get("value"), d.get("verified")))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2f_bench_%s.err' % v).read()[-1500:])
PY
}
run h264_mb4 --workload h264
run h264_mb5 --workload h264 --tune mc_min_blocks=5
run h264_mb6 --workload h264 --tune mc_min_blocks=6
( time timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err ) 2> gpurun_out/r2f_bench_default.time
tail -3 gpurun_out/r2f_bench_default.time; head -c 3000 gpurun_out/r2f_bench_default.json; echo; tail -5 gpurun_out/r2f_bench_default.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; head -c 900 gpurun_out/r2f_bench_reference.json; echo; tail -3 gpurun_out/r2f_bench_reference.err
( time timeout 600 python bench.py --impl reference --workload h264 ) > gpurun_out/r2f_bench_reference_h264.json 2> gpurun_out/r2f_bench_reference_h264.err; head -c 600 gpurun_out/r2f_bench_reference_h264.json; echo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_gpu_tests.log
tail -6 gpurun_out/r2f_gpu_tests.log | cut -c1-300
