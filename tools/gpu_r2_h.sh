#!/bin/bash
# round 2, call H: H.264 per-kernel times after the residual rewrite, variants
mkdir -p gpurun_out
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py --no-secondary --no-verify --steps 50 --warmup 5 "$@" > gpurun_out/r2h_bench_$n.json 2> gpurun_out/r2h_bench_$n.err
  python - "$n" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2h_bench_%s.json' % v).read().strip().splitlines()[-1])
    print("%-22s %.0f Mpix/s  %.4f ms" % (v, d["value"], d["ms_per_step"]))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2h_bench_%s.err' % v).read()[-1500:])
PY
}
run h264_a --workload h264
run h264_b --workload h264
run h264_oldres --workload h264 --tune residual_variant=2
run h264_mb6 --workload h264 --tune mc_min_blocks=6
run h264_mb6_oldres --workload h264 --tune mc_min_blocks=6 --tune residual_variant=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/r2h_launches_h264.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-verify --workload h264 > gpurun_out/r2h_ncu_h264.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_residual_kernel_v2 -s 2 -c 1 -f -o gpurun_out/r2h_residual python bench.py --workload h264 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2h_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_mc_kernel_v2 -s 5 -c 1 -f -o gpurun_out/r2h_mc_avg python bench.py --workload h264 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2h_ncu3.log 2>&1
tail -2 gpurun_out/r2h_ncu3.log | cut -c1-200
