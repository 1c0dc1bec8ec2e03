#!/bin/bash
# round 2, call I: H.264 per-kernel times after the residual rewrite, variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_h264.py tests/test_gpu_h264chain.py tests/test_gpu_h264flush.py tests/test_gpu_h264intra.py -m gpu -q > gpurun_out/r2i_h264_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_h264_tests.log
tail -6 gpurun_out/r2i_h264_tests.log | cut -c1-300
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py --no-secondary --no-verify --steps 50 --warmup 5 "$@" > gpurun_out/r2i_bench_$n.json 2> gpurun_out/r2i_bench_$n.err
  python - "$n" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2i_bench_%s.json' % v).read().strip().splitlines()[-1])
    print("%-22s %.0f Mpix/s  %.4f ms" % (v, d["value"], d["ms_per_step"]))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2i_bench_%s.err' % v).read()[-1500:])
PY
}
run h264_a --workload h264
run h264_b --workload h264
run h264_oldres --workload h264 --tune residual_variant=2
run h264_mb6 --workload h264 --tune mc_min_blocks=6
run h264_mb6_oldres --workload h264 --tune mc_min_blocks=6 --tune residual_variant=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/r2i_launches_h264.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-verify --workload h264 > gpurun_out/r2i_ncu_h264.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_residual_kernel_v2 -s 2 -c 1 -f -o gpurun_out/r2i_residual python bench.py --workload h264 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2i_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_mc_kernel_v2 -s 5 -c 1 -f -o gpurun_out/r2i_mc_avg python bench.py --workload h264 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2i_ncu3.log 2>&1
tail -2 gpurun_out/r2i_ncu3.log | cut -c1-200
