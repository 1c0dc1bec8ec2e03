#!/bin/bash
# round 2, call C: fused sws kernel (4-row TMA stores, pipelined table reads) + the paired-row deblocking kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sws_fused.py -m gpu -q -x > gpurun_out/r2c_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_fused_tests.log
tail -4 gpurun_out/r2c_fused_tests.log | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_h264.py tests/test_gpu_h264chain.py tests/test_gpu_h264flush.py tests/test_gpu_h264lf.py -m gpu -q > gpurun_out/r2c_h264_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_h264_tests.log
tail -12 gpurun_out/r2c_h264_tests.log | cut -c1-300
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py --no-secondary --steps 50 --warmup 5 "$@" > gpurun_out/r2c_bench_$n.json 2> gpurun_out/r2c_bench_$n.err
  python - "$n" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2c_bench_%s.json' % v).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-22s %.0f Mpix/s  %.4f ms  frac %.3f  e2e %s  verified %s" % (v, d["value"], d["ms_per_step"], d["roofline"]["frac"], e.get("value"), d.get("verified")))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2c_bench_%s.err' % v).read()[-1500:])
PY
}
run sws_default --workload sws4k
run sws_w14 --workload sws4k --tune sws_tma_warps=14
run sws_ldg --workload sws4k --tune sws_fused_variant=3
run h264_default --workload h264
run h264_oldblk --workload h264 --tune deblock_variant=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_tma -s 3 -c 1 -f -o gpurun_out/r2c_sws_tma python bench.py --workload sws4k --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2c_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/r2c_launches_h264.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-verify --workload h264 > gpurun_out/r2c_ncu_h264.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_deblock_kernel_v3 -s 2 -c 1 -f -o gpurun_out/r2c_deblock python bench.py --workload h264 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2c_ncu2.log 2>&1
tail -2 gpurun_out/r2c_ncu2.log | cut -c1-200
