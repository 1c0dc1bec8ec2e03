#!/bin/bash
# round 2, call A: the TMA-staged fused sws kernel -- parity, variant sweep, launch list, one --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sws_fused.py -m gpu -q -x > gpurun_out/r2b_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fused_tests.log
tail -25 gpurun_out/r2b_fused_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_sws.py tests/test_sws_colorspace.py tests/test_sws_rgb32_dst.py tests/test_sws_nv12.py -m gpu -q > gpurun_out/r2b_sws_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_sws_tests.log
tail -8 gpurun_out/r2b_sws_tests.log | cut -c1-300
for v in "default" "sws_tma_store=2" "sws_tma_warps=4" "sws_tma_lut=2" "sws_fused_variant=3"; do
  t=""; [ "$v" != "default" ] && t="--tune $v"
  timeout 300 python bench.py --workload sws4k --no-secondary --steps 100 --warmup 10 $t > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2b_bench_%s.json' % v).read().strip().splitlines()[-1])
    print("%-22s %.0f Mpix/s  %.4f ms  frac %.3f  e2e %.0f  clk %s" % (v, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(v, "FAILED", e); print(open('gpurun_out/r2b_bench_%s.err' % v).read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_tma -s 3 -c 1 -f -o gpurun_out/r2b_sws_tma python bench.py --workload sws4k --no-secondary --steps 2 --warmup 3 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log | cut -c1-200
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sws_fused.py -m gpu -q -x -k "272 or 528" > gpurun_out/r2b_memcheck.log 2>&1
tail -5 gpurun_out/r2b_memcheck.log | cut -c1-200
