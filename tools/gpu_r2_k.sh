#!/bin/bash
# round 2, call K: quant_psnr / bit / rd on the GPU, staged 10-bit IDCT / forward DCT kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_late_slots.py -m gpu -q -x -k "idct10 or fdct10 or quant_metrics" > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log
grep -v QMAT_SHIFT gpurun_out/r2k_tests.log | tail -15 | cut -c1-400
timeout 900 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q > gpurun_out/r2k_tests_hbd.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests_hbd.log
tail -15 gpurun_out/r2k_tests_hbd.log | cut -c1-400
timeout 600 python bench.py --no-secondary --steps 30 --warmup 5 --workload idct10 > gpurun_out/r2k_bench_idct10.json 2> gpurun_out/r2k_bench_idct10.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2k_bench_idct10.json').read().strip().splitlines()[-1])
    print("idct10 %.0f Mpix/s  %.4f ms  frac %.3f verified %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("verified")))
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2k_bench_idct10.err').read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simple_idct10_staged -s 6 -c 1 -f -o gpurun_out/r2k_idct10 python bench.py --workload idct10 --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log | cut -c1-200
