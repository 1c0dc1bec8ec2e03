#!/bin/bash
# round 2, call O: 4:2:2 deblocking (decisions + wavefront), the reworked 10-bit wavefront timed
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q > gpurun_out/r2o_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_gpu_tests.log
tail -12 gpurun_out/r2o_gpu_tests.log | cut -c1-400
timeout 600 python bench.py --no-secondary --steps 20 --warmup 3 --workload h264_hbd > gpurun_out/r2o_bench_h264_hbd.json 2> gpurun_out/r2o_bench_h264_hbd.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2o_bench_h264_hbd.json').read().strip().splitlines()[-1])
    print("h264_hbd %.0f Mpix/s  %.4f ms  frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2o_bench_h264_hbd.err').read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file gpurun_out/r2o_launches_h264_hbd.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-verify --workload h264_hbd > gpurun_out/r2o_ncu.log 2>&1
tail -4 gpurun_out/r2o_launches_h264_hbd.csv | cut -c100-420
