#!/bin/bash
# round 2, call Q: gray8 on the GPU; H.264 composite with 2 / 3 / 4 independent chains per step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sws_gray_dst.py tests/test_sws_rgb16_dst.py -m gpu -q > gpurun_out/r2q_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_gpu_tests.log
grep -v "full chroma" gpurun_out/r2q_gpu_tests.log | tail -6 | cut -c1-400
for g in 2 3 4; do
  AVB200_H264_GROUPS=$g timeout 600 python bench.py --no-secondary --steps 30 --warmup 5 --workload h264 > gpurun_out/r2q_bench_h264_g$g.json 2> gpurun_out/r2q_bench_h264_g$g.err
  python - $g <<'PY'
import json, sys
g = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2q_bench_h264_g%s.json' % g).read().strip().splitlines()[-1])
    print("h264 groups=%s %.0f Mpix/s  %.4f ms  verified %s" % (g, d["value"], d["ms_per_step"], d.get("verified")))
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2q_bench_h264_g%s.err' % g).read()[-1500:])
PY
done
