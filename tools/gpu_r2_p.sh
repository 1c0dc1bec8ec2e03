#!/bin/bash
# round 2, call P: FFT with 16-point register passes and bank swizzle
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py tests/test_zz_gpu_h264_hbd.py -m gpu -q -k "fft or mdct or weight_and_dc" > gpurun_out/r2p_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_gpu_tests.log
tail -8 gpurun_out/r2p_gpu_tests.log | cut -c1-400
timeout 600 python bench.py --no-secondary --steps 30 --warmup 5 --workload fft > gpurun_out/r2p_bench_fft.json 2> gpurun_out/r2p_bench_fft.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2p_bench_fft.json').read().strip().splitlines()[-1])
    print("fft %.0f Mpix/s  %.4f ms  frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2p_bench_fft.err').read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 2 -c 1 -f -o gpurun_out/r2p_fft python bench.py --workload fft --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log | cut -c1-200
