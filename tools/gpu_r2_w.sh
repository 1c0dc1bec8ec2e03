#!/bin/bash
# round 2, call W: compute-sanitizer memcheck over the kernels added this round (small cases)
mkdir -p gpurun_out
run() { n=$1; shift; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > gpurun_out/r2w_$n.log 2>&1; echo "$n rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2w_$n.log | head -8 | cut -c1-300; }
run hbd tests/test_zz_gpu_h264_hbd.py -k "3-2 or 7-5 or weight_and_dc or refusals"
run enc tests/test_zz_gpu_late_slots.py -k "quant_metrics_slots or idct10 or fdct10"
run sws tests/test_sws_gray_dst.py tests/test_sws_rgb16_dst.py -k "device_batch"
run fft tests/test_gpu_fft.py
run lf tests/test_gpu_h264lf.py -k "field or decisions_match"
