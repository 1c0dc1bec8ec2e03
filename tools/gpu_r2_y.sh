#!/bin/bash
# round 2, call Y: 4:2:2 intra reconstruction (8 / 9 / 10 bit) + the 4:2:2 flush with intra macroblocks; memcheck on the small cases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q -x -p no:cacheprovider -k "intra or flush" > gpurun_out/r2y_hbd.log 2>&1; echo "hbd rc=$?"; tail -15 gpurun_out/r2y_hbd.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q -x -p no:cacheprovider -k "intra_batch_422 and (3-2 or 7-5) or flush_422 and 7-5 or dc_dequant_batch_422 or (residual_batch_hbd or mc_batch_hbd) and 8-1 and 3-2" > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2y_memcheck.log | head -8 | cut -c1-300
