#!/bin/bash
# round 2, call V: batched 9 / 10-bit intra reconstruction
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q -x -k "intra or flush" > gpurun_out/r2v_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_gpu_tests.log
tail -25 gpurun_out/r2v_gpu_tests.log | cut -c1-500
