#!/bin/bash
# round 2, call X: 8-bit 4:2:2 residual / MC / DC transforms / flush through the generic kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_h264_hbd.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2x_hbd.log 2>&1; echo "hbd rc=$?"; tail -15 gpurun_out/r2x_hbd.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_h264.py tests/test_gpu_h264flush.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2x_h264.log 2>&1; echo "h264 rc=$?"; tail -3 gpurun_out/r2x_h264.log
