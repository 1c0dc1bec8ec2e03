#!/bin/bash
# round 2, call Z6: the GPU suite after the VLC-table cache fix
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z6_gpu_tests.log 2>&1 ) 2> gpurun_out/r2z6_gpu_tests.time; echo "rc=$?" >> gpurun_out/r2z6_gpu_tests.log
grep -v "QMAT\|full chroma\|swscaler" gpurun_out/r2z6_gpu_tests.log | tail -8 | cut -c1-400
