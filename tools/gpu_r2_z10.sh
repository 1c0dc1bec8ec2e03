#!/bin/bash
# round 2, call Z10: the whole GPU suite on the final state (pal8 sources included)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2z10_gpu_tests.log 2>&1 ) 2> gpurun_out/r2z10_gpu_tests.time; echo "rc=$?" >> gpurun_out/r2z10_gpu_tests.log
grep -v "QMAT\|full chroma\|swscaler" gpurun_out/r2z10_gpu_tests.log | tail -10 | cut -c1-400
