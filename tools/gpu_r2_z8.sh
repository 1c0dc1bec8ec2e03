#!/bin/bash
# round 2, call Z8: gray8 sources + rgb2rgb converters on the GPU, then every sws GPU test file (the plan changed)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sws_gray_src.py tests/test_sws_rgb2rgb.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2z8_new.log 2>&1; echo "new rc=$?"; grep -v swscaler gpurun_out/r2z8_new.log | tail -6 | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q -k "sws" -p no:cacheprovider > gpurun_out/r2z8_sws.log 2>&1; echo "sws rc=$?"; grep -v swscaler gpurun_out/r2z8_sws.log | tail -6 | cut -c1-400
