#!/bin/bash
# round 2, call Z2: ncu --set full of the H.264 composite's top kernel (MC) and of the true-rescale tile kernel at the final state
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_mc_kernel_v2 -s 6 -c 1 -f -o gpurun_out/r2z2_h264_mc python bench.py --no-secondary --no-verify --steps 2 --warmup 3 --workload h264 > gpurun_out/r2z2_ncu_mc.log 2>&1
tail -2 gpurun_out/r2z2_ncu_mc.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_tile_rgb24 -s 2 -c 1 -f -o gpurun_out/r2z2_sws_tile python bench.py --no-secondary --no-verify --steps 2 --warmup 3 --workload sws_up > gpurun_out/r2z2_ncu_tile.log 2>&1
tail -2 gpurun_out/r2z2_ncu_tile.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h264_intra_kernel -s 2 -c 1 -f -o gpurun_out/r2z2_h264_intra python bench.py --no-secondary --no-verify --steps 2 --warmup 3 --workload h264_intra > gpurun_out/r2z2_ncu_intra.log 2>&1
tail -2 gpurun_out/r2z2_ncu_intra.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -5
