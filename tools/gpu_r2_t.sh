#!/bin/bash
# round 2, call T: ncu launch list of the default bench command (the metric's config 5 + its end-to-end arm), ncu capture of the headline kernel at the final state
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_default.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/r2t_ncu_launches.log 2>&1
tail -3 gpurun_out/r2t_launches_default.csv | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_tma -s 2 -c 1 -f -o gpurun_out/r2t_sws_tma python bench.py --no-secondary --no-verify --steps 2 --warmup 3 > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log | cut -c1-200
