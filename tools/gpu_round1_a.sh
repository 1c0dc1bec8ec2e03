mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/bench1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simple_idct_kernel -s 3 -c 1 -o gpurun_out/prof_idct python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/ncu_idct.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused -s 3 -c 1 -o gpurun_out/prof_sws python bench.py --steps 3 --warmup 3 --no-secondary --workload sws4k > gpurun_out/ncu_sws.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
