#!/bin/bash
# round-end validation on the GPU box: full GPU suite, smoke, both bench arms, the launch list of the default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final_ncu.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log gpurun_out/final_bench.json
