#!/bin/bash
# First GPU call of the next round: the slots developed against tests/hostsim/ have never run on a B200 (profiles/README.md, last section).
# Runs them on their own first (without -x: every failure is reported), then the whole GPU suite, then the default bench line, so that one
# call answers "did the host-simulated code survive the device" before anything else is built on it.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 600  python -m pytest tests/test_zz_gpu_late_slots.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_late_slots.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
timeout 300  python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1
timeout 600  python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_late_slots.log gpurun_out/r2_gpu_tests.log gpurun_out/r2_smoke.log; head -c 600 gpurun_out/r2_bench.json
