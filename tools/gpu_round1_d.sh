mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sws.py tests/test_gpu_idct.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
AVB200_TUNE=sws_fused_variant=3 timeout 900 python -m pytest tests/test_gpu_sws.py -m gpu -q > gpurun_out/pytest_gpu_v3.log 2>&1; echo "pytest v3 rc=$?" >> gpurun_out/pytest_gpu_v3.log
tail -3 gpurun_out/pytest_gpu_v3.log | cut -c1-200
run() { name=$1; shift; timeout 300 python bench.py --steps 200 --warmup 10 --no-secondary "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err; python - gpurun_out/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "Mpix/s=%.0f ms=%.4f frac=%.3f clk=%s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
except Exception as e: print(sys.argv[2], "ERR", e, open(sys.argv[1]).read()[-300:])
PY
}
run idct_new
run idct_new_mb4 --tune idct_min_blocks=4
run idct_new_g8 --tune idct_grid_mult=8
run sws_v2 --workload sws4k
run sws_v3 --workload sws4k --tune sws_fused_variant=3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_h264.csv python bench.py --steps 5 --warmup 3 --no-secondary --workload h264 > gpurun_out/ncu_h264.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_v3 -s 3 -c 1 -o gpurun_out/prof_sws_v3 python bench.py --steps 3 --warmup 3 --no-secondary --workload sws4k --tune sws_fused_variant=3 > gpurun_out/ncu_sws_v3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simple_idct_kernel -s 3 -c 1 -o gpurun_out/prof_idct_new python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/ncu_idct_new.log 2>&1
