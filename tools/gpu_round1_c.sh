mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sws.py tests/test_gpu_idct.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-200
run() { name=$1; shift; timeout 300 python bench.py --steps 200 --warmup 10 --no-secondary "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err; python - gpurun_out/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "Mpix/s=%.0f ms=%.4f frac=%.3f clk=%s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
except Exception as e: print(sys.argv[2], "ERR", e, open(sys.argv[1]).read()[-300:])
PY
}
run idct_g16 --tune idct_grid_mult=16
run idct_g32 --tune idct_grid_mult=32
run idct_g64 --tune idct_grid_mult=64
run idct_mh1_g16 --tune idct_mulhi=1 --tune idct_grid_mult=16
run idct_mh2_g16 --tune idct_mulhi=2 --tune idct_grid_mult=16
run idct_mh1_g32 --tune idct_mulhi=1 --tune idct_grid_mult=32
run sws_v2 --workload sws4k
run sws_v2_mh --workload sws4k --tune sws_mulhi=1
run h264 --workload h264
run me --workload me
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_v2 -s 3 -c 1 -o gpurun_out/prof_sws_v2mh python bench.py --steps 3 --warmup 3 --no-secondary --workload sws4k --tune sws_mulhi=1 > gpurun_out/ncu_sws_v2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simple_idct_kernel -s 3 -c 1 -o gpurun_out/prof_idct_mh python bench.py --steps 3 --warmup 3 --no-secondary --tune idct_mulhi=1 --tune idct_grid_mult=16 > gpurun_out/ncu_idct_mh.log 2>&1
