mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_misc.py -m gpu -q -k "full_search" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
run() { name=$1; shift; timeout 300 python bench.py --steps 50 --warmup 3 --no-secondary "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err; python - gpurun_out/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "Mpix/s=%.0f ms=%.4f frac=%.3f clk=%s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
except Exception as e: print(sys.argv[2], "ERR", e, open(sys.argv[1]).read()[-300:], open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
run me_v2 --workload me
run me_v1 --workload me --tune full_search_variant=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/launches_h264.csv python bench.py --steps 4 --warmup 3 --no-secondary --workload h264 > gpurun_out/ncu_h264.log 2>&1
