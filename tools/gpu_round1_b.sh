mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-200
for mb in 4 5 6; do timeout 300 python bench.py --steps 200 --warmup 10 --no-secondary --tune idct_min_blocks=$mb > gpurun_out/bench_idct_mb$mb.json 2>gpurun_out/bench_idct_mb$mb.err; done
timeout 300 python bench.py --steps 200 --warmup 10 --no-secondary --tune idct_min_blocks=5 --tune idct_grid_mult=16 > gpurun_out/bench_idct_mb5_g16.json 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-secondary --workload sws4k > gpurun_out/bench_sws_v2.json 2>gpurun_out/bench_sws_v2.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-secondary --workload sws4k --tune sws_fused_variant=1 > gpurun_out/bench_sws_v1.json 2>&1
for f in gpurun_out/bench_idct_mb*.json gpurun_out/bench_sws_v*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["e2e"]["value"])
except Exception as e: print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sws_fused_rgb24_v2 -s 3 -c 1 -o gpurun_out/prof_sws_v2 python bench.py --steps 3 --warmup 3 --no-secondary --workload sws4k > gpurun_out/ncu_sws_v2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simple_idct_kernel -s 3 -c 1 -o gpurun_out/prof_idct_v2 python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/ncu_idct_v2.log 2>&1
