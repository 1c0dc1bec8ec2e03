mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h264.py tests/test_gpu_sws.py tests/test_gpu_slots.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 --no-secondary --workload h264 > gpurun_out/bench_h264.json 2>gpurun_out/bench_h264.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h264.json').read().strip().splitlines()[-1]); print("h264 Mpix/s=%.0f ms=%.4f"%(d["value"], d["ms_per_step"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/launches_h264.csv python bench.py --steps 4 --warmup 3 --no-secondary --workload h264 > gpurun_out/ncu_h264.log 2>&1
