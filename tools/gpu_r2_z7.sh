#!/bin/bash
# round 2, call Z7 (2 GPUs): the driver's torchrun launch of both arms at N=2 on the final state, and config 4 split over 2 GPUs
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z7_bench_n2.json 2> gpurun_out/r2z7_bench_n2.err ) 2> gpurun_out/r2z7_bench_n2.time
tail -3 gpurun_out/r2z7_bench_n2.time
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2z7_bench_n2.json').read().strip().splitlines()[-1])
    print("N=2 value %.0f Mpix/s frac %.3f e2e %s numa %s verified %s" % (d["value"], d["roofline"]["frac"], d.get("e2e"), d.get("numa"), d.get("verified")))
    print({k: v[0] for k, v in d.get("secondary", {}).items()})
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2z7_bench_n2.err').read()[-2500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2z7_bench_ref_n2.json 2> gpurun_out/r2z7_bench_ref_n2.err; tail -c 400 gpurun_out/r2z7_bench_ref_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --workload me --no-secondary > gpurun_out/r2z7_bench_me2.json 2> gpurun_out/r2z7_bench_me2.err; tail -c 500 gpurun_out/r2z7_bench_me2.json; tail -2 gpurun_out/r2z7_bench_me2.err
