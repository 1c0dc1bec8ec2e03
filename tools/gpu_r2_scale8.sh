#!/bin/bash
# round 2: the driver's 8-GPU launch of both arms (torchrun, one rank per GPU), to look at the end-to-end scaling after the NUMA binding
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s_topo.txt 2>&1
for n in 2 8; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-secondary > gpurun_out/r2s_bench_n$n.json 2> gpurun_out/r2s_bench_n$n.err ) 2> gpurun_out/r2s_bench_n$n.time
  tail -3 gpurun_out/r2s_bench_n$n.time
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2s_bench_n%s.json' % n).read().strip().splitlines()[-1])
    print("N=%s value %.0f Mpix/s frac %.3f e2e %s numa %s" % (n, d["value"], d["roofline"]["frac"], d.get("e2e"), d.get("numa")))
    print({k: v[0] for k, v in d.get("secondary", {}).items()})
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r2s_bench_n%s.err' % n).read()[-2500:])
PY
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --workload me --no-secondary > gpurun_out/r2s_bench_me8.json 2> gpurun_out/r2s_bench_me8.err ) 2> gpurun_out/r2s_bench_me8.time
tail -c 600 gpurun_out/r2s_bench_me8.json; tail -3 gpurun_out/r2s_bench_me8.err
