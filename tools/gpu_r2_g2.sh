#!/bin/bash
# round 2, 2-GPU call: weak scaling of the default line (incl. the NUMA-bound end-to-end arm) and the frame-sharded full search
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
for n in 1 2; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"; fi
  timeout 600 $L bench.py --gpus $n --steps 50 --warmup 5 --no-secondary > gpurun_out/r2g_sws_n$n.json 2> gpurun_out/r2g_sws_n$n.err
  timeout 600 $L bench.py --gpus $n --steps 50 --warmup 5 --no-secondary --workload me > gpurun_out/r2g_me_n$n.json 2> gpurun_out/r2g_me_n$n.err
  timeout 600 $L bench.py --gpus $n --steps 20 --warmup 5 --no-secondary --workload idct_put > gpurun_out/r2g_idct_n$n.json 2> gpurun_out/r2g_idct_n$n.err
done
python - <<'PY'
import json
for w in ("sws", "me", "idct"):
    for n in (1, 2):
        try:
            d = json.loads(open('gpurun_out/r2g_%s_n%d.json' % (w, n)).read().strip().splitlines()[-1])
            e = d.get("e2e") or {}
            print(w, n, "value %.0f  ms %.4f  e2e %s  scaling %s  verified %s numa %s" % (d["value"], d["ms_per_step"], e.get("value"), d["scaling"], d.get("verified"), d.get("numa")))
        except Exception as ex:
            print(w, n, "FAILED", ex); print(open('gpurun_out/r2g_%s_n%d.err' % (w, n)).read()[-1200:])
PY
