#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cut -c1-400 gpurun_out/bench_n$N.json; python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f M"%(d["value"]/1e6), "ms", d["ms_per_step"], "e2e %.1f M %.1f ms"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), d["clocks"])
PY
tail -2 gpurun_out/bench_n$N.err
