#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): NCCL tests through the C ABI, then the driver's bench line at N
N=${1:-2}
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > $o/multi${N}_pytest.log 2>&1; tail -8 $o/multi${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
  bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 > $o/multi${N}_bench.json 2> $o/multi${N}_bench.err
tail -3 $o/multi${N}_bench.err
python - <<PY
import json
for l in open("gpurun_out/multi${N}_bench.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N", d["n_gpus"], "value %.1f M"%(d["value"]/1e6), "ms %.2f"%d["ms_per_step"], "e2e %.1f M %.1f ms"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), "e2e_f32 %.1f M"%(d["e2e_f32"]["value"]/1e6), d["clocks"])
        print("cfg3", d["extras"].get("cfg3_time_sharded"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $o/multi${N}_ref.json 2> $o/multi${N}_ref.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/multi${N}_ref.json') if l.startswith('{')][-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('one_thread_value'))"
