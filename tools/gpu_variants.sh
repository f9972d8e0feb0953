#!/bin/bash
# usage: gpu_call3.sh "name ENV=.. ENV=.." ...   (each argument: a label followed by env assignments)
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
}
for spec in "$@"; do run $spec; done > gpurun_out/variants3.log 2>&1
cat gpurun_out/variants3.log
