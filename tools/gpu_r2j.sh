#!/bin/bash
# PV + Spec parity on the default build; bench + Spec probe for default and variants
mkdir -p gpurun_out; o=gpurun_out
true
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>$o/var_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
  for n in 512 1024 2048 4096; do echo -n "$name "; env "$@" python tools/spec_probe.py $n $((n/4)) all 2>>$o/var_$name.err; done
}
run default X=1
for v in "$@"; do run $v MELONIX_B200_LIB=variants/$v.so; done
