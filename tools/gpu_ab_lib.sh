#!/bin/bash
# PV parity tests on the default build, then the resident bench for the default build and for variants/NAME.so
mkdir -p gpurun_out; o=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_pv.py -m gpu -x -q) > $o/ablib_pytest.log 2>&1; tail -2 $o/ablib_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>$o/var_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
}
run default X=1
for v in "$@"; do run $v MELONIX_B200_LIB=variants/$v.so; done
