#!/bin/bash
# K_A2 bring-up: parity of the new analysis kernel, then its timing next to the general kernel
mkdir -p gpurun_out
o=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_pv.py -m gpu -x -q) > $o/ka2_pytest.log 2>&1; tail -25 $o/ka2_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>$o/ka2_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
}
run ka2 MLX_PV_KA2=1
run general MLX_PV_NO_KA2=1
for v in "$@"; do run $v MLX_PV_KA2=1 MELONIX_B200_LIB=variants/$v.so; done
