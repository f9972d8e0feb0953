#!/bin/bash
# PV parity tests on the default build, then the resident bench with and without the environment switches given as
# arguments (NAME=VALUE ...), e.g.  tools/gpu_ab_env.sh MLX_PV_NO_KS32=1
mkdir -p gpurun_out; o=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_pv.py -m gpu -x -q) > $o/abenv_pytest.log 2>&1; tail -4 $o/abenv_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>$o/var_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
}
run default X=1
for v in "$@"; do run "$v" "$v"; done
