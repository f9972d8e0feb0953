#!/bin/bash
# Spec parity tests on the default build, then the batched Spec probe for the default build and the given variants
mkdir -p gpurun_out; o=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_grain.py tests/test_zz_dropin_gpu.py -m gpu -x -q) > $o/r2f_pytest.log 2>&1; tail -3 $o/r2f_pytest.log
run() { name=$1; shift
  for n in 512 1024 2048 4096 8192; do echo -n "$name "; env "$@" python tools/spec_probe.py $n $((n/4)) all 2>>$o/var_$name.err; done
}
run default X=1
for v in "$@"; do run $v MELONIX_B200_LIB=variants/$v.so; done
