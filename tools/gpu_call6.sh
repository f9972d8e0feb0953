#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_pv.py -m gpu -x -q ) > gpurun_out/pytest_pv.log 2>&1
tail -25 gpurun_out/pytest_pv.log
python tools/gpu_check.py 2>&1 | grep -i "pv " | head -20
bash tools/gpu_call3.sh "default X=1" "poly MELONIX_B200_LIB=variants/poly.so"
