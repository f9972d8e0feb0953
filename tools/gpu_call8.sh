#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_grain.py -m gpu -x -q ) > gpurun_out/pytest_grain.log 2>&1
tail -6 gpurun_out/pytest_grain.log
python tools/seg_probe.py
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:grain_c -s 2 -c 2 python tools/seg_probe.py 2>&1 | grep -E "grain_c|duration|dram__bytes" | head -8
