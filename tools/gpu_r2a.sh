#!/bin/bash
# round-2 first pass on one B200: the whole -m gpu suite and a short bench
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
nvidia-smi -L | head -3
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
