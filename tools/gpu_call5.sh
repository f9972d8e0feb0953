#!/bin/bash
# 2-GPU box: multi-GPU tests, weak-scaling bench at N=2, configs[3] time-sharded at full size; then the
# default single-GPU bench line (e2e + cpu baseline) and the reference arm.
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_pv.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1
tail -5 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/cfg4_sharded.py > gpurun_out/cfg4_n2.json 2> gpurun_out/cfg4_n2.err
tail -c 1200 gpurun_out/cfg4_n2.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 800 gpurun_out/bench_ref.json
