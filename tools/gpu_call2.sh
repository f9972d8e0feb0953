#!/bin/bash
# tests + resident bench + ncu source capture of the PV kernels; outputs in gpurun_out/
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1
tail -15 gpurun_out/pytest.log
bash tools/bench_variants.sh "$@" > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
python tools/seg_probe.py > gpurun_out/seg_probe.log 2>&1; cat gpurun_out/seg_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pv_ -s 9 -c 3 -o gpurun_out/prof_pv \
  python bench.py --steps 1 --warmup 3 --tracks 32 --no-e2e --no-cpu > gpurun_out/ncu_pv.log 2>&1
ncu -i gpurun_out/prof_pv.ncu-rep --page raw --csv > gpurun_out/prof_pv.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pv.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_pv.source.csv 2>/dev/null
rm -f gpurun_out/prof_pv.ncu-rep
ls -la gpurun_out
