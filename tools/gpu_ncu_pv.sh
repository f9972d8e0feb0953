#!/bin/bash
# one ncu --set full capture of the PV kernels of the bench step (one launch each), with source; the
# raw / source pages are exported to csv on the box (the .ncu-rep stays there)
mkdir -p gpurun_out
o=gpurun_out
tag=${1:-pv}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pv_ -s 9 -c 3 -o $o/${tag}_prof \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $o/${tag}_ncu.log 2>&1
tail -3 $o/${tag}_ncu.log
ncu -i $o/${tag}_prof.ncu-rep --page raw --csv > $o/${tag}_prof.raw.csv 2>/dev/null
ncu -i $o/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > $o/${tag}_prof.source.csv 2>/dev/null
rm -f $o/${tag}_prof.ncu-rep
ls -la $o/${tag}_prof.*
