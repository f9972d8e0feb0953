#!/bin/bash
# One GPU call: tests, kernel variants, Spec probes, ncu captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1
tail -5 gpurun_out/pytest.log
bash tools/bench_variants.sh gv1 ka2 > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
for cfg in "1024 256" "2048 512" "512 128" "4096 1024" "8192 2048" "1024 1024"; do
  python tools/spec_probe.py $cfg
  MLX_SPEC_GENERIC=1 python tools/spec_probe.py $cfg | sed 's/^/generic: /'
done > gpurun_out/spec_probe.log 2>&1
cat gpurun_out/spec_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pv_ -s 9 -c 3 -o gpurun_out/prof_pv \
  python bench.py --steps 1 --warmup 3 --tracks 32 --no-e2e --no-cpu > gpurun_out/ncu_pv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_frames -s 3 -c 1 -o gpurun_out/prof_spec \
  python tools/spec_probe.py 1024 256 > gpurun_out/ncu_spec.log 2>&1
ls -la gpurun_out
for r in prof_pv prof_spec; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$r.source.csv 2>/dev/null
done
du -sh gpurun_out/*
