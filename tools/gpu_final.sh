#!/bin/bash
# Round-end capture on one B200: tests, the driver's bench lines, secondary sweeps, ncu launch list and
# full captures.  Everything lands in gpurun_out/final_*; tools/summarise_profiles.py turns it into profiles/.
mkdir -p gpurun_out
o=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $o/final_pytest.log 2>&1; tail -4 $o/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/final_smoke.log 2>&1; tail -2 $o/final_smoke.log
# opt-in: the reference's own front-end on the product's GPU Spec / SpecCache (oracle/_ref/libapp_dropin.so)
MLX_RUN_DROPIN_GPU=1 python -m pytest tests/test_zz_dropin_gpu.py -m gpu -q > $o/final_dropin.log 2>&1; tail -3 $o/final_dropin.log
python bench.py --impl reference > $o/final_bench_ref.json 2> $o/final_bench_ref.err
python bench.py > $o/final_bench_n1.json 2> $o/final_bench_n1.err; tail -c 600 $o/final_bench_n1.json
python tools/extra_bench.py > $o/final_extra.json 2> $o/final_extra.err
python tools/seg_probe.py > $o/final_seg_probe.log 2>&1; cat $o/final_seg_probe.log
python tools/picks_probe.py > $o/final_picks_probe.log 2>&1; cat $o/final_picks_probe.log
python tools/grain_probe.py > $o/final_grain_probe.log 2>&1; cat $o/final_grain_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pv_|spec_|grain_" -c 60 --csv --log-file $o/final_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $o/final_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pv_ -s 9 -c 3 -o $o/final_prof_pv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $o/final_ncu_pv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_frames -s 3 -c 1 -o $o/final_prof_spec \
  python tools/spec_probe.py 1024 256 > $o/final_ncu_spec.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grain_c -s 2 -c 2 -o $o/final_prof_seg \
  python tools/seg_probe.py 64 > $o/final_ncu_seg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:picks_ -s 4 -c 2 -o $o/final_prof_picks \
  python tools/picks_probe.py 64 > $o/final_ncu_picks.log 2>&1
for r in pv spec seg picks; do
  ncu -i $o/final_prof_$r.ncu-rep --page raw --csv > $o/final_prof_$r.raw.csv 2>/dev/null
  ncu -i $o/final_prof_$r.ncu-rep --page source --csv --print-source cuda,sass > $o/final_prof_$r.source.csv 2>/dev/null
done
rm -f $o/final_prof_*.ncu-rep
du -sh $o/final_* | tail -20
