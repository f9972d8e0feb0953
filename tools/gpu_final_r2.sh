#!/bin/bash
# Round-2 end capture on one B200: tests, smoke, the driver's two bench arms, launch list, full ncu captures
# (PV kernels on the bench batch; the batched Spec launch).  Everything lands in gpurun_out/final_*;
# tools/summarise_profiles.py r2 turns it into profiles/.
mkdir -p gpurun_out
o=gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > $o/final_pytest.log 2>&1; tail -4 $o/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/final_smoke.log 2>&1; tail -2 $o/final_smoke.log
python bench.py --impl reference > $o/final_bench_ref.json 2> $o/final_bench_ref.err
python bench.py > $o/final_bench_n1.json 2> $o/final_bench_n1.err; tail -c 400 $o/final_bench_n1.json
python tools/seg_probe.py > $o/final_seg_probe.log 2>&1; cat $o/final_seg_probe.log
python tools/picks_probe.py > $o/final_picks_probe.log 2>&1; cat $o/final_picks_probe.log
python tools/grain_probe.py > $o/final_grain_probe.log 2>&1; cat $o/final_grain_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pv_|spec_|grain_|pcm16" -c 60 --csv --log-file $o/final_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > $o/final_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pv_ -s 9 -c 3 -o $o/final_prof_pv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > $o/final_ncu_pv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_frames -s 3 -c 1 -o $o/final_prof_spec \
  python tools/spec_probe.py 1024 256 all > $o/final_ncu_spec.log 2>&1
for r in pv spec; do
  ncu -i $o/final_prof_$r.ncu-rep --page raw --csv > $o/final_prof_$r.raw.csv 2>/dev/null
  ncu -i $o/final_prof_$r.ncu-rep --page source --csv --print-source cuda,sass > $o/final_prof_$r.source.csv 2>/dev/null
done
rm -f $o/final_prof_*.ncu-rep
du -sh $o/final_* | tail -20
