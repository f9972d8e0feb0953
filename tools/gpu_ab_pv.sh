#!/bin/bash
# A/B on one B200: PV parity tests on the default build, the resident bench for the default build and for variants/NAME.so
# (tools/build_variant.sh), then a short ncu metrics pass (time, instructions, shared-memory wavefronts and conflicts,
# L1 data pipe, issue slots) of the PV kernels
mkdir -p gpurun_out; o=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_pv.py -m gpu -x -q) > $o/r2e_pytest.log 2>&1; tail -3 $o/r2e_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>$o/var_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$name', 'Mframes/s %.2f'%(d['value']/1e6), 'ms %.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
}
run default X=1
for v in "$@"; do run $v MELONIX_B200_LIB=variants/$v.so; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:pv_ -s 9 -c 3 --csv --log-file $o/r2e_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > $o/r2e_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2e_ncu.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:28], r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
