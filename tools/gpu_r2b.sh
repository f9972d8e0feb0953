#!/bin/bash
# round-2 pass on one B200: the whole -m gpu suite, the full default bench line, the reference arm
mkdir -p gpurun_out
o=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $o/r2b_pytest.log 2>&1; tail -12 $o/r2b_pytest.log
(time python bench.py) > $o/r2b_bench.json 2> $o/r2b_bench.err; tail -4 $o/r2b_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2b_bench.json") if l.startswith("{")][-1])
print("value %.2f M ms %.2f"%(d["value"]/1e6,d["ms_per_step"]), d["roofline"]["kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"])
for k in ("e2e","e2e_f32"):
    e=d.get(k) or {}
    print(k, "%.2f M frames/s, %.1f ms"%(e.get("value",0)/1e6, e.get("ms_per_step",0)), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"))
print("cpu", d["cpu_baseline"])
x=d.get("extras",{})
print("cfg3", x.get("cfg3_time_sharded"))
oc=x.get("other_configs",{})
for r in oc.get("spec",[]): print("spec", r["fftN"], "%.1f M f/s %.0f GB/s %.3f"%(r["frames_per_s"]/1e6, r["achieved_gbs"], r["frac_of_hbm_peak"]), "%.3f ms"%r["kernel_ms_per_launch"])
for r in oc.get("pv",[]): print("pv", r["fftN"], "%.1f M f/s %.3f"%(r["frames_per_s"]/1e6, r["frac_of_hbm_peak"]), {k:round(v,2) for k,v in r["kernel_ms"].items()})
print(oc.get("cfg1_single_60s_track"), oc.get("cfg0_spec_10s"), oc.get("error"))
PY
