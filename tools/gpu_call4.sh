#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1
tail -6 gpurun_out/pytest.log
python tools/extra_bench.py > gpurun_out/extra_default.json 2> gpurun_out/extra_default.err
MELONIX_B200_LIB=variants/ka4096.so python tools/extra_bench.py > gpurun_out/extra_ka4096.json 2> gpurun_out/extra_ka4096.err
python - <<'PY'
import json
for n in ("default","ka4096"):
    d=json.load(open(f"gpurun_out/extra_{n}.json"))
    print(n, "pv", [(p["fftN"], round(p["frames_per_s"]/1e6,1), {k:round(v,2) for k,v in p["kernel_ms"].items()}) for p in d["pv"]])
    print(n, "spec", [(p["fftN"], p["hop"], round(p["frames_per_s"]/1e6,1), round(p["frac_of_hbm_peak"],3)) for p in d["spec"]])
PY
