#!/bin/bash
# compute-sanitizer over the small parity tests of every kernel family (memcheck, racecheck, synccheck).
mkdir -p gpurun_out
SEL='tiled or golden or ragged or edge_jobs or empty_and_short or per_frame_rate or degenerate or cap_and_device or kat4'
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
      python -m pytest tests/test_gpu_spec.py tests/test_gpu_pv.py tests/test_gpu_grain.py -m gpu -x -q -k "$SEL" ) \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Race|hazard|Invalid|error" gpurun_out/sanitize_$tool.log | head -12
done
