#!/bin/bash
# compute-sanitizer over the small parity tests of every kernel family (memcheck, racecheck, synccheck),
# including this round's paths: TMA-staged K_S, int16 wire, batched Spec launch, split analyze/synth, K_A2.
mkdir -p gpurun_out
SEL='tiled or golden or ragged or edge_jobs or empty_and_short or per_frame_rate or degenerate or cap_and_device or kat4 or int16_wire or all_tracks_in_one_launch or host_pipeline'
for tool in memcheck racecheck synccheck; do
  ( time timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
      python -m pytest tests/test_gpu_spec.py tests/test_gpu_pv.py tests/test_gpu_grain.py -m gpu -x -q -k "$SEL" ) \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Race|hazard|Invalid|error" gpurun_out/sanitize_$tool.log | head -12
done
# K_A2 (opt-in kernel) and the general kernel on a ragged four-track batch, analysis + stage export
for tool in memcheck synccheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python tools/dbg_ka2.py 2048 3.0 ) \
      > gpurun_out/sanitize_ka2_$tool.log 2>&1
  echo "== ka2 $tool rc=$?"; grep -E "ERROR SUMMARY|Race|hazard|Invalid|differ" gpurun_out/sanitize_ka2_$tool.log | head -8
done
