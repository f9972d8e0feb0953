"""-m gpu: the reference's own front-end on the product's GPU classes.

oracle/_ref/libapp_dropin.so is what INTEGRATION.md section 1 produces: /root/reference/app.cpp and
save-wav.cpp, unmodified, compiled against melonix_b200/host/{spec,spec-cache,range,texture} and linked
with libmelonix_b200.so (oracle/Makefile target `dropin`; it travels to the GPU box prebuilt).  The test
drives the reference's App through it -- preproc() constructs OUR Spec on the B200, SpecCache::getTex
uploads OUR fused colour-ramp texels through the recording GL shim -- and compares every column with
the all-reference build (oracle/_ref/libapp_ref.so, CPU)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402

pytestmark = pytest.mark.gpu


def test_reference_front_end_on_gpu_spec_matches_all_reference_build(oracle):
    if not (oracle.have_ref_app() and oracle.have_dropin_app()):
        pytest.skip("oracle/_ref/libapp_ref.so / libapp_dropin.so not built")
    x = S.vibrato_tone(3.0, seed=11)
    width, range_time = 1280, 10.0
    with oracle.RefApp(x, 48000, []) as ref, oracle.RefApp(x, 48000, [], build="dropin") as gpu:
        assert all(np.array_equal(a, b) for a, b in zip(ref.grains(), gpu.grains()))     # same front-end code
        for k in (2.0 ** 15, 2.0 ** 13, 2.0 ** 11):
            for t in (0.5, 1.7, 2.4):
                a = ref.speccache_column(k, width, range_time, t).astype(np.int32)
                b = gpu.speccache_column(k, width, range_time, t).astype(np.int32)
                d = np.abs(a - b)
                # FP32 magnitudes on the GPU vs double in the reference: a texel may move by one
                assert d.max() <= 1 and (d > 0).mean() < 1e-3, (k, t, int(d.max()), float((d > 0).mean()))
