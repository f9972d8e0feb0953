"""-m gpu: grain path (K6) and the C++ drop-in classes.  The grain resampler is float arithmetic
that the kernel reproduces operation for operation: bit-exact against the oracle's restatement of
App::process / App::exportWav (reference app.cpp:294-345, 1194-1215)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("markers", [
    [], "p3", "m3",
    [(100000, 0, 0.5, 2.0), (400000, 0, -0.3, -1.5), (900000, 0, 0.0, 4.0)],
])
def test_export_bit_exact(engine, oracle, markers):
    from melonix_b200 import hostlib as H
    x = S.two_tone(20.0) if not isinstance(markers, list) or markers else S.two_tone(6.0)
    if markers == "p3":
        markers = [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)]
    elif markers == "m3":
        markers = [(10, 0, 0, -3.0), (x.size - 10, 0, 0, -3.0)]
    engine.upload_tracks([x])
    pcm, pcm16 = H.export_wav(engine, 0, x, 48000, markers)
    o = oracle.grain_export(x, 48000, markers)
    assert pcm.size == o["pcm"].size
    assert np.array_equal(pcm.view(np.uint32), o["pcm"].view(np.uint32))   # float bit patterns
    assert np.array_equal(pcm16, o["pcm16"])


def test_kat4_identity_and_golden(engine):
    from melonix_b200 import hostlib as H
    x = S.two_tone(20.0)
    engine.upload_tracks([x])
    pcm, _ = H.export_wav(engine, 0, x, 48000, [])
    gs, gl = H.grain_segment(x)
    total = int(gl.sum())
    assert pcm.size == total + 1500 and np.array_equal(pcm[:total], x[:total]) and not pcm[total:].any()
    g = np.load(GOLD / "grain_p3.npz")
    x = S.two_tone(float(g["seconds"]))
    engine.upload_tracks([x])
    pcm, pcm16 = H.export_wav(engine, 0, x, 48000, [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)])
    assert np.array_equal(pcm16, g["pcm16"]) and np.array_equal(pcm, g["pcm"])


def test_degenerate_inputs(engine, oracle):
    from melonix_b200 import hostlib as H
    for x in (np.zeros(1200, np.float32), S.two_tone(0.05)):
        engine.upload_tracks([x])
        pcm, pcm16 = H.export_wav(engine, 0, x, 48000, [])
        o = oracle.grain_export(x, 48000, [])
        assert np.array_equal(pcm, o["pcm"]) and np.array_equal(pcm16, o["pcm16"])


def _host_test(*args, env=None):
    exe = ROOT / "melonix_b200" / "host_test"
    assert exe.exists(), "run __graft_entry__.build()"
    import os
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_spec_class_async_contract_and_values(tmp_path, oracle):
    """The C++ `Spec` drop-in (melonix_b200/host/spec.cpp): first getSpec of a key returns {} (KAT-3),
    later calls return SpectrSize/2 floats equal to the oracle / the reference's spec.cpp."""
    x = S.vibrato_tone(1.0, seed=9)
    jobs = S.regular_jobs(x.size, 375)[:96]
    x.tofile(tmp_path / "wav.f32")
    jobs.tofile(tmp_path / "jobs.i32")
    out = _host_test("spec", tmp_path / "wav.f32", tmp_path / "jobs.i32", tmp_path / "out.f32")
    assert "first_call_empty=96" in out and "remaining=0" in out and "half=16384" in out
    got = np.fromfile(tmp_path / "out.f32", np.float32).reshape(96, 16384)
    ref = oracle.ref_spec_run(x, jobs) if oracle.have_ref() else oracle.spec_batch(x, 32768, jobs)
    assert np.sqrt(np.mean((got.astype(np.float64) - ref) ** 2)) < 1e-8


def test_cpp_speccache_textures(tmp_path, oracle):
    x = S.vibrato_tone(0.5, seed=10)
    x.tofile(tmp_path / "wav.f32")
    width, range_time, k = 48, 0.5, 2.0 ** 13
    out = _host_test("speccache", tmp_path / "wav.f32", k, width, range_time, tmp_path / "out.u8",
                     env={"MELONIX_SPECTR_SIZE": "4096"})
    assert "not_ready=0" in out
    got = np.fromfile(tmp_path / "out.u8", np.uint8).reshape(width, 2048, 3).astype(np.int32)
    jobs = np.array([[int((c * range_time / width) * 48000), int((c * range_time / width + range_time / width) * 48000)]
                     for c in range(width)], np.int32)
    ref = oracle.colormap(oracle.spec_batch(x, 4096, jobs), k).astype(np.int32)
    d = np.abs(got - ref)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_cpp_spec_recolours_cached_columns_at_once(tmp_path, oracle):
    """Brightness change on warm columns: with the floats cached, Spec::getSpecRgb answers on
    its first call with the host colour ramp -- byte-identical to the reference's populateTex arithmetic
    on those floats -- instead of relaunching and returning {} (a black flash) meanwhile."""
    x = S.vibrato_tone(0.5, seed=12)
    jobs = S.regular_jobs(x.size, 375)[:40]
    x.tofile(tmp_path / "wav.f32")
    jobs.tofile(tmp_path / "jobs.i32")
    k = 2.0 ** 12
    out = _host_test("recolour", tmp_path / "wav.f32", tmp_path / "jobs.i32", k, tmp_path / "out.u8",
                     env={"MELONIX_SPECTR_SIZE": "4096"})
    assert "not_ready=0" in out and "immediate=40" in out
    got = np.fromfile(tmp_path / "out.u8", np.uint8).reshape(40, 2048, 3).astype(np.int32)
    ref = oracle.colormap(oracle.spec_batch(x, 4096, jobs), k).astype(np.int32)
    d = np.abs(got - ref)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3   # FP32 magnitudes on the GPU vs double in the oracle


def test_cpp_export_path(tmp_path, oracle):
    x = S.two_tone(5.0)
    x.tofile(tmp_path / "wav.f32")
    _host_test("export", tmp_path / "wav.f32", 48000, 3.0, tmp_path / "out.i16")
    got = np.fromfile(tmp_path / "out.i16", np.int16)
    o = oracle.grain_export(x, 48000, [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)])
    assert np.array_equal(got, o["pcm16"])


# ------------------------------------------------------------------------------------------------
# K8: grain segmentation on the device (App::preproc, reference app.cpp:156-235) -- integer-exact
def _seg_cases():
    fs = 48000
    rng = np.random.default_rng(8)

    def tone(f, sec, noise=0.0):
        t = np.arange(int(sec * fs)) / fs
        x = 0.4 * np.sin(2 * np.pi * f * t) + 0.2 * np.sin(2 * np.pi * 2.01 * f * t + 1.0)
        return (x + noise * (rng.random(t.size) - 0.5)).astype(np.float32)

    weird = tone(440, 5)
    weird[::97] = -0.0           # -0.0 counts as ">= 0" (app.cpp:175)
    weird[5::1013] = np.nan      # NaN passes both sign tests
    return {
        "two_tone": S.two_tone(20.0),
        "vibrato": S.vibrato_tone(30.0, seed=4),
        "noisy": tone(220, 20, 0.05),
        "sparse_33Hz": tone(33, 20),
        "fallback_9Hz": tone(9, 30),               # no look-7 crossing within +-749: forward look-3 scan
        "fallback_9Hz_noisy": tone(9, 30, 0.02),
        "white": tone(0, 10, 1.0),
        "silence": np.zeros(100000, np.float32),
        "negative_dc": np.full(100000, -0.25, np.float32),
        "clip_1400": tone(220, 1400 / fs),          # shorter than one grain: the loop never runs (app.cpp:161)
        "clip_1502": tone(220, 1502 / fs),
        "clip_3100": tone(220, 3100 / fs),
        "empty": np.zeros(0, np.float32),
        "neg_zero_and_nan": weird,
        "long": S.vibrato_tone(200.0, seed=5),      # several 256 Ki-sample stage refills per chain
    }


def test_grain_segmentation_matches_reference_restatement(engine, oracle):
    from melonix_b200 import hostlib as H
    cases = _seg_cases()
    names = list(cases)
    engine.upload_tracks([cases[k] for k in names])      # all tracks segmented by ONE call
    got = engine.grain_segment()
    for name, (gs, gl) in zip(names, got):
        os_, ol = oracle.grain_segment(cases[name])
        assert gs.size == os_.size, name
        assert np.array_equal(gs, os_) and np.array_equal(gl, ol), name
        hs, hl = H.grain_segment(cases[name])              # host C++ mirror
        assert np.array_equal(gs, hs) and np.array_equal(gl, hl), name
    counts = dict(zip(names, (g[0].size for g in got)))
    assert counts["two_tone"] == 628                      # KAT-4 probe of SURVEY.md section 4
    assert counts["silence"] == counts["negative_dc"] == counts["clip_1400"] == counts["empty"] == 0
    assert counts["fallback_9Hz"] > 0 and counts["long"] > 6000


def test_grain_segmentation_cap_and_device_entry(engine, oracle):
    import torch
    x = S.two_tone(20.0)
    engine.upload_tracks([x, x[:200000]])
    os_, ol = oracle.grain_segment(x)
    got = engine.grain_segment(cap=100)                     # too small: rows truncated, counts complete
    assert got[0][0].size == 100 and int(engine.last_grain_counts[0]) == os_.size
    assert np.array_equal(got[0][0], os_[:100]) and np.array_equal(got[0][1], ol[:100])
    engine.use_torch_stream()
    cap = x.size // 751 + 1
    gs = torch.zeros((2, cap), dtype=torch.int32, device="cuda")
    gl = torch.zeros((2, cap), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    engine.grain_segment_dev(gs, gl, cnt, cap)
    torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    assert c[0] == os_.size
    assert np.array_equal(gs[0, :c[0]].cpu().numpy(), os_) and np.array_equal(gl[0, :c[0]].cpu().numpy(), ol)
    o2s, o2l = oracle.grain_segment(x[:200000])
    assert c[1] == o2s.size and np.array_equal(gs[1, :c[1]].cpu().numpy(), o2s)


def test_grain_segmentation_golden(engine):
    """Committed fixture (tests/golden/grain_p3.npz, made by tests/golden/make_golden.py from the oracle)."""
    g = np.load(GOLD / "grain_p3.npz")
    x = S.two_tone(float(g["seconds"]))
    engine.upload_tracks([x])
    gs, gl = engine.grain_segment()[0]
    assert np.array_equal(gs, g["g_start"]) and np.array_equal(gl, g["g_len"])


def test_grain_segmentation_two_hour_track(engine):
    """BASELINE configs[3] length (2 h at 48 kHz, 345.6 M samples): positions beyond 2^28, ~10.8 M words of
    crossing bits, ~1300 stage refills in the chain; against the ORACLE (oracle/grain_ref.c, pinned to the
    reference's own App::preproc in tests/test_oracle.py) and the product's host mirror."""
    from oracle import oracle as O
    from melonix_b200 import hostlib as H
    base = S.vibrato_tone(60.0, seed=21)
    x = np.tile(base, 120)
    x[1::100003] *= -1.0                      # break the exact periodicity of the tiling
    assert x.size == 345_600_000
    engine.upload_tracks([x])
    gs, gl = engine.grain_segment()[0]
    os_, ol = O.grain_segment(x)
    assert gs.size == os_.size > 200_000
    assert np.array_equal(gs, os_) and np.array_equal(gl, ol)
    hs, hl = H.grain_segment(x)
    assert np.array_equal(gs, hs) and np.array_equal(gl, hl)
    assert int(gs[-1]) > (1 << 28)
    engine.upload_tracks([np.zeros(16, np.float32)])   # release the 1.4 GB track buffer for later tests
