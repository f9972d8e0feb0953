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


def test_cpp_export_path(tmp_path, oracle):
    x = S.two_tone(5.0)
    x.tofile(tmp_path / "wav.f32")
    _host_test("export", tmp_path / "wav.f32", 48000, 3.0, tmp_path / "out.i16")
    got = np.fromfile(tmp_path / "out.i16", np.int16)
    o = oracle.grain_export(x, 48000, [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)])
    assert np.array_equal(got, o["pcm16"])
