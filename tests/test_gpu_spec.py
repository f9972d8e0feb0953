"""-m gpu: Spec STFT path (K1 / K7) through the C ABI against the oracle and golden vectors.
Tolerances: magnitudes are float (FP32 FFT on the GPU vs double in the reference) -> RMS <= 1e-4
as BASELINE.json states, in practice ~1e-9; RGB texels (truncating casts) may differ by 1 LSB."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def test_config1_sweep_1024_256(engine, oracle):
    """BASELINE configs[0]: 10 s 48 kHz sine sweep, 1024-FFT / 256-hop, all 1875 frames."""
    x = S.sine_sweep(10.0)
    jobs = S.regular_jobs(x.size, 256)
    assert jobs.shape[0] == 1875
    engine.upload_tracks([x])
    got = engine.spec_batch(0, 1024, jobs)
    ref = oracle.spec_batch(x, 1024, jobs)
    assert got.shape == (1875, 512)
    assert rms(got, ref) <= 1e-4 and rms(got, ref) < 1e-7
    assert np.array_equal(got.argmax(1)[50:-50], ref.argmax(1)[50:-50])


def test_reference_geometry_and_golden(engine, oracle):
    """N = 32768 (reference SpectrSize), hop 375; golden rows come from the reference's spec.cpp."""
    g = np.load(GOLD / "spec_ref_geometry.npz")
    n = int(g["n"])
    kat = (0.5 * np.sin(2 * np.pi * 300 * np.arange(n) / 32768)).astype(np.float32)
    engine.upload_tracks([kat])
    got = engine.spec_batch(0, 32768, g["jobs"])
    assert got.shape == g["out"].shape
    assert rms(got, g["out"]) < 1e-8
    assert got[0].argmax() == 300 and abs(float(got[0].max()) - 0.03340026) < 1e-6   # KAT-1
    assert not got[2].any() and not got[4].any()                                       # KAT-2
    if oracle.have_ref():
        x = S.vibrato_tone(1.0, seed=5)
        jobs = S.regular_jobs(x.size, 375)[:64]
        engine.upload_tracks([x])
        assert rms(engine.spec_batch(0, 32768, jobs), oracle.ref_spec_run(x, jobs)) < 1e-8


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_all_sizes_and_edge_jobs(engine, oracle, N):
    x = S.vibrato_tone(2.0, seed=N)
    engine.upload_tracks([x])
    jobs = np.concatenate([
        S.regular_jobs(x.size, N // 4)[:: max(1, (x.size // (N // 4)) // 40)],
        np.array([[-5000, -10], [x.size + 40000, x.size + 40100], [-3, 100], [x.size - 7, x.size + 300],
                  [1000, 1001], [5000, 9000], [0, 0], [x.size, x.size + N]], np.int32)])
    got = engine.spec_batch(0, N, jobs)
    ref = oracle.spec_batch(x, N, jobs)
    assert rms(got, ref) < 1e-7 and np.abs(got - ref).max() < 1e-6


def test_empty_and_short_tracks(engine, oracle):
    for x in (np.zeros(0, np.float32), np.full(3, 0.25, np.float32)):
        engine.upload_tracks([x])
        jobs = np.array([[0, 256], [-100, 3], [1, 2]], np.int32)
        got = engine.spec_batch(0, 1024, jobs)
        ref = oracle.spec_batch(x, 1024, jobs) if x.size else np.zeros((3, 512), np.float32)
        assert np.abs(got - ref).max() < 1e-7
    assert engine.spec_batch(0, 1024, np.zeros((0, 2), np.int32)).shape == (0, 512)


def test_regular_hop_device_path(engine, oracle):
    import torch
    x = S.sine_sweep(3.0)
    engine.upload_tracks([x])
    F = (x.size + 511) // 512
    out = torch.empty((F, 1024), dtype=torch.float32, device="cuda")
    engine.use_torch_stream()
    engine.spec_frames_dev(0, 2048, 512, 0, F, out)
    torch.cuda.synchronize()
    assert rms(out.cpu().numpy(), oracle.spec_batch(x, 2048, S.regular_jobs(x.size, 512))) < 1e-7


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096, 8192])
def test_regular_hop_tiled_kernel(engine, oracle, N):
    """K1r (TMA-tiled regular-hop kernel): hops N/4, N and a 300-sample "pixel", first_frame > 0,
    frames that reach into the zero padding on both sides; against the oracle and against the
    general job-list kernel on the same jobs."""
    import torch
    x = S.vibrato_tone(1.5, seed=3 * N)
    engine.upload_tracks([x])
    engine.use_torch_stream()
    for hop, first, count in ((N // 4, 0, None), (N, 0, None), (300, 0, None), (N // 4, 7, None), (4, 1000, 333),
                              (N - 4, 3, 5), (N // 4, 40, 1)):
        if hop > N:
            continue
        F = count or (x.size + hop - 1) // hop + 5 - first  # five frames past the end of the track
        out = torch.empty((F, N // 2), dtype=torch.float32, device="cuda")
        engine.spec_frames_dev(0, N, hop, first, F, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        f = np.arange(first, first + F)
        jobs = np.stack([f * hop, (f + 1) * hop], 1).astype(np.int32)
        ref = oracle.spec_batch(x, N, jobs)
        assert rms(got, ref) < 1e-7 and np.abs(got - ref).max() < 1e-6, (N, hop, first)
        perm = np.random.default_rng(N + hop).permutation(F)  # a shuffled list is not a regular run
        gen = engine.spec_batch(0, N, jobs[perm])
        assert np.abs(gen - got[perm]).max() < 1e-6, (N, hop, first)


def test_regular_hop_many_batches(engine, oracle):
    """K1r with several double-buffered tile batches per CTA (120 s, 1024/256: 22 500 frames) through
    the host job-list entry point, which recognises the regular run."""
    x = S.vibrato_tone(120.0, seed=77)
    jobs = S.regular_jobs(x.size, 256)
    engine.upload_tracks([x])
    got = engine.spec_batch(0, 1024, jobs)
    ref = oracle.spec_batch(x, 1024, jobs)
    assert got.shape == ref.shape == (22500, 512)
    assert rms(got, ref) < 1e-7 and np.abs(got - ref).max() < 1e-6
    rgb = engine.spec_batch_rgb(0, 1024, jobs[:4000], 2.0 ** 12).astype(np.int32)
    rref = oracle.colormap(ref[:4000], 2.0 ** 12).astype(np.int32)
    d = np.abs(rgb - rref)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_fused_colour_ramp(engine, oracle):
    """K7: RGB texels vs the oracle's restatement of spec-cache.cpp:77-96 applied to the oracle's
    spectrum.  Casts truncate, so a 1e-7 relative magnitude difference can move a texel by one."""
    x = S.vibrato_tone(1.0, seed=11)
    engine.upload_tracks([x])
    jobs = S.regular_jobs(x.size, 375)[:32]
    seen = set()
    for k in (2.0 ** 15, 2.0 ** 13, 2.0 ** 11):
        got = engine.spec_batch_rgb(0, 32768, jobs, k).astype(np.int32)
        ref = oracle.colormap(oracle.spec_batch(x, 32768, jobs), k).astype(np.int32)
        diff = np.abs(got - ref)
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
        seen |= {"low"} if (ref[..., 1] == 0).any() else set()
        seen |= {"mid"} if ((ref[..., 1] > 0) & (ref[..., 2] == 0)).any() else set()
        seen |= {"high"} if (ref[..., 2] > 0).any() else set()
    assert seen == {"low", "mid", "high"}  # all three segments of the ramp are exercised
    c = np.load(GOLD / "colormap.npz")
    assert c["rgb"].shape[1] == 3


def test_errors_are_reported(engine):
    import melonix_b200 as m
    engine.upload_tracks([np.zeros(100, np.float32)])
    with pytest.raises(m.MlxError):
        engine.spec_batch(0, 1000, np.array([[0, 10]], np.int32))       # not a power of two
    with pytest.raises(m.MlxError):
        engine.spec_batch(3, 1024, np.array([[0, 10]], np.int32))       # no such track
    with pytest.raises(m.MlxError):
        engine.pv_run(2048, 500, 1.0)                                   # hop != N/4


def test_all_tracks_in_one_launch_equals_per_track_launches(engine, oracle):
    """mlx_spec_frames_all_dev (one batched K1r launch over every uploaded track, ragged lengths) against
    the per-track launches bit for bit, and against the oracle (reference spec.cpp:44-66 semantics)."""
    import torch
    xs = [S.vibrato_tone(2.0, seed=91), S.vibrato_tone(0.7, seed=92), S.sine_sweep(1.3), np.zeros(300, np.float32)]
    engine.upload_tracks(xs)
    engine.use_torch_stream()
    for N, hop in ((1024, 256), (2048, 512), (512, 128), (4096, 1024), (2048, 300)):
        Fs = [(x.size + hop - 1) // hop for x in xs]
        outs = [torch.full((F, N // 2), -1.0, dtype=torch.float32, device="cuda") for F in Fs]
        engine.spec_frames_all_dev(N, hop, outs)
        torch.cuda.synchronize()
        for t, (x, F) in enumerate(zip(xs, Fs)):
            one = torch.full((F, N // 2), -2.0, dtype=torch.float32, device="cuda")
            engine.spec_frames_dev(t, N, hop, 0, F, one)
            torch.cuda.synchronize()
            assert torch.equal(outs[t], one), (N, hop, t)
            ref = oracle.spec_batch(x, N, S.regular_jobs(x.size, hop), nthreads=2)
            assert np.sqrt(np.mean((outs[t].cpu().numpy().astype(np.float64) - ref) ** 2)) < 1e-7
