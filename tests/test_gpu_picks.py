"""-m gpu: waveform min/max pyramid (K9) through the C ABI against the oracle's restatement of
App::calcPicks / App::getMinMaxFromRange (reference app.cpp:347-426).  min / max of floats: bit-exact,
including the std::min / std::max argument order that decides NaN and signed-zero cases."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _ranges(n, rng, count=4000):
    """random ranges of every size class plus the reference's edge cases (app.cpp:382-396)."""
    if n <= 0:
        return np.array([[0, 0], [0, 1], [-1, 3], [2, 1]], np.int32)
    lens = (2.0 ** rng.uniform(0, np.log2(max(n, 2)), count)).astype(np.int64)
    start = rng.integers(0, max(n - 1, 1), count)
    end = np.minimum(start + lens, n - 1)
    edge = np.array([[0, 0], [5, 5], [7, 3], [n, n], [n - 1, n - 1], [-3, 10], [10, -3], [0, n], [0, n - 1],
                     [n - 2, n - 1], [n - 1, n], [n + 5, n + 9], [0, 1], [0, 2], [1, 3], [3, 4], [0, min(n - 1, 4096)],
                     [1, min(n - 1, 4097)], [min(5, n - 1), n - 1]], np.int64)
    return np.concatenate([np.stack([start, end], 1), edge]).astype(np.int32)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 16, 17, 4095, 4096, 4097, 8191, 8193, 100003, 1_440_000])
def test_pyramid_and_range_queries_bit_exact(engine, oracle, n):
    rng = np.random.default_rng(n + 1)
    x = (S.vibrato_tone(max(n, 1) / 48000.0 + 0.01, seed=n)[:n] if n else np.zeros(0, np.float32)).astype(np.float32)
    assert x.size == n
    engine.upload_tracks([x])
    pairs, off = engine.picks_build(0)
    ref, roff = oracle.picks_build(x)
    assert np.array_equal(off, roff) and pairs.shape == ref.shape
    assert np.array_equal(bits(pairs), bits(ref))
    r = _ranges(n, rng)
    got = engine.minmax_ranges(0, r)
    exp = oracle.minmax_ranges(x, ref, roff, r)
    assert np.array_equal(bits(got), bits(exp))


def test_nan_and_signed_zero_follow_std_min_max(engine, oracle):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(70001).astype(np.float32)
    x[::7] = 0.0
    x[3::7] = -0.0
    x[5::101] = np.nan
    engine.upload_tracks([x, x[:5000]])
    for t, xx in ((0, x), (1, x[:5000])):
        pairs, off = engine.picks_build(t)
        ref, roff = oracle.picks_build(xx)
        assert np.array_equal(bits(pairs), bits(ref))
        r = _ranges(xx.size, rng, 2000)
        assert np.array_equal(bits(engine.minmax_ranges(t, r)), bits(oracle.minmax_ranges(xx, ref, roff, r)))


def test_device_entry_and_cache_invalidation(engine, oracle):
    import torch
    x = S.two_tone(2.0)
    engine.upload_tracks([x])
    engine.use_torch_stream()
    off = engine.picks_layout(x.size)
    ref, _ = oracle.picks_build(x)
    for shift in (0, 1):                            # a caller buffer that is only 8-byte aligned works too
        buf = torch.zeros((int(off[-1]) + 1, 2), dtype=torch.float32, device="cuda")
        engine.picks_build_dev(0, buf[shift:])
        torch.cuda.synchronize()
        assert np.array_equal(bits(buf[shift:shift + int(off[-1])].cpu().numpy()), bits(ref))
    # all uploaded tracks in one launch (ragged lengths)
    tracks = [x, x[:5001], np.zeros(0, np.float32), x[:70000] * 0.5]
    engine.upload_tracks(tracks)
    bufs = [torch.zeros((max(int(engine.picks_layout(t.size)[-1]), 1), 2), dtype=torch.float32, device="cuda")
            for t in tracks]
    engine.picks_build_all_dev(bufs)
    torch.cuda.synchronize()
    for t, b in zip(tracks, bufs):
        rt, ro = oracle.picks_build(t)
        assert np.array_equal(bits(b[:int(ro[-1])].cpu().numpy()), bits(rt))
    engine.upload_tracks([x])
    r = np.array([[100, 90000], [0, 95999]], np.int32)
    a = engine.minmax_ranges(0, r)
    y = (0.5 * x + 0.1).astype(np.float32)
    engine.upload_tracks([y])                       # a new upload must not reuse the cached pyramid
    b = engine.minmax_ranges(0, r)
    ry, ro = oracle.picks_build(y)
    assert np.array_equal(bits(b), bits(oracle.minmax_ranges(y, ry, ro, r))) and not np.array_equal(a, b)


def test_golden_fixture(engine):
    """tests/golden/picks_5003.npz (made by tests/golden/make_golden.py from the oracle)."""
    g = np.load(Path(__file__).resolve().parent / "golden" / "picks_5003.npz")
    x = S.vibrato_tone(5003 / 48000.0 + 0.01, seed=int(g["seed"]))[:int(g["n"])]
    engine.upload_tracks([x])
    pairs, off = engine.picks_build(0)
    assert np.array_equal(off, g["level_off"]) and np.array_equal(bits(pairs), bits(g["pairs"]))
    assert np.array_equal(bits(engine.minmax_ranges(0, g["ranges"])), bits(g["minmax"]))
