"""CPU tests of the oracle (oracle/): pinned against the reference's own spec.cpp compiled
unmodified (oracle/_ref), the analytic KATs of SURVEY.md section 4, an independent numpy
restatement of the PV spec, and the committed golden vectors."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402
from np_pv_reference import pv_numpy  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


def test_fft_against_numpy(oracle):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(3000).astype(np.float32)
    jobs = np.array([[1000, 1256], [-5, 100], [2900, 3100], [-3000, -100], [7000, 7256]], np.int32)
    for N in (512, 1024, 4096):
        got = oracle.spec_batch(x, N, jobs, nthreads=1)
        for j, (a, b) in enumerate(jobs):
            i = np.arange(b - N, b)
            ok = (i >= 0) & (i < x.size)
            xi = np.where(ok, x[np.clip(i, 0, x.size - 1)], 0).astype(np.float32)
            w = np.where(i >= a, np.float32(1), np.exp(np.float32(-2.5e-4) * (a - i).astype(np.float32)).astype(np.float32))
            ref = np.abs(np.fft.fft((w * xi).astype(np.float32).astype(np.float64)))[:N // 2] / N
            assert np.abs(got[j] - ref).max() < 5e-8


def test_kat1_spec_peak(oracle):
    # analytic peak of the reference window
    n = 200000
    wav = (0.5 * np.sin(2 * np.pi * 300 * np.arange(n) / 32768)).astype(np.float32)
    s = oracle.spec_batch(wav, 32768, np.array([[100000, 100375]], np.int32))[0]
    assert s.size == 16384 and s.argmax() == 300
    assert abs(float(s.max()) - 0.03340026) < 1e-7
    a = 2.5e-4
    W = 375 + np.exp(-a) * (1 - np.exp(-a * (32768 - 375))) / (1 - np.exp(-a))
    assert abs(float(s.max()) - 0.25 * W / 32768) < 1e-4  # closed form is approximate (leakage)


def test_kat2_spec_edges(oracle):
    x = S.sine_sweep(0.2)
    jobs = np.array([[-5000, 0], [-100, -1], [x.size + 1024, x.size + 1280]], np.int32)
    assert not oracle.spec_batch(x, 1024, jobs).any()


def test_restatement_equals_reference_spec_cpp(oracle):
    """oracle/spec_ref.c vs the reference's own Spec class (spec.cpp compiled unmodified)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    x = S.vibrato_tone(1.0, seed=3)
    jobs = np.array([[0, 375], [20000, 20375], [-200, 175], [x.size - 100, x.size + 275], [40000, 41000]], np.int32)
    ref = oracle.ref_spec_run(x, jobs)
    got = oracle.spec_batch(x, 32768, jobs, nthreads=2)
    assert np.array_equal(ref, got)


def test_golden_spec(oracle):
    g = np.load(GOLD / "spec_ref_geometry.npz")
    n = int(g["n"])
    kat = (0.5 * np.sin(2 * np.pi * 300 * np.arange(n) / 32768)).astype(np.float32)
    assert np.array_equal(oracle.spec_batch(kat, 32768, g["jobs"]), g["out"])  # golden = reference output
    g = np.load(GOLD / "spec_cfg1.npz")
    x = S.sine_sweep(float(g["seconds"]))
    assert np.array_equal(oracle.spec_batch(x, 1024, g["jobs"]), g["out"])


@pytest.mark.parametrize("N,semis", [(1024, 0.0), (2048, 3.0), (512, -5.0), (4096, -12.0), (512, 11.0)])
def test_pv_oracle_against_numpy_restatement(oracle, N, semis):
    fs = 48000
    t = np.arange(fs // 3) / fs
    rng = np.random.default_rng(3)
    x = (0.4 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 1230 * t + 1) + 1e-3 * rng.standard_normal(t.size)).astype(np.float32)
    r = np.float32(2.0) ** (np.float32(semis) / np.float32(12.0))
    o = oracle.pv_run(x, N, N // 4, r)
    y, pk, f0 = pv_numpy(x, N, N // 4, r)
    assert np.sqrt(np.mean((o["y"].astype(np.float64) - y) ** 2)) < 1e-9
    assert np.array_equal(pk, o["peak"])
    assert np.abs(f0 - o["f0"]).max() < 1e-3


def test_pv_identity_and_pitch(oracle):
    fs = 48000
    x = (0.5 * np.sin(2 * np.pi * 440 * np.arange(fs) / fs)).astype(np.float32)
    o = oracle.pv_run(x, 2048, 512, 1.0)
    i = slice(2048, x.size - 2048)
    assert np.sqrt(np.mean((o["y"][i] - x[i]) ** 2)) < 1e-6          # r = 1 reproduces the input
    assert set(o["peak"][8:-8]) == {19}                               # 440 Hz -> bin 19 at 2048/48k
    assert abs(np.median(o["f0"][8:-8]) - 440.0) < 0.01
    r = np.float32(2.0) ** (np.float32(3.0) / np.float32(12.0))
    y = oracle.pv_run(x, 2048, 512, r)["y"].astype(np.float64)
    spec = np.abs(np.fft.rfft(y[8192:8192 + 32768] * np.hanning(32768)))
    assert abs(spec.argmax() * fs / 32768 - 440 * 2 ** 0.25) < 3.0     # shifted to ~523 Hz


def test_golden_pv(oracle):
    for name, N in (("pv_2048_p3.npz", 2048), ("pv_1024_m5.npz", 1024)):
        g = np.load(GOLD / name)
        x = S.vibrato_tone(float(g["seconds"]), seed=int(g["seed"]))
        o = oracle.pv_run(x, N, N // 4, float(g["rate"]))
        assert np.array_equal(o["y"], g["y"]) and np.array_equal(o["peak"], g["peak"])
        assert np.allclose(o["f0"], g["f0"], rtol=0, atol=1e-4)


def test_kat4_kat5_grains(oracle):
    x = S.two_tone(20.0)
    gs, gl = oracle.grain_segment(x)
    assert gs.size == 628 and gl.min() >= 1519 and gl.max() <= 1528   # SURVEY.md KAT-4 probe
    e = oracle.grain_export(x, 48000, [], gs, gl)
    total = int(gl.sum())
    assert e["pcm"].size == total + 1500
    assert np.array_equal(e["pcm"][:total], x[:total]) and not e["pcm"][total:].any()
    for semis, steps in ((3.0, 746), (-3.0, 528)):                    # KAT-5: duration preserved
        mk = [(10, 0, 0, semis), (x.size - 10, 0, 0, semis)]
        e = oracle.grain_export(x, 48000, mk, gs, gl)
        assert abs(e["pcm"].size - x.size) < 2300
        assert e["schedule"]["rate"].size == steps


def test_golden_grain_and_colormap(oracle):
    g = np.load(GOLD / "grain_p3.npz")
    x = S.two_tone(float(g["seconds"]))
    gs, gl = oracle.grain_segment(x)
    assert np.array_equal(gs, g["g_start"]) and np.array_equal(gl, g["g_len"])
    e = oracle.grain_export(x, 48000, [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)], gs, gl)
    assert np.array_equal(e["pcm16"], g["pcm16"]) and np.array_equal(e["pcm"], g["pcm"])
    c = np.load(GOLD / "colormap.npz")
    assert np.array_equal(oracle.colormap(c["v"], float(c["k"])), c["rgb"])


def test_colormap_matches_numpy_restatement(oracle):
    v = np.linspace(0, 3, 997).astype(np.float32)
    k = np.float32(100.0)
    tmp = np.clip(v * k, np.float32(0), np.float32(255)).astype(np.float32)
    exp = np.zeros((v.size, 3), np.uint8)
    for i, t in enumerate(tmp):
        if t < 85:
            exp[i] = (int(t), 0, 0)
        elif t < 170:
            a = float(np.float32(np.float32(t - np.float32(85)) / np.float32(85))) * 3.141592 / 2
            exp[i] = (int(float(t) * np.cos(a)), int(float(t) * np.sin(a)), 0)
        else:
            lk = int(np.float32(np.float32(t - np.float32(170)) * np.float32(3)))
            exp[i] = (lk, int(t), lk)
    assert np.array_equal(oracle.colormap(v, float(k)), exp)


def test_picks_oracle_against_brute_force(oracle):
    """Restatement of App::calcPicks / getMinMaxFromRange (app.cpp:347-426): every level equals the
    brute-force min/max of its blocks, aligned range queries are exact, and the reference's edge
    behaviour (empty / out-of-range ranges, app.cpp:382-396) is reproduced."""
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 3, 4, 5, 8, 9, 1000, 4097, 100003):
        x = rng.standard_normal(n).astype(np.float32)
        pairs, off = oracle.picks_build(x)
        L = len(off) - 1
        assert L == sum(1 for l in range(40) if n > (1 << (l + 1)))
        for l in range(L):
            cnt = n >> (l + 1)
            assert off[l + 1] - off[l] == cnt
            blk = x[:cnt << (l + 1)].reshape(cnt, -1)
            assert np.array_equal(pairs[off[l]:off[l + 1], 0], blk.min(1))
            assert np.array_equal(pairs[off[l]:off[l + 1], 1], blk.max(1))
        if n > 16:
            r = []
            for _ in range(200):
                l = int(rng.integers(1, int(np.log2(n))))
                i = int(rng.integers(0, n >> l))
                if ((i + 1) << l) < n:
                    r.append((i << l, (i + 1) << l))
            r = np.array(r, np.int32)
            bf = np.array([[x[s:e].min(), x[s:e].max()] for s, e in r], np.float32)
            assert np.array_equal(oracle.minmax_ranges(x, pairs, off, r), bf)
            edge = np.array([[5, 5], [7, 3], [n, n], [-3, 10], [10, -3], [0, n], [3, 4]], np.int32)
            e = oracle.minmax_ranges(x, pairs, off, edge)
            assert np.array_equal(e[0], [x[5], x[5]]) and np.array_equal(e[1], [x[7], x[7]])
            assert not e[2].any() and not e[3].any() and not e[5].any()
            assert np.array_equal(e[4], [x[10], x[10]])   # start >= end is tested first (app.cpp:382)
            assert np.array_equal(e[6], [x[3], x[3]])


def test_golden_picks(oracle):
    g = np.load(GOLD / "picks_5003.npz")
    x = S.vibrato_tone(5003 / 48000.0 + 0.01, seed=int(g["seed"]))[:int(g["n"])]
    pairs, off = oracle.picks_build(x)
    assert np.array_equal(off, g["level_off"]) and np.array_equal(pairs.view(np.uint32), g["pairs"].view(np.uint32))
    mm = oracle.minmax_ranges(x, pairs, off, g["ranges"])
    assert np.array_equal(mm.view(np.uint32), g["minmax"].view(np.uint32))


# ------------------------------------------------------------------------------------------------
# The oracle against the reference's OWN application code (oracle/_ref/libapp_ref.so: app.cpp,
# spec-cache.cpp, save-wav.cpp, spec.cpp compiled unmodified; see oracle/ref_app.cpp).
needs_ref_app = pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).have_ref_app(),
                                   reason="oracle/_ref/libapp_ref.so not built (needs /root/reference)")

MARKER_SETS = [
    [],
    [(10, 0, 0, 3.0), (-10, 0, 0, 3.0)],                                            # constant +3 st
    [(10, 0, 0, -3.0), (-10, 0, 0, -3.0)],
    [(100000, 0, 0.5, 2.0), (200000, 0, -0.3, -1.5), (280000, 0, 0.0, 4.0)],        # time warp + pitch bends
]


def _mk(markers, n):
    return [(m[0] if m[0] >= 0 else n + m[0], m[1], m[2], m[3]) for m in markers]


@needs_ref_app
@pytest.mark.parametrize("markers", MARKER_SETS)
def test_grain_path_pinned_by_reference_app(oracle, markers, tmp_path):
    """App::preproc grains, App::process float output, App::exportWav int16 output and the warp maps of the
    reference itself equal the oracle's restatements (oracle/grain_ref.c) bit for bit."""
    x = S.two_tone(6.0)
    mk = _mk(markers, x.size)
    with oracle.RefApp(x, 48000, mk) as app:
        gs, gl = app.grains()
        os_, ol = oracle.grain_segment(x)
        assert np.array_equal(gs, os_) and np.array_equal(gl, ol)
        o = oracle.grain_export(x, 48000, mk)
        pcm = app.render()
        assert pcm.size == o["pcm"].size and np.array_equal(pcm.view(np.uint32), o["pcm"].view(np.uint32))
        pcm16 = app.export_wav(tmp_path / "out.wav")
        assert pcm16.size == o["pcm16"].size
        assert np.array_equal(pcm16[2:], o["pcm16"][2:])      # samples 0-1: clobbered by the reference's saveWav
        assert not pcm16[:2].any()
        for t in np.linspace(-0.5, 7.0, 301):
            assert app.time2sample(t) == oracle.time2sample(mk, 48000, float(t))
            assert np.float32(app.time2pitchbend(t)) == np.float32(oracle.time2pitchbend(mk, 48000, x.size, float(t)))
        for s_ in range(-100, x.size, 2999):
            assert app.sample2time(s_) == oracle.sample2time(mk, 48000, s_)
        assert app.duration() == oracle.sample2time(mk, 48000, x.size - 1)


@needs_ref_app
def test_grain_segmentation_pinned_on_hard_signals(oracle):
    """look-3 fallback, noise, silence, short clips, -0.0 / NaN: the reference's own loop vs the oracle."""
    fs = 48000
    rng = np.random.default_rng(8)
    t = np.arange(20 * fs) / fs
    weird = (0.4 * np.sin(2 * np.pi * 440 * t[:5 * fs])).astype(np.float32)
    weird[::97] = -0.0
    weird[5::1013] = np.nan
    cases = [
        (0.4 * np.sin(2 * np.pi * 9 * t)).astype(np.float32),
        (0.4 * np.sin(2 * np.pi * 33 * t) + 0.02 * (rng.random(t.size) - 0.5)).astype(np.float32),
        (rng.random(5 * fs) - 0.5).astype(np.float32),
        np.zeros(100000, np.float32), np.full(50000, -0.25, np.float32),
        S.two_tone(1502 / fs)[:1502], S.two_tone(3100 / fs)[:3100], weird,
    ]
    for x in cases:
        with oracle.RefApp(x, fs, []) as app:
            gs, gl = app.grains()
        os_, ol = oracle.grain_segment(x)
        assert np.array_equal(gs, os_) and np.array_equal(gl, ol)


@needs_ref_app
def test_picks_pinned_by_reference_app(oracle):
    rng = np.random.default_rng(5)
    for n in (3, 9, 1000, 4097, 288000):
        x = rng.standard_normal(n).astype(np.float32)
        if n > 50:
            x[::7] = 0.0
            x[3::7] = -0.0
            x[5::101] = np.nan
        with oracle.RefApp(x, 48000, []) as app:
            rp, roff = app.picks()
            pairs, off = oracle.picks_build(x)
            assert np.array_equal(roff, off) and np.array_equal(rp.view(np.uint32), pairs.view(np.uint32))
            s_ = rng.integers(0, max(n - 1, 1), 3000)
            e_ = np.minimum(s_ + (2.0 ** rng.uniform(0, np.log2(n), 3000)).astype(np.int64), n - 1)
            r = np.concatenate([np.stack([s_, e_], 1), [[5, 5], [7, 3], [n, n], [-3, 10], [10, -3], [0, n], [0, n - 1]]])
            r = r.astype(np.int32)
            assert np.array_equal(app.minmax_ranges(r).view(np.uint32),
                                  oracle.minmax_ranges(x, pairs, off, r).view(np.uint32))


@needs_ref_app
def test_colour_ramp_pinned_by_reference_speccache(oracle):
    """SpecCache::populateTex (spec-cache.cpp:52-110) of the reference: the texels it uploads equal
    mlxo_colormap applied to the oracle spectrum of the same job, byte for byte, in all three segments."""
    x = S.vibrato_tone(3.0, seed=11)
    seen = set()
    with oracle.RefApp(x, 48000, []) as app:
        for k in (2.0 ** 15, 2.0 ** 13, 2.0 ** 11):
            for t in (0.5, 1.7, 2.4):
                width, range_time = 1280, 10.0
                rgb = app.speccache_column(k, width, range_time, t)
                key = int(t * width / range_time)                         # spec-cache.cpp:12
                start, px = key * range_time / width, range_time / width  # spec-cache.cpp:63-64
                job = np.array([[oracle.time2sample([], 48000, start), oracle.time2sample([], 48000, start + px)]],
                               np.int32)
                ref = oracle.colormap(oracle.spec_batch(x, 32768, job), k)[0]
                assert np.array_equal(rgb, ref)
                seen |= {"low"} if (ref[..., 1] == 0).any() else set()
                seen |= {"mid"} if ((ref[..., 1] > 0) & (ref[..., 2] == 0)).any() else set()
                seen |= {"high"} if (ref[..., 2] > 0).any() else set()
    assert seen == {"low", "mid", "high"}


@needs_ref_app
def test_golden_fixtures_equal_reference_outputs(oracle, tmp_path):
    """The committed grain / picks fixtures are what the reference's own code produces."""
    g = np.load(GOLD / "grain_p3.npz")
    x = S.two_tone(float(g["seconds"]))
    mk = [(10, 0.0, 0.0, 3.0), (x.size - 10, 0.0, 0.0, 3.0)]
    with oracle.RefApp(x, 48000, mk) as app:
        gs, gl = app.grains()
        assert np.array_equal(gs, g["g_start"]) and np.array_equal(gl, g["g_len"])
        assert np.array_equal(app.render().view(np.uint32), g["pcm"].view(np.uint32))
        assert np.array_equal(app.export_wav(tmp_path / "g.wav")[2:], g["pcm16"][2:])
    p = np.load(GOLD / "picks_5003.npz")
    x = S.vibrato_tone(5003 / 48000.0 + 0.01, seed=int(p["seed"]))[:int(p["n"])]
    with oracle.RefApp(x, 48000, []) as app:
        rp, roff = app.picks()
        assert np.array_equal(roff, p["level_off"]) and np.array_equal(rp.view(np.uint32), p["pairs"].view(np.uint32))
        assert np.array_equal(app.minmax_ranges(p["ranges"]).view(np.uint32), p["minmax"].view(np.uint32))


@needs_ref_app
def test_random_marker_sets_reference_oracle_and_host_schedule(oracle):
    """40 random marker sets (time warp dTime in +-0.4 s, pitch bends in +-12 st, 1-6 markers): the
    reference's own process() loop, the oracle's restatement and the product's host schedule builder
    (melonix_b200/host/grain_schedule.cpp, which feeds mlx_grain_render) agree exactly."""
    from melonix_b200 import hostlib as H
    x = S.two_tone(4.0)
    gs, gl = oracle.grain_segment(x)
    hs, hl = H.grain_segment(x)
    assert np.array_equal(gs, hs) and np.array_equal(gl, hl)
    rng = np.random.default_rng(2024)
    for it in range(40):
        nm = int(rng.integers(1, 7))
        samples = np.sort(rng.choice(np.arange(2000, x.size - 2000), nm, replace=False))
        mk = [(int(s_), 0.0, float(np.round(rng.uniform(-0.4, 0.4), 3)), float(np.round(rng.uniform(-12, 12), 2)))
              for s_ in samples]
        o = oracle.grain_export(x, 48000, mk, gs, gl)
        with oracle.RefApp(x, 48000, mk) as app:
            pcm = app.render()
        assert pcm.size == o["pcm"].size, (it, mk)
        assert np.array_equal(pcm.view(np.uint32), o["pcm"].view(np.uint32)), (it, mk)
        s = H.export_schedule(x, 48000, mk, hs, hl)
        for k in ("gstart", "glen", "rate", "next"):
            assert np.array_equal(s[k], o["schedule"][k]), (it, k, mk)
        assert np.array_equal(s["out_off"][:-1], o["schedule"]["out_off"])
        assert s["out_off"][-1] + s["tail_zeros"] == o["pcm"].size
