"""-m gpu: phase-vocoder path (K_A / scan / K_S) through the C ABI against the double-precision
oracle (PV-spec v1; NOT IN REFERENCE -- parity unpinned by reference, self-consistency target).
Tolerances (BASELINE.json north_star): output RMS <= 1e-4 absolute; peak bins bit-exact on every
frame whose top-two magnitude margin exceeds 1e-5; f0 within 1e-3 Hz."""
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


def ratio(semis):
    import melonix_b200 as m
    return m.semitone_ratio(semis)


def check(g, o, tol=1e-4):
    ok = o["margin"] > 1e-5
    assert rms(g["y"], o["y"]) <= tol
    assert np.array_equal(g["peak"][ok], o["peak"][ok])
    assert np.abs(g["f0"][ok] - o["f0"][ok]).max() < 1e-3
    return int((~ok).sum())


def test_config2_full_pitch_shift(engine, oracle):
    """BASELINE configs[1]: 60 s mono, 2048-FFT / 512-hop, +3 semitones."""
    x = S.vibrato_tone(60.0, seed=1234)
    engine.upload_tracks([x])
    g = engine.pv_run(2048, 512, ratio(3.0))[0]
    o = oracle.pv_run(x, 2048, 512, ratio(3.0))
    assert g["peak"].size == 5625
    excluded = check(g, o)
    assert excluded == 0
    assert rms(g["y"], o["y"]) < 1e-6          # what the kernels actually achieve


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("semis", [3.0, -4.0, 0.0, 12.0, -24.0])
def test_sizes_and_ratios(engine, oracle, N, semis):
    x = S.vibrato_tone(3.0, seed=N + int(semis))
    engine.upload_tracks([x])
    g = engine.pv_run(N, N // 4, ratio(semis))[0]
    o = oracle.pv_run(x, N, N // 4, ratio(semis))
    check(g, o)


def test_golden_vectors(engine):
    for name, N in (("pv_2048_p3.npz", 2048), ("pv_1024_m5.npz", 1024)):
        gold = np.load(GOLD / name)
        x = S.vibrato_tone(float(gold["seconds"]), seed=int(gold["seed"]))
        engine.upload_tracks([x])
        g = engine.pv_run(N, N // 4, float(gold["rate"]))[0]
        ok = gold["margin"] > 1e-5
        assert rms(g["y"], gold["y"]) <= 1e-4
        assert np.array_equal(g["peak"][ok], gold["peak"][ok])


def test_ragged_batch_silence_and_tiny_tracks(engine, oracle):
    xs = [S.vibrato_tone(2.0, seed=1), S.vibrato_tone(1.37, seed=2, f_base=330.0), S.vibrato_tone(0.2, seed=3),
          np.zeros(30000, np.float32),                                # digital silence: the |Z| gate
          np.concatenate([np.zeros(9000, np.float32), S.vibrato_tone(0.5, seed=4), np.zeros(7000, np.float32)]),
          S.vibrato_tone(0.004, seed=5),                              # shorter than one hop
          np.full(1, 0.5, np.float32)]
    engine.upload_tracks(xs)
    r = ratio(3.0)
    gs = engine.pv_run(2048, 512, r)
    for x, g in zip(xs, gs):
        o = oracle.pv_run(x, 2048, 512, r)
        assert rms(g["y"], o["y"]) <= 1e-4
        ok = o["margin"] > 1e-5
        assert np.array_equal(g["peak"][ok], o["peak"][ok])
    assert not gs[3]["y"].any()


def test_identity_property_full_size(engine):
    """Size-independent property at BASELINE's full track length (5 min): r = 1 reproduces the input
    away from the two edges (PV-spec A.8)."""
    x = S.vibrato_tone(300.0, seed=77)
    engine.upload_tracks([x])
    g = engine.pv_run(2048, 512, 1.0)[0]
    i = slice(2048, x.size - 2048)
    assert rms(g["y"][i], x[i]) < 2e-6
    assert g["peak"].size == 28125


def test_no_drift_on_stationary_tones(engine, oracle):
    """A smooth per-frame phase error would grow linearly with time; the integer-turn phase
    difference makes it telescope.  300 s of stationary partials: error flat from start to end."""
    n = 300 * 48000
    t = np.arange(n) / 48000.0
    x = (0.3 * np.sin(2 * np.pi * 441.3 * t) + 0.15 * np.sin(2 * np.pi * 1237.7 * t + 0.3)
         + 1e-4 * np.random.default_rng(5).standard_normal(n)).astype(np.float32)
    engine.upload_tracks([x])
    g = engine.pv_run(2048, 512, ratio(3.0))[0]
    o = oracle.pv_run(x, 2048, 512, ratio(3.0))
    seg = 5 * 48000
    first, last = rms(g["y"][:seg], o["y"][:seg]), rms(g["y"][-seg:], o["y"][-seg:])
    assert first < 1e-6 and last < 1e-6 and last < 3 * first + 1e-7
    assert np.array_equal(g["peak"], o["peak"])


def test_cut_decision_on_noise(engine, oracle):
    """White noise puts every bin's phase advance uniformly on the circle: hundreds of full-scale
    bin-frames fall within 1.5e-3 rad of the +-pi cut, where K_A takes the sign from the DOUBLE product
    X conj(Xprev) (-i)^k.  A wrong sign there offsets that bin's phase by frac(rate) turns for the
    rest of the track, so the RMS bound checks the cut logic on all of them."""
    N, H = 2048, 512
    n = 48000 * 6
    x = (0.1 * np.random.default_rng(2024).standard_normal(n)).astype(np.float32)
    F = (n + H - 1) // H
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32).astype(np.float64)
    xp = np.concatenate([np.zeros(N), x.astype(np.float64), np.zeros(N)])
    X = np.stack([np.fft.rfft(xp[(f + 1) * H:(f + 1) * H + N] * w) for f in range(F)])
    Z = X[1:] * np.conj(X[:-1]) * ((-1j) ** np.arange(N // 2 + 1))
    near = (np.abs(np.abs(np.angle(Z)) - np.pi) < 1.5e-3) & (np.abs(Z) > 1.0)
    assert near.sum() >= 100
    engine.upload_tracks([x])
    for semis in (3.0, -7.0):
        g = engine.pv_run(N, H, ratio(semis))[0]
        o = oracle.pv_run(x, N, H, ratio(semis))
        assert rms(g["y"], o["y"]) <= 1e-4 and rms(g["y"], o["y"]) < 1e-6


def test_wave_tiling_and_chunking_are_bitwise_invisible(engine, monkeypatch):
    xs = [S.vibrato_tone(6.0, seed=21), S.vibrato_tone(4.3, seed=22)]
    engine.upload_tracks(xs)
    r = ratio(-2.0)
    a = engine.pv_run(2048, 512, r, wave_mib=-1)
    b = engine.pv_run(2048, 512, r, wave_mib=1)
    monkeypatch.setenv("MLX_PV_CHUNK", "24")
    c = engine.pv_run(2048, 512, r, wave_mib=2)
    for u, v, w in zip(a, b, c):
        assert np.array_equal(u["y"], v["y"]) and np.array_equal(u["y"], w["y"])
        assert np.array_equal(u["peak"], v["peak"]) and np.array_equal(u["f0"], w["f0"])


def test_time_range_shards_equal_unsharded_bitwise(engine):
    """S logical shards run one after the other on one GPU (seam exchange =
    slicing, phase carry = prefix of the shard totals) must equal the unsharded result bit for bit."""
    import torch
    from melonix_b200 import dist as D
    N, H = 4096, 1024
    x = S.vibrato_tone(20.0, seed=31)
    r = ratio(3.0)
    engine.upload_tracks([x])
    full = engine.pv_run(N, H, r)[0]
    xd = torch.from_numpy(x).cuda()
    engine.use_torch_stream()
    for world in (2, 5):
        shards = D.plan_time_shards(x.size, N, H, world)
        carry = torch.zeros(N // 2 + 1, dtype=torch.int64, device="cuda")
        y = np.zeros_like(x)
        peak = np.zeros_like(full["peak"])
        for s in shards:
            win = xd[s.need_lo:s.need_hi].contiguous()
            engine.upload_tracks_dev([win])
            tot = torch.zeros(N // 2 + 1, dtype=torch.int32, device="cuda")
            engine.pv_phase_totals_dev(N, H, r, [tot], frame_begin=s.local_frame_begin, frame_end=s.local_frame_end,
                                       wave_mib=-1)
            c32 = torch.where(carry >= 2 ** 31, carry - 2 ** 32, carry).to(torch.int32)
            yd = torch.zeros_like(win)
            pk = torch.zeros(D.num_frames(win.numel(), H), dtype=torch.int32, device="cuda")
            engine.pv_run_dev(N, H, r, [yd], [pk], None, frame_begin=s.local_frame_begin,
                              frame_end=s.local_frame_end, phase_in=[c32], wave_mib=-1)
            torch.cuda.synchronize()
            y[s.own_lo:s.own_hi] = yd[s.left_halo:s.left_halo + (s.own_hi - s.own_lo)].cpu().numpy()
            peak[s.frame_begin:s.frame_end] = pk[s.local_frame_begin:s.local_frame_end].cpu().numpy()
            carry = (carry + (tot.to(torch.int64) & 0xFFFFFFFF)) & 0xFFFFFFFF
        assert np.array_equal(y, full["y"]), f"world={world}"
        assert np.array_equal(peak, full["peak"])


def test_split_analyze_synth_is_bitwise_the_one_call_pipeline(engine):
    """mlx_pv_analyze_dev + mlx_pv_synth_dev (ONE analysis pass, intermediates staged on the device) equal
    mlx_pv_run_dev bit for bit -- whole tracks, a ragged batch, and logical time shards with a carried-in
    phase (the configs[3] protocol without the second analysis pass round 1 needed)."""
    import torch
    from melonix_b200 import dist as D
    import melonix_b200 as m
    engine.use_torch_stream()
    # (a) whole tracks, ragged batch, pitch up and down
    xs = [S.vibrato_tone(5.0, seed=61), S.vibrato_tone(3.3, seed=62)]
    for N, semis in ((2048, 3.0), (2048, -5.0), (4096, 3.0), (1024, 7.0)):
        H = N // 4
        r = ratio(semis)
        engine.upload_tracks(xs)
        ref = engine.pv_run(N, H, r)
        nb = N // 2 + 1
        tots = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in xs]
        pks = [torch.zeros((x.size + H - 1) // H, dtype=torch.int32, device="cuda") for x in xs]
        f0s = [torch.zeros((x.size + H - 1) // H, dtype=torch.float32, device="cuda") for x in xs]
        ys = [torch.zeros(x.size, dtype=torch.float32, device="cuda") for x in xs]
        engine.pv_analyze_dev(N, H, r, tots, pks, f0s)
        engine.pv_synth_dev(N, H, r, ys)
        torch.cuda.synchronize()
        for u, y, pk, f0 in zip(ref, ys, pks, f0s):
            assert np.array_equal(u["y"], y.cpu().numpy()), (N, semis)
            assert np.array_equal(u["peak"], pk.cpu().numpy()) and np.array_equal(u["f0"], f0.cpu().numpy())
        # the totals equal those of the analysis-only entry point
        t2 = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in xs]
        engine.pv_phase_totals_dev(N, H, r, t2, wave_mib=-1)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(tots, t2))
    # (b) synth without a staged analysis must fail loudly, not synthesise stale data
    engine.upload_tracks(xs)
    with pytest.raises(m.MlxError):
        engine.pv_synth_dev(2048, 512, ratio(3.0), [torch.zeros(x.size, dtype=torch.float32, device="cuda") for x in xs])
    # (c) logical shards on one GPU
    N, H = 4096, 1024
    x = S.vibrato_tone(20.0, seed=31)
    r = ratio(3.0)
    engine.upload_tracks([x])
    full = engine.pv_run(N, H, r)[0]
    xd = torch.from_numpy(x).cuda()
    for world in (3,):
        carry = torch.zeros(N // 2 + 1, dtype=torch.int64, device="cuda")
        y = np.zeros_like(x)
        for s in D.plan_time_shards(x.size, N, H, world):
            win = xd[s.need_lo:s.need_hi].contiguous()
            engine.upload_tracks_dev([win])
            tot = torch.zeros(N // 2 + 1, dtype=torch.int32, device="cuda")
            engine.pv_analyze_dev(N, H, r, [tot], frame_begin=s.local_frame_begin, frame_end=s.local_frame_end)
            c32 = torch.where(carry >= 2 ** 31, carry - 2 ** 32, carry).to(torch.int32)
            yd = torch.zeros_like(win)
            engine.pv_synth_dev(N, H, r, [yd], frame_begin=s.local_frame_begin, frame_end=s.local_frame_end,
                                phase_in=[c32])
            torch.cuda.synchronize()
            y[s.own_lo:s.own_hi] = yd[s.left_halo:s.left_halo + (s.own_hi - s.own_lo)].cpu().numpy()
            carry = (carry + (tot.to(torch.int64) & 0xFFFFFFFF)) & 0xFFFFFFFF
        assert np.array_equal(y, full["y"]), f"world={world}"


def test_bench_tracks_full_length_against_oracle(engine, oracle):
    """Two real bench tracks (BASELINE configs[2] as bench.py generates them on the device: 300 s, per-track
    f_base, seeds 1234 + track): tracks 0 and 63 against the oracle at full length."""
    import torch
    import bench as B
    dev = torch.device("cuda", 0)
    n = 300 * 48000
    r = ratio(3.0)
    for track in (0, 63):
        # gen_tracks_gpu(rank = track, ntracks = 1) produces global track `track` of the bench batch
        xt = B.gen_tracks_gpu(torch, dev, 1, n, track)[0].cpu().numpy()
        engine.upload_tracks([xt])
        g = engine.pv_run(2048, 512, r)[0]
        o = oracle.pv_run(xt, 2048, 512, r)
        excluded = check(g, o)
        assert excluded == 0 and rms(g["y"], o["y"]) < 1e-6, track
        # the far end of the track is as exact as its start (no drift over 28 125 frames)
        assert rms(g["y"][-480000:], o["y"][-480000:]) < 1e-6


def test_sinusoid_in_shifted_sinusoid_out(engine):
    """A check that shares no code with either restatement of the spec: a steady sinusoid in gives a
    sinusoid at rate * f out (frequency read off the output by a zero-padded FFT in numpy, double), with
    steady amplitude, and the detected f0 / peak bin of the analysis matches the input tone."""
    fs, N, H = 48000.0, 2048, 512
    n = 4 * 48000
    t = np.arange(n) / fs
    for f_in, semis in ((440.0, 3.0), (997.0, -4.0), (233.3, 7.0)):
        x = (0.4 * np.sin(2 * np.pi * f_in * t)).astype(np.float32)
        r = float(ratio(semis))
        engine.upload_tracks([x])
        g = engine.pv_run(N, H, ratio(semis))[0]
        # analysis side: detected pitch
        mid = slice(8, g["f0"].size - 8)
        assert np.all(g["peak"][mid] == int(round(f_in * N / fs)))
        assert np.abs(g["f0"][mid] - f_in).max() < 1e-3
        # synthesis side: output frequency and amplitude steadiness on the interior
        y = g["y"][48000:-48000].astype(np.float64)
        w = np.hanning(y.size)
        Y = np.abs(np.fft.rfft(y * w, 1 << 20))
        k = int(np.argmax(Y))
        a, b, c = np.log(Y[k - 1]), np.log(Y[k]), np.log(Y[k + 1])
        f_out = (k + 0.5 * (a - c) / (a - 2 * b + c)) * fs / (1 << 20)
        assert abs(f_out - r * f_in) < 1e-3, (f_in, semis, f_out)   # the double oracle lands within 5e-6 Hz
        env = np.sqrt(np.convolve(y * y, np.ones(4800) / 4800, "valid") * 2.0)
        # (bin remapping spreads the main lobe: the level drops by a ratio-dependent factor but must be steady)
        assert env.min() > 0.35 * 0.4 and env.max() < 1.0 * 0.4 and env.max() / env.min() < 1.02


@pytest.mark.parametrize("N", [1024, 2048])
def test_pitch_up_kernel_is_bitwise_the_general_kernel(engine, monkeypatch, N):
    """K_A2 (csrc/pv_analyze2.cu: constant ratio >= 1; tail FFT stage in the bin threads, bin shift as a
    scatter) against the general analysis kernel (MLX_PV_NO_KA2=1): identical output samples, peak bins,
    f0 and phase totals -- ratios 1 (no empty bins), +3 st, +12 st (every other output bin empty), +24 st
    (three empty bins after each), ragged batch, short chunks, a wave boundary inside a chunk."""
    import torch
    H = N // 4
    xs = [S.vibrato_tone(4.0, seed=71), S.vibrato_tone(1.3, seed=72), np.zeros(700, np.float32),
          S.two_tone(2.0)]
    engine.use_torch_stream()
    for semis in (0.0, 3.0, 0.37, 12.0, 24.0):
        r = ratio(semis)
        for env in ({}, {"MLX_PV_CHUNK": "24"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            res = []
            for no_ka2 in ("0", "1"):
                if no_ka2 == "1":
                    monkeypatch.setenv("MLX_PV_NO_KA2", "1")
                else:
                    monkeypatch.delenv("MLX_PV_NO_KA2", raising=False)
                    monkeypatch.setenv("MLX_PV_KA2", "1")
                engine.upload_tracks(xs)
                out = engine.pv_run(N, H, r, wave_mib=(1 if env else -1))
                tots = [torch.zeros(N // 2 + 1, dtype=torch.int32, device="cuda") for _ in xs]
                engine.pv_phase_totals_dev(N, H, r, tots, wave_mib=-1)
                torch.cuda.synchronize()
                res.append((out, [t.cpu().numpy() for t in tots]))
            monkeypatch.delenv("MLX_PV_NO_KA2", raising=False)
            monkeypatch.delenv("MLX_PV_KA2", raising=False)
            for k in env:
                monkeypatch.delenv(k, raising=False)
            (a, ta), (b, tb) = res
            for u, v in zip(a, b):
                assert np.array_equal(u["y"], v["y"]), (N, semis, env)
                assert np.array_equal(u["peak"], v["peak"]) and np.array_equal(u["f0"], v["f0"]), (N, semis, env)
            for u, v in zip(ta, tb):
                assert np.array_equal(u, v), (N, semis, env)


def test_per_frame_rate_array(engine, oracle):
    import torch
    x = S.vibrato_tone(2.0, seed=41)
    F = (x.size + 511) // 512
    rates = (2.0 ** (np.linspace(-3, 5, F) / 12.0)).astype(np.float32)
    engine.upload_tracks([x])
    engine.use_torch_stream()
    y = torch.zeros(x.size, dtype=torch.float32, device="cuda")
    engine.pv_run_dev(2048, 512, 1.0, [y], rate_per_frame=[torch.from_numpy(rates).cuda()])
    torch.cuda.synchronize()
    o = oracle.pv_run(x, 2048, 512, 1.0, rate_per_frame=rates)
    assert rms(y.cpu().numpy(), o["y"]) <= 1e-4


def test_host_pipeline_equals_resident_run(engine):
    xs = [S.vibrato_tone(3.0, seed=51), S.vibrato_tone(2.2, seed=52), S.vibrato_tone(0.7, seed=53)]
    r = ratio(3.0)
    engine.upload_tracks(xs)
    a = engine.pv_run(2048, 512, r)
    ys = [np.zeros_like(x) for x in xs]
    pk = [np.zeros((x.size + 511) // 512, np.int32) for x in xs]
    f0 = [np.zeros((x.size + 511) // 512, np.float32) for x in xs]
    engine.pv_process_host(xs, 2048, 512, r, ys, pk, f0)
    for u, y, p, f in zip(a, ys, pk, f0):
        assert np.array_equal(u["y"], y) and np.array_equal(u["peak"], p) and np.array_equal(u["f0"], f)


def test_int16_wire_formats_are_exact(engine):
    """mlx_pv_process_host_fmt: int16 PCM in means x = s / 32768 (the same output, bit for bit, as handing in
    those floats); int16 out is the reference's export conversion int16(x * 32767.) by truncation
    (app.cpp:1209-1212) of the float output, bit for bit -- computed on the device, fused into K_S."""
    xs = [S.vibrato_tone(3.0, seed=81), S.vibrato_tone(1.1, seed=82), S.two_tone(0.4), S.vibrato_tone(2.0, seed=83),
          S.vibrato_tone(0.9, seed=84)]                               # five tracks: a group of four and a ragged one
    r = ratio(3.0)
    q = [np.round(x * 32767.0).astype(np.int16) for x in xs]
    xf = [(s.astype(np.float32) / np.float32(32768.0)) for s in q]    # exact in float32
    nf = [(x.size + 511) // 512 for x in xs]

    def run(ins, out_dtype):
        ys = [np.zeros(x.size, out_dtype) for x in xs]
        pk = [np.zeros(n, np.int32) for n in nf]
        f0 = [np.zeros(n, np.float32) for n in nf]
        engine.pv_process_host(ins, 2048, 512, r, ys, pk, f0)
        return ys, pk, f0

    y_ff, pk_ff, f0_ff = run(xf, np.float32)
    y_if, pk_if, f0_if = run(q, np.float32)
    y_fi, _, _ = run(xf, np.int16)
    y_ii, pk_ii, _ = run(q, np.int16)
    engine.upload_tracks(xf)
    ref = engine.pv_run(2048, 512, r)
    for t in range(len(xs)):
        assert np.array_equal(y_ff[t], ref[t]["y"]) and np.array_equal(pk_ff[t], ref[t]["peak"])
        assert np.array_equal(y_if[t], y_ff[t]) and np.array_equal(pk_if[t], pk_ff[t]) and np.array_equal(f0_if[t], f0_ff[t])
        want = np.trunc(y_ff[t].astype(np.float64) * 32767.0).astype(np.int16)
        assert np.array_equal(y_fi[t], want) and np.array_equal(y_ii[t], want) and np.array_equal(pk_ii[t], pk_ff[t])


@pytest.mark.parametrize("N,semis", [(2048, 3.0), (2048, -4.0), (4096, 3.0), (1024, 7.0)])
def test_every_bin_of_every_frame_against_the_oracle(engine, oracle, N, semis):
    """Not only the audio: the staged analysis itself (mlx_pv_stage_export_dev) against the oracle's per-frame
    debug output -- shifted magnitudes (PV-spec A.5) and accumulated synthesis phases (A.6) of EVERY bin of
    EVERY frame.  A single cut decision on the other side of +-pi than the oracle's would move that bin's
    phase by frac(rate) turns (8e8 counts at +3 st) for the rest of the track; what is allowed is the
    telescoping error of the float phase (1e-7 rad = 70 counts of 2^-32 turn) plus one rounding per frame."""
    import torch
    H = N // 4
    x = S.vibrato_tone(20.0, seed=1234)
    r = ratio(semis)
    o = oracle.pv_run(x, N, H, r, want_debug=True, want_audio=False)
    acc = np.cumsum(o["inc"].astype(np.uint64), axis=0).astype(np.uint32)      # mod 2^32
    F, nb = acc.shape
    engine.use_torch_stream()
    engine.upload_tracks([x])
    tot = torch.zeros(nb, dtype=torch.int32, device="cuda")
    engine.pv_analyze_dev(N, H, r, [tot])
    smag = torch.empty((F, nb), dtype=torch.float32, device="cuda")
    phase = torch.empty((F, nb), dtype=torch.int32, device="cuda")
    engine.pv_stage_export_dev(0, 0, F, smag, phase)
    torch.cuda.synchronize()
    sm = smag.cpu().numpy()
    ph = phase.cpu().numpy().view(np.uint32)
    # magnitudes: float32 from the FP64 spectrum
    scale = np.abs(o["smag"]).max()
    assert np.abs(sm - o["smag"]).max() <= 2e-6 * scale
    assert np.allclose(sm, o["smag"], rtol=5e-6, atol=1e-7 * scale)
    # phases: signed distance on the circle, in counts of 2^-32 turn
    d = (ph - acc).astype(np.int32).astype(np.int64)
    worst = int(np.abs(d).max())
    assert worst < (1 << 15), f"max phase deviation {worst} counts at frame/bin {np.unravel_index(np.abs(d).argmax(), d.shape)}"
    # ... and it does not grow along the track (first vs last quarter of the frames)
    q = F // 4
    assert np.abs(d[-q:]).max() < 4 * max(64, np.abs(d[:q]).max()) or np.abs(d[-q:]).max() < (1 << 13)
    # the totals handed to the next shard are the last frame's phases
    assert np.array_equal(tot.cpu().numpy().view(np.uint32), ph[-1])


@pytest.mark.parametrize("N", [1024, 2048, 4096])
def test_parseval_on_the_analysis_magnitudes(engine, N):
    """Energy conservation of K_A's spectrum, checked with numpy alone (no code shared with the oracle or with
    tests/np_pv_reference.py): at rate 1 the bin shift is the identity, so the staged magnitudes are |X_k| of
    the Hann-windowed frame and  sum_n (w x)^2 = (|X_0|^2 + |X_{N/2}|^2 + 2 sum_{0<k<N/2} |X_k|^2) / N
    must hold for every frame (float32 magnitudes: 1e-5 relative)."""
    import torch
    H = N // 4
    x = S.vibrato_tone(6.0, seed=77)
    engine.use_torch_stream()
    engine.upload_tracks([x])
    F = (x.size + H - 1) // H
    nb = N // 2 + 1
    tot = torch.zeros(nb, dtype=torch.int32, device="cuda")
    engine.pv_analyze_dev(N, H, 1.0, [tot])
    smag = torch.empty((F, nb), dtype=torch.float32, device="cuda")
    phase = torch.empty((F, nb), dtype=torch.int32, device="cuda")
    engine.pv_stage_export_dev(0, 0, F, smag, phase)
    torch.cuda.synchronize()
    m2 = smag.cpu().numpy().astype(np.float64) ** 2
    spec_energy = (m2[:, 0] + m2[:, -1] + 2.0 * m2[:, 1:-1].sum(axis=1)) / N
    # frame f covers samples [(f+1)H - N, (f+1)H), zero outside the track (spec.cpp:47-54 convention)
    xp = np.concatenate([np.zeros(N, np.float64), x.astype(np.float64), np.zeros(N, np.float64)])
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N) / N)).astype(np.float32).astype(np.float64)
    idx = (np.arange(F)[:, None] + 1) * H + np.arange(N)[None, :]       # + N (front padding) - N (frame start)
    time_energy = ((xp[idx] * w[None, :]) ** 2).sum(axis=1)
    assert np.allclose(spec_energy, time_energy, rtol=1e-5, atol=1e-9)
