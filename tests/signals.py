"""Seeded synthetic inputs of SURVEY.md section 8(d) (float32, generated in float64 then cast)."""
import numpy as np

FS = 48000


def sine_sweep(seconds=10.0, f_lo=110.0, f_hi=7040.0, amp=0.5, fs=FS):
    """config 1: linear sine sweep (no RNG)."""
    n = int(round(seconds * fs))
    t = np.arange(n) / fs
    phase = 2 * np.pi * (f_lo * t + 0.5 * (f_hi - f_lo) / seconds * t * t)
    return (amp * np.sin(phase)).astype(np.float32)


def vibrato_tone(seconds=60.0, seed=1234, f_base=220.0, fs=FS, noise_db=-50.0):
    """config 2/3: 8-harmonic tone (amps 1/h, peak 0.5), f0(t) = f_base * 2^(sin(2 pi 0.5 t)/12),
    plus white noise at noise_db dBFS."""
    n = int(round(seconds * fs))
    t = np.arange(n) / fs
    f0 = f_base * 2.0 ** (np.sin(2 * np.pi * 0.5 * t) / 12.0)
    ph = 2 * np.pi * np.cumsum(f0) / fs
    x = np.zeros(n)
    for h in range(1, 9):
        x += np.sin(h * ph) / h
    x *= 0.5 / np.abs(x).max()
    rng = np.random.default_rng(seed)
    x += rng.standard_normal(n) * 10.0 ** (noise_db / 20.0)
    return x.astype(np.float32)


def two_tone(seconds=20.0, fs=FS):
    """grain-path probe signal of SURVEY.md section 4 (KAT-4): 220 + 440 Hz."""
    n = int(round(seconds * fs))
    t = np.arange(n) / fs
    return (0.3 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)


def regular_jobs(n, hop):
    """reference job convention: frame f = (f*hop, (f+1)*hop) (spec-cache.cpp:63-65)."""
    F = (n + hop - 1) // hop
    f = np.arange(F, dtype=np.int64)
    return np.stack([f * hop, (f + 1) * hop], axis=1).astype(np.int32)
