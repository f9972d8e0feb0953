"""Ad-hoc GPU diagnostics (development aid): runs the C-ABI paths at small sizes against the oracle
and prints error summaries.  Not part of the test suite."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import melonix_b200 as m  # noqa: E402
from oracle import oracle as O  # noqa: E402
import signals as S  # noqa: E402


def rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def main():
    eng = m.Engine(0)
    print(eng.device_info())
    x = S.vibrato_tone(4.0, seed=1234)
    eng.upload_tracks([x])
    for N in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        hop = N // 4
        jobs = S.regular_jobs(x.size, hop)[:300]
        extra = np.array([[-5000, -10], [x.size + 40000, x.size + 40100], [-3, 100], [x.size - 7, x.size + 300],
                          [1000, 1001], [5000, 9000]], np.int32)
        jobs = np.concatenate([jobs, extra])
        t0 = time.time()
        g = eng.spec_batch(0, N, jobs)
        t1 = time.time()
        o = O.spec_batch(x, N, jobs)
        print(f"spec N={N:6d} jobs={len(jobs)} rms={rms(g, o):.3e} max={np.abs(g - o).max():.3e} "
              f"peak={o.max():.4f} gpu_s={t1 - t0:.4f}")
    for N in (2048, 512, 1024, 4096, 8192):
        hop = N // 4
        for semis in (3.0, -4.0, 0.0):
            r = m.semitone_ratio(semis)
            t0 = time.time()
            g = eng.pv_run(N, hop, r)[0]
            t1 = time.time()
            o = O.pv_run(x, N, hop, r)
            ok = o["margin"] > 1e-5
            print(f"pv N={N:5d} st={semis:+.0f} rms={rms(g['y'], o['y']):.3e} max={np.abs(g['y'] - o['y']).max():.3e} "
                  f"peak_eq={np.array_equal(g['peak'][ok], o['peak'][ok])} ({(~ok).sum()} excl) "
                  f"f0_max={np.abs(g['f0'] - o['f0']).max():.3e} gpu_s={t1 - t0:.4f}")
    # drift: long stationary tones (systematic per-frame phase errors would grow linearly in time)
    n = 300 * 48000
    t = np.arange(n) / 48000.0
    xs = (0.3 * np.sin(2 * np.pi * 441.3 * t) + 0.15 * np.sin(2 * np.pi * 1237.7 * t + 0.3)
          + 1e-4 * np.random.default_rng(5).standard_normal(n)).astype(np.float32)
    eng.upload_tracks([xs])
    r = m.semitone_ratio(3.0)
    g = eng.pv_run(2048, 512, r)[0]
    o = O.pv_run(xs, 2048, 512, r)
    seg = 5 * 48000
    print(f"drift 300 s stationary: rms first 5 s {rms(g['y'][:seg], o['y'][:seg]):.3e}  "
          f"mid {rms(g['y'][n // 2:n // 2 + seg], o['y'][n // 2:n // 2 + seg]):.3e}  "
          f"last 5 s {rms(g['y'][-seg:], o['y'][-seg:]):.3e}  whole {rms(g['y'], o['y']):.3e} "
          f"peak_eq={np.array_equal(g['peak'], o['peak'])}")
    # multi-track ragged + wave tiling
    xs = [S.vibrato_tone(2.0, seed=1), S.vibrato_tone(1.37, seed=2, f_base=330.0), S.vibrato_tone(0.2, seed=3)]
    eng.upload_tracks(xs)
    r = m.semitone_ratio(3.0)
    a = eng.pv_run(2048, 512, r, wave_mib=-1)
    b = eng.pv_run(2048, 512, r, wave_mib=1)
    for i, xx in enumerate(xs):
        o = O.pv_run(xx, 2048, 512, r)
        print(f"ragged track {i}: rms={rms(a[i]['y'], o['y']):.3e} waves_bitwise={np.array_equal(a[i]['y'], b[i]['y'])} "
              f"peak_eq={np.array_equal(a[i]['peak'], o['peak'])}")
    eng.close()


if __name__ == "__main__":
    main()
