"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference for the
reference-geometry vectors, which come from the reference's own spec.cpp compiled unmodified):

    python tests/golden/make_golden.py

Inputs are rebuilt from formulas / seeds by the tests (tests/signals.py); the files hold only the
job lists and the expected outputs."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import signals as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    O.build()
    # --- Spec, reference geometry (N = 32768, hop = 375: 10 s / 1280 px at 48 kHz), from the
    #     reference's own Spec class (oracle/_ref).  KAT-1 of SURVEY.md section 4 included.
    assert O.have_ref(), "needs oracle/_ref (built from /root/reference/spec.cpp)"
    n = 200000
    kat = (0.5 * np.sin(2 * np.pi * 300 * np.arange(n) / 32768)).astype(np.float32)
    jobs = np.array([[100000, 100375], [0, 375], [-400, -25], [199900, 200275], [n + 32768, n + 33143],
                     [5000, 5375]], np.int32)
    ref = O.ref_spec_run(kat, jobs)
    np.savez_compressed(OUT / "spec_ref_geometry.npz", jobs=jobs, out=ref, n=n)
    print("spec_ref_geometry: peak", ref[0].max(), "argmax", ref[0].argmax())

    # --- Spec, config 1 geometry (1024 / 256) on the sine sweep, oracle restatement
    x = S.sine_sweep(2.0)
    jobs = S.regular_jobs(x.size, 256)
    sel = np.r_[0:8, 180:188, jobs.shape[0] - 8:jobs.shape[0]]
    out = O.spec_batch(x, 1024, jobs[sel], nthreads=1)
    np.savez_compressed(OUT / "spec_cfg1.npz", jobs=jobs[sel], out=out, seconds=2.0)

    # --- PV (NOT IN REFERENCE): oracle self-consistency vectors
    x = S.vibrato_tone(0.5, seed=1234)
    r = np.float32(2.0) ** (np.float32(3.0) / np.float32(12.0))
    o = O.pv_run(x, 2048, 512, r)
    np.savez_compressed(OUT / "pv_2048_p3.npz", y=o["y"], peak=o["peak"], f0=o["f0"], margin=o["margin"],
                        seconds=0.5, seed=1234, rate=np.float32(r))
    r2 = np.float32(2.0) ** (np.float32(-5.0) / np.float32(12.0))
    o = O.pv_run(x, 1024, 256, r2)
    np.savez_compressed(OUT / "pv_1024_m5.npz", y=o["y"], peak=o["peak"], f0=o["f0"], margin=o["margin"],
                        seconds=0.5, seed=1234, rate=np.float32(r2))

    # --- grains: segmentation + export at +3 semitones (markers near both ends)
    x = S.two_tone(3.0)
    gs, gl = O.grain_segment(x)
    mk = [(10, 0.0, 0.0, 3.0), (x.size - 10, 0.0, 0.0, 3.0)]
    e = O.grain_export(x, 48000, mk, gs, gl)
    np.savez_compressed(OUT / "grain_p3.npz", g_start=gs, g_len=gl, pcm=e["pcm"], pcm16=e["pcm16"],
                        s_gstart=e["schedule"]["gstart"], s_glen=e["schedule"]["glen"],
                        s_rate=e["schedule"]["rate"], s_out_off=e["schedule"]["out_off"],
                        s_next=e["schedule"]["next"], seconds=3.0)

    # --- waveform min/max pyramid + range queries (App::calcPicks / getMinMaxFromRange, app.cpp:347-426)
    x = S.vibrato_tone(5003 / 48000.0 + 0.01, seed=31)[:5003]
    pairs, off = O.picks_build(x)
    rng = np.random.default_rng(31)
    start = rng.integers(0, 5000, 300)
    ranges = np.stack([start, np.minimum(start + (2.0 ** rng.uniform(0, 12, 300)).astype(np.int64), 5002)], 1)
    ranges = np.concatenate([ranges, [[5, 5], [7, 3], [5003, 5003], [-3, 10], [10, -3], [0, 5003], [0, 5002], [3, 4]]])
    ranges = ranges.astype(np.int32)
    np.savez_compressed(OUT / "picks_5003.npz", n=5003, seed=31, pairs=pairs, level_off=off, ranges=ranges,
                        minmax=O.minmax_ranges(x, pairs, off, ranges))

    # --- colour ramp
    v = np.linspace(0, 3.0, 4001).astype(np.float32)
    np.savez_compressed(OUT / "colormap.npz", v=v, k=np.float32(100.0), rgb=O.colormap(v, 100.0))
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
