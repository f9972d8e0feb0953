"""Secondary measurements for profiles/ (not the driver's bench): Spec STFT path (BASELINE configs[0]
geometry at scale, reference geometry, N sweep) and the PV N sweep (configs[4]), with achieved
algorithmic GB/s against the measured HBM peak, plus the reference's own spec.cpp timed on the host.

    python tools/extra_bench.py > profiles/r1_extra.json
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import melonix_b200 as m  # noqa: E402
from bench import gen_tracks_gpu  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
FS = 48000


def main():
    dev = torch.device("cuda", 0)
    eng = m.Engine(0)
    eng.use_torch_stream()
    nt, seconds = 16, 300.0
    n = int(seconds * FS)
    x = gen_tracks_gpu(torch, dev, nt, n, 0)
    eng.upload_tracks_dev([x[i] for i in range(nt)])
    out = {"peak_gbs": PEAK, "tracks": nt, "seconds": seconds, "spec": [], "pv": []}

    # ---- Spec path: regular-hop jobs over every track, reference window (spec.cpp:44-66)
    for N, hop in [(512, 128), (1024, 256), (2048, 512), (4096, 1024), (8192, 2048), (32768, 375)]:
        F = (n + hop - 1) // hop
        buf = torch.empty((F, N // 2), dtype=torch.float32, device=dev)
        def run():
            for t in range(nt):
                eng.spec_frames_dev(t, N, hop, 0, F, buf)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        eng.profile_enable(True); eng.profile_read()
        reps = 3
        for _ in range(reps):
            run()
        ms, launches = eng.profile_read()["spec"]
        eng.profile_enable(False)
        frames = nt * F * reps
        fps = frames / (ms * 1e-3)
        algo = 4 * hop + 2 * N
        out["spec"].append(dict(fftN=N, hop=hop, frames_per_s=fps, algorithmic_bytes_per_frame=algo,
                                achieved_gbs=fps * algo / 1e9, frac_of_hbm_peak=fps * algo / 1e9 / PEAK,
                                kernel_ms_per_launch=ms / launches, frames_per_launch=F))
        del buf
    # ---- PV sweep (BASELINE configs[4]): N at 4:1 hop, +3 semitones
    r = m.semitone_ratio(3.0)
    y = torch.empty_like(x)
    for N in (512, 1024, 2048, 4096, 8192):
        hop = N // 4
        F = (n + hop - 1) // hop
        peak = torch.empty((nt, F), dtype=torch.int32, device=dev)
        f0 = torch.empty((nt, F), dtype=torch.float32, device=dev)
        outs = ([y[i] for i in range(nt)], [peak[i] for i in range(nt)], [f0[i] for i in range(nt)])
        for _ in range(2):
            eng.pv_run_dev(N, hop, r, *outs)
        torch.cuda.synchronize()
        eng.profile_enable(True); eng.profile_read()
        reps = 3
        for _ in range(reps):
            eng.pv_run_dev(N, hop, r, *outs)
        prof = eng.profile_read()
        eng.profile_enable(False)
        ms = sum(prof[k][0] for k in ("pv_analyze", "pv_scan", "pv_synth")) / reps
        fps = nt * F / (ms * 1e-3)
        algo = 8 * hop + 8
        out["pv"].append(dict(fftN=N, hop=hop, frames_per_s=fps, algorithmic_bytes_per_frame=algo,
                              achieved_gbs=fps * algo / 1e9, frac_of_hbm_peak=fps * algo / 1e9 / PEAK,
                              kernel_ms={k: prof[k][0] / reps for k in ("pv_analyze", "pv_scan", "pv_synth")}))
        del peak, f0
    # ---- the reference's own spec.cpp on the host (one worker thread, as the reference runs it)
    try:
        from oracle import oracle as O
        if O.have_ref():
            xs = x[0, : 48000 * 20].cpu().numpy()
            jobs = np.stack([np.arange(1280) * 375, (np.arange(1280) + 1) * 375], 1).astype(np.int32)
            t0 = time.perf_counter()
            O.ref_spec_run(xs, jobs)
            el = time.perf_counter() - t0
            out["reference_spec_cpp"] = dict(frames_per_s=1280 / el, jobs=1280, fftN=32768, threads=1,
                                             note="reference spec.cpp unmodified + shim FFT (not FFTW), one screen of 1280 columns")
            t0 = time.perf_counter()
            O.spec_batch(xs, 32768, jobs, nthreads=0)
            el = time.perf_counter() - t0
            out["oracle_spec_all_cores"] = dict(frames_per_s=1280 / el, threads=O.num_threads())
    except Exception as e:  # noqa: BLE001
        out["reference_spec_cpp"] = dict(error=str(e))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
