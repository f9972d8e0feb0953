#!/usr/bin/env python
"""bench.py -- phase-vocoder frames/sec (2048-FFT, hop 512, 48 kHz mono) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU implementation of the same path, host cores

A "step" is one pass of the whole hot path (K_A analysis -> scan -> K_S synthesis) over
one batch of synthetic tracks.  Workload = BASELINE.json configs[2]: 64 mono tracks x 5 min per GPU,
2048-FFT / 512-hop, +3 semitones (tracks shard across GPUs with no data-path collective: weak
scaling, every rank processes its own 64 tracks).  Inputs (3.7 GB) and outputs (3.7 GB) per GPU are
far larger than the 126 MB L2, so no flush is needed between timed steps.

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM (CUDA
events, max over ranks); `e2e` = the same through the host-buffer C-ABI call with pinned host input
and output, copies inside the timed region; `roofline` = algorithmic bytes (8*hop+8 per frame,
SURVEY.md section 8d) over the summed device time of the path's kernels, against the measured HBM
peak; `cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FS = 48000
FFT_N = 2048
HOP = 512
SEMITONES = 3.0
METRIC = "phase-vocoder frames/sec (2048-FFT, hop 512, 48 kHz mono)"
ALGO_BYTES_PER_FRAME = 8 * HOP + 8  # 4H in + 4H out + peakBin i32 + f0 f32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=64, help="tracks per GPU")
    ap.add_argument("--seconds", type=float, default=300.0, help="seconds per track")
    ap.add_argument("--wave-mib", type=int, default=0, help="intermediate-stage budget in MiB (0 = library default: one wave)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (configs[3] time-sharded, ...)")
    ap.add_argument("--cfg3-seconds", type=float, default=7200.0, help="length of the configs[3] file")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU baseline")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arm (0 = all this process may use)")
    return ap.parse_args()


def semitone_ratio(st):
    return np.float32(np.power(np.float32(2.0), np.float32(st) / np.float32(12.0), dtype=np.float32))


def vibrato_track_numpy(seconds, seed, f_base=220.0):
    """cfg-2/3 signal (tests/signals.py:vibrato_tone), used for the CPU sample."""
    n = int(round(seconds * FS))
    t = np.arange(n) / FS
    f0 = f_base * 2.0 ** (np.sin(2 * np.pi * 0.5 * t) / 12.0)
    ph = 2 * np.pi * np.cumsum(f0) / FS
    x = np.zeros(n)
    for h in range(1, 9):
        x += np.sin(h * ph) / h
    x *= 0.5 / np.abs(x).max()
    x += np.random.default_rng(seed).standard_normal(n) * 10.0 ** (-50.0 / 20.0)
    return x.astype(np.float32)


# ------------------------------------------------------------------------------------------------
def host_cores():
    """Host threads this process may use.  Taken from the scheduler affinity, NOT from OpenMP:
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which would silently turn the
    "all host cores" CPU arm into a single-core run (round-1 SCALE artefact)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_one_thread(sec=10.0):
    """The same port on ONE thread (the reference's actual concurrency: one worker, spec.cpp:11)."""
    from oracle import oracle as O
    x = vibrato_track_numpy(sec, 1234)[None, :]
    r = semitone_ratio(SEMITONES)
    F = O.pv_num_frames(x.shape[1], HOP)
    t0 = time.perf_counter()
    O.pv_run_batch(x, FFT_N, HOP, r, FS, 1)
    return F / (time.perf_counter() - t0)


def cpu_baseline(target_seconds, nthreads=0):
    """Times the oracle port of the path (oracle/pv_ref.c, OpenMP over tracks) on the host cores.
    Sample: `threads` tracks x 20 s of the cfg-3 signal, repeated until ~target_seconds elapse."""
    from oracle import oracle as O
    O.build()
    threads = nthreads or host_cores()
    sec = 20.0
    base = vibrato_track_numpy(sec, 1234)
    x = np.ascontiguousarray(np.tile(base, (threads, 1)))
    r = semitone_ratio(SEMITONES)
    F = O.pv_num_frames(x.shape[1], HOP)
    O.pv_run_batch(x[: max(1, threads // 8)], FFT_N, HOP, r, FS, threads)  # warm-up (page in, spin up OpenMP)
    reps, t0 = 0, time.perf_counter()
    while True:
        O.pv_run_batch(x, FFT_N, HOP, r, FS, threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= target_seconds or reps >= 50:
            break
    return dict(value=threads * F * reps / el, unit="frames/s", cores=threads, kind="port",
                one_thread_value=cpu_one_thread(),
                sample=f"{threads} tracks x {sec:.0f} s (cfg-3 signal, {F} frames each), {reps} passes, "
                       f"{el:.1f} s wall; double-precision oracle/pv_ref.c, OpenMP over tracks "
                       f"(FFT engine: in-repo radix-4 double FFT, not FFTW)")


def reference_arm(args):
    """--impl reference: the reference has no phase vocoder and its spec.cpp
    is a different transform, so the CPU implementation of this metric's path is the oracle port.
    Each step is one pass over a bounded sample (threads tracks x 20 s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = args.cpu_threads or host_cores()   # explicit: never inherit OMP_NUM_THREADS=1 from torchrun
    sec = 20.0
    base = vibrato_track_numpy(sec, 1234)
    x = np.ascontiguousarray(np.tile(base, (threads, 1)))
    r = semitone_ratio(SEMITONES)
    F = O.pv_num_frames(x.shape[1], HOP)
    for _ in range(min(args.warmup, 2)):
        O.pv_run_batch(x, FFT_N, HOP, r, FS, threads)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.pv_run_batch(x, FFT_N, HOP, r, FS, threads)
    el = time.perf_counter() - t0
    v = threads * F * steps / el
    sample = (f"{threads} tracks x {sec:.0f} s per step (cfg-3 signal), {steps} timed steps; oracle/pv_ref.c "
              f"(double precision, OpenMP over tracks; in-repo FFT, not FFTW)")
    line = dict(metric=METRIC, value=v, unit="frames/s", n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 2),
                ms_per_step=1e3 * el / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="cfg-3 signal, 2048-FFT/512-hop, +3 st, bounded CPU sample", fft=FFT_N, hop=HOP,
                            semitones=SEMITONES, tracks=threads, seconds=sec),
                cpu_baseline=dict(value=v, unit="frames/s", cores=threads, kind="port", sample=sample,
                                  one_thread_value=cpu_one_thread(), omp_num_threads_env=os.environ.get("OMP_NUM_THREADS")),
                e2e=dict(value=v, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=0)
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples), power_w_max=max(self.power) if self.power else None)


def gen_tracks_gpu(torch, dev, ntracks, n, rank):
    """cfg-3 synthetic batch generated on the device (float64 phase, cast to float32):
    8-harmonic vibrato tone, per-track base 110*2^(t/64*2) Hz, -50 dBFS white noise, seed 1234+track."""
    x = torch.empty((ntracks, n), dtype=torch.float32, device=dev)
    t = torch.arange(n, device=dev, dtype=torch.float64) / FS
    vib = torch.sin(2 * np.pi * 0.5 * t) / 12.0
    for tr in range(ntracks):
        gtrack = rank * ntracks + tr
        f_base = 110.0 * 2.0 ** ((gtrack % 64) / 64.0 * 2.0)
        f0 = f_base * torch.pow(torch.tensor(2.0, device=dev, dtype=torch.float64), vib)
        ph = 2 * np.pi * torch.cumsum(f0, 0) / FS
        s = torch.zeros(n, device=dev, dtype=torch.float64)
        for h in range(1, 9):
            s += torch.sin(h * ph) / h
        s *= 0.5 / s.abs().max()
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + gtrack)
        s += torch.randn(n, device=dev, dtype=torch.float64, generator=gen) * 10.0 ** (-50.0 / 20.0)
        x[tr] = s.to(torch.float32)
        del f0, ph, s
    return x


def gen_long_file_gpu(torch, dev, lo, hi, seed, block=1 << 24):
    """Samples [lo, hi) of the configs[3] synthetic channel, a closed-form function of the GLOBAL sample index
    (so that every rank can produce its own time range and rank 0 the whole file, bit-identically):
    8-harmonic tone, f0(t) = 220 (1 + 0.0578 sin(2 pi 0.5 t)) Hz (+-1 semitone), phase integrated in closed
    form in float64, amplitude 0.25, plus a -50 dBFS hash-noise of the sample index."""
    out = torch.empty(hi - lo, dtype=torch.float32, device=dev)
    for b0 in range(lo, hi, block):
        b1 = min(hi, b0 + block)
        i = torch.arange(b0, b1, device=dev, dtype=torch.int64)
        t = i.to(torch.float64) / FS
        ph = 2 * np.pi * 220.0 * (t - (0.0578 / (2 * np.pi * 0.5)) * torch.cos(2 * np.pi * 0.5 * t))
        sig = torch.zeros_like(t)
        for h in range(1, 9):
            sig += torch.sin(h * ph) / h
        # counter-based noise: two rounds of a 64-bit mix of (index, seed) -> uniform -> centred, unit variance
        z = i * 0x9E3779B97F4A7C1 + (seed * 0x632BE59BD9B4E019 & 0x7FFFFFFFFFFFFFFF)
        z = (z ^ (z >> 30)) * 0x3F58476D1CE4E5B
        z = (z ^ (z >> 27)) * 0x14D049BB133111EB
        u = ((z ^ (z >> 31)) & 0xFFFFFF).to(torch.float64) / float(1 << 24)
        sig = 0.25 / 1.9 * sig + (u - 0.5) * (12.0 ** 0.5) * 10.0 ** (-50.0 / 20.0)
        out[b0 - lo:b1 - lo] = sig.to(torch.float32)
    return out


def cfg3_time_sharded(torch, dist, eng, dev, rank, world, seconds=7200.0, reps=3):
    """BASELINE configs[3]: one 2 h "stereo" file (two planar mono channels), 4096-FFT / 1024-hop, +3 st,
    sharded by contiguous time range across the ranks through mlx_pv_run_sharded_dev (NCCL seam send/recv of
    the overlap-region samples + all-gather of the uint32 phase totals behind the C ABI, ONE analysis pass).
    Rank 0 also runs the whole file unsharded on its own GPU: the 1-GPU time the speed-up refers to and the
    bits the gathered sharded output must equal."""
    from melonix_b200 import dist as D
    N, H = 4096, 1024
    n = int(round(seconds * FS))
    F = (n + H - 1) // H
    rate = semitone_ratio(SEMITONES)
    sh = D.shard_frames(n, N, H, world, rank)
    owns = [gen_long_file_gpu(torch, dev, sh.own_lo, sh.own_hi, 1234 + c) for c in range(2)]
    nf = sh.frame_end - sh.frame_begin
    outs = ([torch.empty_like(o) for o in owns], [torch.empty(nf, dtype=torch.int32, device=dev) for _ in owns],
            [torch.empty(nf, dtype=torch.float32, device=dev) for _ in owns])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    ms_1gpu = None
    ref = None
    if rank == 0:   # the unsharded run (whole file on one GPU)
        full = [gen_long_file_gpu(torch, dev, 0, n, 1234 + c) for c in range(2)]
        ref = [torch.empty_like(f) for f in full]
        pk = [torch.empty(F, dtype=torch.int32, device=dev) for _ in full]
        f0 = [torch.empty(F, dtype=torch.float32, device=dev) for _ in full]
        eng.use_torch_stream()
        eng.upload_tracks_dev(full)
        del full
        eng.pv_run_dev(N, H, rate, ref, pk, f0, sample_rate=FS, wave_mib=-1)   # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.pv_run_dev(N, H, rate, ref, pk, f0, sample_rate=FS, wave_mib=-1)
        e1.record()
        torch.cuda.synchronize()
        ms_1gpu = e0.elapsed_time(e1) / reps
        del pk, f0
    if world > 1:
        sync_all()
        D.run_time_sharded(eng, owns, n, N, H, rate, sample_rate=FS, outs=outs)   # warm-up: communicator, tables
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            D.run_time_sharded(eng, owns, n, N, H, rate, sample_rate=FS, outs=outs)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # bit-for-bit check on rank 0: every rank ships its owned output
        equal = True
        for c in range(2):
            for r in range(world):
                s = D.shard_frames(n, N, H, world, r)
                if rank == 0:
                    got = outs[0][c] if r == 0 else torch.empty(s.own_hi - s.own_lo, dtype=torch.float32, device=dev)
                    if r != 0:
                        dist.recv(got, r)
                    equal = equal and bool(torch.equal(got, ref[c][s.own_lo:s.own_hi]))
                    del got
                elif rank == r:
                    dist.send(outs[0][c], 0)
        comm = getattr(eng, "_comm", None)
        info = dict(ms=ms, frames=2 * F, frames_per_s=2 * F / (ms * 1e-3), bitwise_equal=equal,
                    nccl_version=comm.nccl_version if comm is not None else None)
    else:
        info = dict(ms=ms_1gpu, frames=2 * F, frames_per_s=2 * F / (ms_1gpu * 1e-3), bitwise_equal=None)
    if rank == 0:
        info.update(ms_1gpu=ms_1gpu, frames_per_s_1gpu=2 * F / (ms_1gpu * 1e-3),
                    speedup_vs_1gpu=ms_1gpu / info["ms"], n_gpus=world,
                    workload=f"configs[3]: {seconds / 3600:.1f} h x 2 planar channels, 4096-FFT/1024-hop, +3 st, "
                             f"time-range sharded; timed region = seam exchange + analysis + phase all-gather + synthesis",
                    seam_floats_per_track=[N, 3 * H], phase_words_per_track=N // 2 + 1,
                    api="mlx_pv_run_sharded_dev (NCCL inside libmelonix_b200.so)" if world > 1 else "mlx_pv_run_dev")
    del owns, outs, ref
    torch.cuda.empty_cache()
    return info if rank == 0 else None


def extras_single_gpu(torch, eng, dev, peak_gbs, nt=16, seconds=300.0):
    """The other BASELINE configs on one GPU (secondary numbers; the headline stays configs[2]):
    configs[0] geometry (Spec STFT 1024/256) and the FFT-size sweeps of configs[4] for the Spec path
    (4H + 2N algorithmic bytes per frame, one batched launch over all tracks) and the PV path (8H + 8),
    configs[1] (one 60 s track, full pitch shift)."""
    import melonix_b200 as m
    out = dict(tracks=nt, seconds_per_track=seconds, spec=[], pv=[])
    n = int(seconds * FS)
    x = gen_tracks_gpu(torch, dev, nt, n, 0)
    eng.use_torch_stream()
    eng.upload_tracks_dev([x[i] for i in range(nt)])
    for N, hop in [(512, 128), (1024, 256), (2048, 512), (4096, 1024), (8192, 2048)]:
        Fr = (n + hop - 1) // hop
        buf = torch.empty((nt, Fr, N // 2), dtype=torch.float32, device=dev)
        outs = [buf[i] for i in range(nt)]
        for _ in range(2):
            eng.spec_frames_all_dev(N, hop, outs)
        torch.cuda.synchronize()
        eng.profile_enable(True)
        eng.profile_read()
        reps = 5
        for _ in range(reps):
            eng.spec_frames_all_dev(N, hop, outs)
        ms, launches = eng.profile_read()["spec"]
        eng.profile_enable(False)
        fps = nt * Fr * reps / (ms * 1e-3)
        algo = 4 * hop + 2 * N
        out["spec"].append(dict(fftN=N, hop=hop, frames_per_s=fps, algorithmic_bytes_per_frame=algo,
                                achieved_gbs=fps * algo / 1e9, frac_of_hbm_peak=fps * algo / 1e9 / peak_gbs,
                                kernel_ms_per_launch=ms / max(launches, 1), frames_per_launch=nt * Fr))
        del buf, outs
    r = semitone_ratio(SEMITONES)
    y = torch.empty_like(x)
    for N in (512, 1024, 2048, 4096, 8192):
        hop = N // 4
        Fr = (n + hop - 1) // hop
        pk = torch.empty((nt, Fr), dtype=torch.int32, device=dev)
        f0 = torch.empty((nt, Fr), dtype=torch.float32, device=dev)
        o = ([y[i] for i in range(nt)], [pk[i] for i in range(nt)], [f0[i] for i in range(nt)])
        for _ in range(2):
            eng.pv_run_dev(N, hop, r, *o, sample_rate=FS)
        torch.cuda.synchronize()
        eng.profile_enable(True)
        eng.profile_read()
        reps = 3
        for _ in range(reps):
            eng.pv_run_dev(N, hop, r, *o, sample_rate=FS)
        prof = eng.profile_read()
        eng.profile_enable(False)
        ms = sum(prof[k][0] for k in ("pv_analyze", "pv_scan", "pv_synth")) / reps
        fps = nt * Fr / (ms * 1e-3)
        algo = 8 * hop + 8
        out["pv"].append(dict(fftN=N, hop=hop, frames_per_s=fps, algorithmic_bytes_per_frame=algo,
                              achieved_gbs=fps * algo / 1e9, frac_of_hbm_peak=fps * algo / 1e9 / peak_gbs,
                              kernel_ms={k: prof[k][0] / reps for k in ("pv_analyze", "pv_scan", "pv_synth")}))
        del pk, f0
    # configs[1]: ONE 60 s track, 2048/512 (23 MB of compulsory traffic: launch- and latency-bound)
    n1 = 60 * FS
    eng.upload_tracks_dev([x[0, :n1].contiguous()])
    F1 = (n1 + HOP - 1) // HOP
    y1 = torch.empty(n1, dtype=torch.float32, device=dev)
    p1 = torch.empty(F1, dtype=torch.int32, device=dev)
    g1 = torch.empty(F1, dtype=torch.float32, device=dev)
    for _ in range(3):
        eng.pv_run_dev(FFT_N, HOP, r, [y1], [p1], [g1], sample_rate=FS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.pv_run_dev(FFT_N, HOP, r, [y1], [p1], [g1], sample_rate=FS)
    e1.record()
    torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / 20
    out["cfg1_single_60s_track"] = dict(frames=F1, ms=ms1, frames_per_s=F1 / (ms1 * 1e-3),
                                        note="one 60 s track per call: 23 MB of traffic, three launches, latency-bound")
    # configs[0]: the 10 s sweep, Spec STFT 1024/256 (1875 frames: a 5.8 MB job, reported for completeness)
    n0 = 10 * FS
    eng.upload_tracks_dev([x[0, :n0].contiguous()])
    F0 = (n0 + 255) // 256
    b0 = torch.empty((F0, 512), dtype=torch.float32, device=dev)
    for _ in range(3):
        eng.spec_frames_dev(0, 1024, 256, 0, F0, b0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        eng.spec_frames_dev(0, 1024, 256, 0, F0, b0)
    e1.record()
    torch.cuda.synchronize()
    ms0 = e0.elapsed_time(e1) / 50
    out["cfg0_spec_10s"] = dict(frames=F0, ms=ms0, frames_per_s=F0 / (ms0 * 1e-3),
                                note="10 s, 1024-FFT/256-hop Spec STFT: one 1875-frame launch, launch-latency-bound")
    del x, y
    torch.cuda.empty_cache()
    return out


def load_traffic():
    """dram bytes per launch from the committed ncu capture (profiles/), if any."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import melonix_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner (any NCCL_DEBUG level >= VERSION) to stdout: keep stdout to the one
        # JSON line by sending NCCL's log to stderr instead of changing its level
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and os.environ.get("MLX_BENCH_NUMA", "1") != "0":
        from melonix_b200.dist import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local)  # pinned e2e buffers local to the GPU's PCIe root

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nt = args.tracks
    n = int(round(args.seconds * FS))
    F = (n + HOP - 1) // HOP
    frames_per_rank = nt * F
    rate = semitone_ratio(SEMITONES)

    eng = m.Engine(local)
    x = gen_tracks_gpu(torch, dev, nt, n, rank)
    y = torch.empty_like(x)
    peak = torch.empty((nt, F), dtype=torch.int32, device=dev)
    f0 = torch.empty((nt, F), dtype=torch.float32, device=dev)
    eng.use_torch_stream()
    eng.upload_tracks_dev([x[i] for i in range(nt)])
    outs = ([y[i] for i in range(nt)], [peak[i] for i in range(nt)], [f0[i] for i in range(nt)])

    def step():
        eng.pv_run_dev(FFT_N, HOP, rate, outs[0], outs[1], outs[2], sample_rate=FS, wave_mib=args.wave_mib)

    # ---- device-resident timing: W warm-up steps, then exactly K steps between CUDA events
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.profile_enable(True)
    eng.profile_read(reset=True)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - l0
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)
    ms_per_step = ms_total / args.steps
    value = world * frames_per_rank / (ms_per_step * 1e-3)

    # ---- roofline of the path's kernels (this rank; per GPU)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak_gbs, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kern_ms = {k: v[0] / args.steps for k, v in prof.items() if v[1] > 0}
    path_ms = sum(kern_ms.values())
    algo_bytes = ALGO_BYTES_PER_FRAME * frames_per_rank
    achieved = algo_bytes / (path_ms * 1e-3) / 1e9 if path_ms > 0 else None
    dom = max(kern_ms, key=kern_ms.get) if kern_ms else None
    traffic = load_traffic()
    roofline = dict(bound="hbm", achieved=achieved, peak=peak_gbs, unit="GB/s",
                    frac=(achieved / peak_gbs) if achieved else None,
                    traffic=(traffic or {}).get("path_bytes_per_step"),
                    # what actually binds the FFT kernels: the L1 / shared-memory data pipe (same ncu capture)
                    l1_data_pipe_pct_of_peak=(traffic or {}).get("l1_data_pipe_pct_of_peak"),
                    peak_source=peak_src,
                    kernel="pv_analyze + pv_scan + pv_synth (the path is three launches per wave)",
                    algorithmic_bytes_per_frame=ALGO_BYTES_PER_FRAME, frames_per_step=frames_per_rank,
                    kernel_ms_per_step=kern_ms, dominant=dom,
                    kernel_share={k: v / path_ms for k, v in kern_ms.items()} if path_ms else None,
                    launches_per_step={k: v[1] / args.steps for k, v in prof.items() if v[1] > 0})

    # ---- end to end through the host-buffer C-ABI call (pinned host in/out, copies timed).  Headline `e2e`:
    #      int16 PCM on both sides of the wire (the reference's export sink takes int16, app.cpp:1209-1212;
    #      16-bit PCM in is x = s / 32768) -- half the bytes of the float32 form, which is reported next to it.
    e2e = e2e_f32 = None
    if not args.no_e2e:
        del y
        torch.cuda.empty_cache()

        def run_e2e(dtype, steps):
            hx = torch.empty((nt, n), dtype=dtype, pin_memory=True)
            if dtype == torch.int16:
                for i in range(nt):   # 16-bit PCM of the same synthetic tracks
                    hx[i].copy_((x[i] * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
            else:
                hx.copy_(x)
            hy = torch.empty((nt, n), dtype=dtype, pin_memory=True)
            hp = torch.empty((nt, F), dtype=torch.int32, pin_memory=True)
            hf = torch.empty((nt, F), dtype=torch.float32, pin_memory=True)
            ins = [hx[i] for i in range(nt)]
            ho = ([hy[i] for i in range(nt)], [hp[i] for i in range(nt)], [hf[i] for i in range(nt)])

            def estep():
                eng.pv_process_host(ins, FFT_N, HOP, rate, ho[0], ho[1], ho[2], sample_rate=FS, wave_mib=args.wave_mib)

            estep()  # warm-up (allocations)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                estep()
            barrier()
            el = max_over_ranks(time.perf_counter() - t0)
            # pure-copy floor of the same bytes on this box (H2D and D2H at once, no kernels): what the host <->
            # device path alone allows, measured in the same run because it differs from box to box
            d_in = torch.empty((nt, n), dtype=dtype, device=dev)
            d_out = torch.empty((nt, n), dtype=dtype, device=dev)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            floor = None
            for _ in range(2):
                torch.cuda.synchronize()
                barrier()
                t1 = time.perf_counter()
                with torch.cuda.stream(s1):
                    d_in.copy_(hx, non_blocking=True)
                with torch.cuda.stream(s2):
                    hy.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                barrier()
                floor = max_over_ranks(time.perf_counter() - t1)
            del d_in, d_out
            bps = hx.element_size()
            fmt = "int16 PCM" if dtype == torch.int16 else "float32"
            return dict(value=world * frames_per_rank * steps / el, unit="frames/s",
                        h2d_bytes_per_step=int(nt * n * bps), d2h_bytes_per_step=int(nt * n * bps + nt * F * 8),
                        ms_per_step=1e3 * el / steps, steps=steps, sample_format=f"{fmt} in, {fmt} out",
                        copy_floor_ms=1e3 * floor,
                        api="mlx_pv_process_host_fmt (C ABI; pinned host buffers; H2D / D2H on copy streams overlapped "
                            "with the kernels, tracks in groups of 2 per launch)")

        e2e_f32 = run_e2e(torch.float32, max(1, min(2, args.e2e_steps)))
        e2e = run_e2e(torch.int16, args.e2e_steps)
        del x
        torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras:
        eng.upload_tracks([np.zeros(16, np.float32)])   # release the batch's track buffer
        try:
            c3 = cfg3_time_sharded(torch, dist, eng, dev, rank, world, seconds=args.cfg3_seconds)
            if rank == 0:
                extras["cfg3_time_sharded"] = c3
        except Exception as e:  # noqa: BLE001  (the headline numbers above must survive a failing extra)
            if rank == 0:
                extras["cfg3_time_sharded"] = dict(error=f"{type(e).__name__}: {e}")
        if world == 1:
            try:
                extras["other_configs"] = extras_single_gpu(torch, eng, dev, peak_gbs)
            except Exception as e:  # noqa: BLE001
                extras["other_configs"] = dict(error=f"{type(e).__name__}: {e}")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.cpu_seconds, args.cpu_threads)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="BASELINE configs[2]: 64 mono tracks x 5 min per GPU, 2048-FFT/512-hop, "
                                         "pitch detect + shift +3 st (tracks shard across GPUs, no collective)",
                                tracks_per_gpu=nt, seconds_per_track=args.seconds, fft=FFT_N, hop=HOP,
                                semitones=SEMITONES, sample_rate=FS, frames_per_gpu=frames_per_rank,
                                analysis_fft="f64", synthesis_fft="f32", phase_accumulator="u32",
                                wave_mib=args.wave_mib, numa_binding=numa,
                                l2="inputs and outputs (3.7 GB each per GPU) exceed the 126 MB L2; no flush needed"),
                    clocks=clocks, e2e=e2e, e2e_f32=e2e_f32, gpu_launches=launches, roofline=roofline,
                    cpu_baseline=cpu, extras=extras)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
