"""Python mirror of the C ABI: an `Engine` owns one mlx_ctx on one GPU.

Host (numpy) entry points copy through the library; the `*_dev` entry points take torch CUDA
tensors (PyTorch is only the device-memory / stream plumbing) and launch on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import capi
from .capi import PvParams, check, ptr_array


def semitone_ratio(semitones: float) -> np.float32:
    """rate = powf(2, pitchBend / 12) evaluated in float like the reference (app.cpp:297)."""
    return np.float32(np.power(np.float32(2.0), np.float32(semitones) / np.float32(12.0), dtype=np.float32))


def num_frames(n: int, hop: int) -> int:
    return (n + hop - 1) // hop


class Engine:
    """One GPU context.  Mirrors, on the Python side, what the C++ `Spec` does with the C ABI."""

    def __init__(self, device: int = 0):
        self._L = capi.lib()
        h = C.c_void_p()
        check(self._L.mlx_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        self._keep = []  # pointer arrays of the call being issued (the C call reads them before it returns)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.mlx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ context
    def device_info(self) -> dict:
        sm, cc, mem = C.c_int(), C.c_int(), C.c_size_t()
        check(self._L.mlx_device_info(self._h, C.byref(sm), C.byref(cc), C.byref(mem)))
        return dict(sm_count=sm.value, cc=cc.value, total_mem=mem.value)

    def set_stream(self, cuda_stream: int) -> None:
        check(self._L.mlx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def use_torch_stream(self) -> None:
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self) -> None:
        check(self._L.mlx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.mlx_launch_count(self._h))

    KERNEL_KINDS = ("pv_analyze", "pv_scan", "pv_synth", "spec", "grain")

    def profile_enable(self, on: bool = True) -> None:
        check(self._L.mlx_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel: (total_ms, launches)} measured with CUDA events on the engine stream."""
        ms = (C.c_double * 5)()
        ln = (C.c_int64 * 5)()
        check(self._L.mlx_profile_read(self._h, ms, ln, int(reset)))
        return {k: (ms[i], int(ln[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    # ------------------------------------------------------------------ tracks
    def upload_tracks(self, tracks: Sequence[np.ndarray]) -> None:
        arrs = [np.ascontiguousarray(t, np.float32) for t in tracks]
        ptrs = ptr_array([a.ctypes.data for a in arrs])
        ns = (C.c_int64 * len(arrs))(*[a.size for a in arrs])
        check(self._L.mlx_upload_tracks(self._h, ptrs, ns, len(arrs)))
        self._lens = [a.size for a in arrs]

    def upload_tracks_dev(self, tensors) -> None:
        """tensors: list of 1-D float32 CUDA torch tensors (device-to-device copy)."""
        ptrs = ptr_array([t.data_ptr() for t in tensors])
        ns = (C.c_int64 * len(tensors))(*[t.numel() for t in tensors])
        check(self._L.mlx_upload_tracks_dev(self._h, ptrs, ns, len(tensors)))
        self._lens = [t.numel() for t in tensors]

    @property
    def track_lens(self):
        return list(self._lens)

    # ------------------------------------------------------------------ Spec path
    def spec_batch(self, track: int, fftN: int, start_end: np.ndarray) -> np.ndarray:
        """Mirror of Spec::getSpec for a list of (start,end) jobs -> [count][fftN/2] float32."""
        se = np.ascontiguousarray(start_end, np.int32).reshape(-1, 2)
        out = np.empty((se.shape[0], fftN // 2), np.float32)
        check(self._L.mlx_spec_batch(self._h, track, fftN, se.ctypes.data, se.shape[0], out.ctypes.data))
        return out

    def spec_batch_rgb(self, track: int, fftN: int, start_end: np.ndarray, k: float) -> np.ndarray:
        se = np.ascontiguousarray(start_end, np.int32).reshape(-1, 2)
        out = np.empty((se.shape[0], fftN // 2, 3), np.uint8)
        check(self._L.mlx_spec_batch_rgb(self._h, track, fftN, se.ctypes.data, se.shape[0],
                                         C.c_float(k), out.ctypes.data))
        return out

    def spec_frames_dev(self, track: int, fftN: int, hop: int, first_frame: int, count: int, out) -> None:
        """out: float32 CUDA tensor [count, fftN/2]; asynchronous on the engine stream."""
        assert out.is_cuda and out.is_contiguous() and out.numel() >= count * (fftN // 2)
        check(self._L.mlx_spec_frames_dev(self._h, track, fftN, hop, first_frame, count, out.data_ptr()))

    def spec_frames_all_dev(self, fftN: int, hop: int, outs) -> None:
        """Every uploaded track in one launch: outs[t] = float32 CUDA tensor [F_t, fftN/2]."""
        check(self._L.mlx_spec_frames_all_dev(self._h, fftN, hop, ptr_array([o.data_ptr() for o in outs])))

    # ------------------------------------------------------------------ PV path
    def _params(self, fftN, hop, rate, sample_rate, frame_begin=-1, frame_end=-1, wave_mib=0,
                phase_in=None, rate_per_frame=None) -> PvParams:
        self._keep.clear()  # the previous call has returned: its host pointer arrays were consumed
        p = PvParams()
        p.fftN, p.hop, p.rate, p.sample_rate = int(fftN), int(hop), float(rate), float(sample_rate)
        p.frame_begin, p.frame_end, p.wave_mib = int(frame_begin), int(frame_end), int(wave_mib)
        p.rate_per_frame_dev = None
        p.phase_in_dev = None
        if phase_in is not None:
            arr = ptr_array([t.data_ptr() if t is not None else None for t in phase_in])
            self._keep.append(arr)
            p.phase_in_dev = C.cast(arr, C.POINTER(C.c_void_p))
        if rate_per_frame is not None:
            arr = ptr_array([t.data_ptr() if t is not None else None for t in rate_per_frame])
            self._keep.append(arr)
            p.rate_per_frame_dev = C.cast(arr, C.POINTER(C.c_void_p))
        return p

    def pv_run(self, fftN: int, hop: int, rate: float, sample_rate: float = 48000.0, want_audio=True,
               wave_mib: int = 0, frame_begin: int = -1, frame_end: int = -1):
        """Host-buffer pipeline on the uploaded tracks -> list of dict(y, peak, f0) (numpy)."""
        p = self._params(fftN, hop, rate, sample_rate, frame_begin, frame_end, wave_mib)
        outs = []
        for n in self._lens:
            F = num_frames(n, hop)
            outs.append(dict(y=np.zeros(n, np.float32) if want_audio else None,
                             peak=np.zeros(F, np.int32), f0=np.zeros(F, np.float32)))
        pw = ptr_array([o["y"].ctypes.data if o["y"] is not None else None for o in outs])
        pp = ptr_array([o["peak"].ctypes.data for o in outs])
        pf = ptr_array([o["f0"].ctypes.data for o in outs])
        check(self._L.mlx_pv_run(self._h, C.byref(p), pw, pp, pf))
        return outs

    def pv_run_dev(self, fftN: int, hop: int, rate: float, out_wav, out_peak=None, out_f0=None,
                   sample_rate: float = 48000.0, wave_mib: int = 0, frame_begin: int = -1,
                   frame_end: int = -1, phase_in=None, rate_per_frame=None) -> None:
        """Device-resident pipeline: out_* are lists of CUDA tensors (entries may be None)."""
        p = self._params(fftN, hop, rate, sample_rate, frame_begin, frame_end, wave_mib, phase_in,
                         rate_per_frame)
        nt = len(self._lens)
        pw = ptr_array([t.data_ptr() if t is not None else None for t in (out_wav or [None] * nt)])
        pp = ptr_array([t.data_ptr() if t is not None else None for t in (out_peak or [None] * nt)])
        pf = ptr_array([t.data_ptr() if t is not None else None for t in (out_f0 or [None] * nt)])
        check(self._L.mlx_pv_run_dev(self._h, C.byref(p), pw, pp, pf))

    def pv_phase_totals_dev(self, fftN: int, hop: int, rate: float, totals, out_peak=None, out_f0=None,
                            sample_rate: float = 48000.0, frame_begin: int = -1, frame_end: int = -1,
                            wave_mib: int = 0) -> None:
        p = self._params(fftN, hop, rate, sample_rate, frame_begin, frame_end, wave_mib)
        nt = len(self._lens)
        pt = ptr_array([t.data_ptr() for t in totals])
        pp = ptr_array([t.data_ptr() if t is not None else None for t in (out_peak or [None] * nt)])
        pf = ptr_array([t.data_ptr() if t is not None else None for t in (out_f0 or [None] * nt)])
        check(self._L.mlx_pv_phase_totals_dev(self._h, C.byref(p), pt, pp, pf))

    def pv_analyze_dev(self, fftN: int, hop: int, rate: float, totals, out_peak=None, out_f0=None,
                       sample_rate: float = 48000.0, frame_begin: int = -1, frame_end: int = -1) -> None:
        """First half of the split pipeline (mlx_pv_analyze_dev): ONE analysis pass; the intermediates stay
        staged on the device, `totals` (list of int32/uint32 CUDA tensors of fftN/2+1) receive the phase
        totals of the owned frames."""
        p = self._params(fftN, hop, rate, sample_rate, frame_begin, frame_end, -1)
        nt = len(self._lens)
        pt = ptr_array([t.data_ptr() if t is not None else None for t in totals])
        pp = ptr_array([t.data_ptr() if t is not None else None for t in (out_peak or [None] * nt)])
        pf = ptr_array([t.data_ptr() if t is not None else None for t in (out_f0 or [None] * nt)])
        check(self._L.mlx_pv_analyze_dev(self._h, C.byref(p), pt, pp, pf))

    def pv_stage_export_dev(self, track: int, frame_begin: int, count: int, smag, phase) -> None:
        """After pv_analyze_dev: shifted magnitudes (float32 [count, fftN/2+1]) and accumulated synthesis
        phases (int32 view of uint32, same shape) of `count` frames of one track, into CUDA tensors."""
        check(self._L.mlx_pv_stage_export_dev(self._h, int(track), int(frame_begin), int(count), smag.data_ptr(),
                                              phase.data_ptr()))

    def pv_synth_dev(self, fftN: int, hop: int, rate: float, out_wav, sample_rate: float = 48000.0,
                     frame_begin: int = -1, frame_end: int = -1, phase_in=None) -> None:
        """Second half (mlx_pv_synth_dev): carried-in phase + synthesis on the staged analysis."""
        p = self._params(fftN, hop, rate, sample_rate, frame_begin, frame_end, -1, phase_in)
        pw = ptr_array([t.data_ptr() if t is not None else None for t in out_wav])
        check(self._L.mlx_pv_synth_dev(self._h, C.byref(p), pw))

    def pv_run_sharded_dev(self, comm: "Comm", owns, n_total: int, fftN: int, hop: int, rate: float, out_own,
                           peak_own=None, f0_own=None, sample_rate: float = 48000.0) -> None:
        """One long file by time range across the ranks of `comm` (mlx_pv_run_sharded_dev): owns / out_own
        are lists (one per planar channel) of CUDA tensors holding this rank's owned samples."""
        p = self._params(fftN, hop, rate, sample_rate)
        nt = len(owns)
        po = ptr_array([t.data_ptr() for t in owns])
        pw = ptr_array([t.data_ptr() if t is not None else None for t in (out_own or [None] * nt)])
        pp = ptr_array([t.data_ptr() if t is not None else None for t in (peak_own or [None] * nt)])
        pf = ptr_array([t.data_ptr() if t is not None else None for t in (f0_own or [None] * nt)])
        check(self._L.mlx_pv_run_sharded_dev(self._h, comm._h, C.byref(p), po, nt, int(n_total),
                                             pw if out_own else None, pp if peak_own else None,
                                             pf if f0_own else None))
        self._lens = [int(self._L.mlx_track_len(self._h, t)) for t in range(nt)]

    def pv_process_host(self, tracks, fftN: int, hop: int, rate: float, out_wav, out_peak=None,
                        out_f0=None, sample_rate: float = 48000.0, wave_mib: int = 0) -> None:
        """End to end from host buffers (numpy arrays or pinned torch CPU tensors) into host buffers.
        float32 or int16 PCM on either side (mlx_pv_process_host_fmt): int16 in = s / 32768, int16 out =
        int16(x * 32767.) by truncation, the reference's export conversion (app.cpp:1209-1212)."""
        def addr(a):
            if a is None:
                return None
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()

        def size(a):
            return a.size if isinstance(a, np.ndarray) else a.numel()

        def fmt(seq):
            kinds = {str(a.dtype).replace("torch.", "") for a in seq if a is not None}
            if kinds <= {"float32"}:
                return 0
            if kinds == {"int16"}:
                return 1
            raise TypeError(f"tracks must be all float32 or all int16, got {sorted(kinds)}")

        p = self._params(fftN, hop, rate, sample_rate, -1, -1, wave_mib)
        nt = len(tracks)
        pin = ptr_array([addr(t) for t in tracks])
        ns = (C.c_int64 * nt)(*[size(t) for t in tracks])
        pw = ptr_array([addr(t) for t in (out_wav or [None] * nt)])
        pp = ptr_array([addr(t) for t in (out_peak or [None] * nt)])
        pf = ptr_array([addr(t) for t in (out_f0 or [None] * nt)])
        check(self._L.mlx_pv_process_host_fmt(self._h, C.byref(p), pin, fmt(tracks), ns, nt, pw,
                                              fmt(out_wav or []), pp, pf))
        self._lens = [size(t) for t in tracks]

    # ------------------------------------------------------------------ grain path
    # ------------------------------------------------------------------ waveform pyramid
    def picks_layout(self, n: int) -> np.ndarray:
        """level_off[levels + 1] (in pairs) of the min/max pyramid of an n-sample track (app.cpp:347-378)."""
        L = int(self._L.mlx_picks_levels(n))
        off = np.zeros(L + 1, np.int64)
        self._L.mlx_picks_layout(n, off.ctypes.data)
        return off

    def picks_build(self, track: int):
        """(pairs [total, 2] float32, level_off) of an uploaded track -- App::calcPicks on the GPU."""
        off = self.picks_layout(int(self._L.mlx_track_len(self._h, track)))
        pairs = np.zeros((max(int(off[-1]), 1), 2), np.float32)
        check(self._L.mlx_picks_build(self._h, track, pairs.ctypes.data))
        return pairs[:int(off[-1])], off

    def picks_build_dev(self, track: int, pairs) -> None:
        assert pairs.is_cuda and pairs.is_contiguous()
        check(self._L.mlx_picks_build_dev(self._h, track, pairs.data_ptr()))

    def picks_build_all_dev(self, pairs_list) -> None:
        """One launch for all uploaded tracks; pairs_list[t]: float32 CUDA tensor [total_t, 2]."""
        check(self._L.mlx_picks_build_all_dev(self._h, ptr_array([p.data_ptr() for p in pairs_list])))

    def minmax_ranges(self, track: int, ranges) -> np.ndarray:
        """App::getMinMaxFromRange (app.cpp:380-426) for every (start, end) row of `ranges`."""
        r = np.ascontiguousarray(ranges, np.int32).reshape(-1, 2)
        out = np.zeros((r.shape[0], 2), np.float32)
        check(self._L.mlx_minmax_ranges(self._h, track, r.ctypes.data, r.shape[0], out.ctypes.data))
        return out

    def grain_segment(self, cap: int | None = None):
        """Zero-crossing grain segmentation of every uploaded track on the GPU (App::preproc,
        reference app.cpp:156-235).  Returns [(g_start, g_len)] per track (int32 arrays)."""
        nt = int(self._L.mlx_num_tracks(self._h))
        lens = [int(self._L.mlx_track_len(self._h, t)) for t in range(nt)]
        if cap is None:
            cap = max(lens) // 751 + 1 if lens else 1
        gs = np.zeros((nt, max(cap, 1)), np.int32)
        gl = np.zeros((nt, max(cap, 1)), np.int32)
        counts = np.zeros(nt, np.int32)
        check(self._L.mlx_grain_segment(self._h, ptr_array([gs[t].ctypes.data for t in range(nt)]),
                                        ptr_array([gl[t].ctypes.data for t in range(nt)]), cap, counts.ctypes.data))
        self.last_grain_counts = counts
        return [(gs[t, :min(int(counts[t]), cap)].copy(), gl[t, :min(int(counts[t]), cap)].copy()) for t in range(nt)]

    def grain_segment_dev(self, g_start, g_len, counts, cap: int) -> None:
        """g_start / g_len: int32 CUDA tensors [ntracks, cap]; counts: int32 CUDA tensor [ntracks]."""
        nt = int(self._L.mlx_num_tracks(self._h))
        assert g_start.is_cuda and g_len.is_cuda and counts.is_cuda and g_start.shape[0] == nt
        check(self._L.mlx_grain_segment_dev(self._h, ptr_array([g_start[t].data_ptr() for t in range(nt)]),
                                            ptr_array([g_len[t].data_ptr() for t in range(nt)]), cap,
                                            counts.data_ptr()))

    def grain_render(self, track: int, g_start, g_len, g_rate, out_off, g_next, tail_zeros: int = 1500,
                     want_i16: bool = True):
        gs = np.ascontiguousarray(g_start, np.int32)
        gl = np.ascontiguousarray(g_len, np.int32)
        gr = np.ascontiguousarray(g_rate, np.float32)
        oo = np.ascontiguousarray(out_off, np.int64)
        gn = np.ascontiguousarray(g_next, np.float32)
        ng = gs.size
        assert oo.size == ng + 1
        total = int(oo[-1]) + tail_zeros
        out = np.zeros(total, np.float32)
        out16 = np.zeros(total, np.int16) if want_i16 else None
        check(self._L.mlx_grain_render(self._h, track, gs.ctypes.data, gl.ctypes.data, gr.ctypes.data,
                                       oo.ctypes.data, gn.ctypes.data, ng, tail_zeros, out.ctypes.data,
                                       out16.ctypes.data if want_i16 else None))
        return out, out16


class Comm:
    """NCCL communicator behind the C ABI (mlx_comm): one per rank, bound to the rank's Engine."""

    def __init__(self, engine: Engine, unique_id: bytes, world: int, rank: int):
        assert len(unique_id) == 128
        self._L = engine._L
        self.engine = engine
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(self._L.mlx_comm_create(C.byref(h), engine._h, buf, int(world), int(rank)))
        self._h = h
        self.world, self.rank = int(world), int(rank)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(capi.lib().mlx_comm_unique_id(buf))
        return buf.raw

    @property
    def nccl_version(self) -> int:
        v = C.c_int()
        check(self._L.mlx_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.mlx_comm_destroy(self._h)
            self._h = None
