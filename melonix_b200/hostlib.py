"""ctypes binding of include/melonix_host.h (libmelonix_host.so): the host-side, serial parts of the
grain path -- zero-crossing segmentation, marker warp maps, exportWav's cursor recurrence -- that
produce the render schedule for Engine.grain_render (mlx_grain_render).  Pure host code, product
side; it does not use the oracle."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmelonix_host.so"
HEADER_PATH = _PKG.parent / "include" / "melonix_host.h"


class Marker(C.Structure):
    _fields_ = [("sample", C.c_int), ("note", C.c_double), ("dTime", C.c_double), ("pitchBend", C.c_double)]


def declared_symbols() -> list[str]:
    return sorted(set(re.findall(r"MLXH_API[^;(]*?\b(mlxh_\w+)\s*\(", HEADER_PATH.read_text())))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.mlxh_sample2time.restype = C.c_double
        L.mlxh_duration.restype = C.c_double
        L.mlxh_time2pitchbend.restype = C.c_float
        L.mlxh_picks_layout.restype = C.c_int64
        _lib = L
    return _lib


def _markers(markers):
    """markers: iterable of (sample, note, dTime, pitchBend), sorted by sample."""
    arr = (Marker * max(1, len(markers)))()
    for i, m in enumerate(markers):
        arr[i] = Marker(int(m[0]), float(m[1]), float(m[2]), float(m[3]))
    return arr, len(markers)


def grain_segment(wav: np.ndarray):
    wav = np.ascontiguousarray(wav, np.float32)
    cap = max(16, wav.size // 700 + 16)
    gs, gl = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    ng = lib().mlxh_grain_segment(wav.ctypes.data_as(C.c_void_p), C.c_int64(wav.size),
                                  gs.ctypes.data_as(C.c_void_p), gl.ctypes.data_as(C.c_void_p), C.c_int(cap))
    assert ng <= cap
    return gs[:ng].copy(), gl[:ng].copy()


def time2sample(markers, sr: int, t: float) -> int:
    arr, nm = _markers(markers)
    return int(lib().mlxh_time2sample(arr, C.c_int(nm), C.c_int(sr), C.c_double(t)))


def sample2time(markers, sr: int, sample: int) -> float:
    arr, nm = _markers(markers)
    return float(lib().mlxh_sample2time(arr, C.c_int(nm), C.c_int(sr), C.c_int(sample)))


def time2pitchbend(markers, sr: int, n: int, t: float) -> float:
    arr, nm = _markers(markers)
    return float(lib().mlxh_time2pitchbend(arr, C.c_int(nm), C.c_int(sr), C.c_int64(n), C.c_double(t)))


def export_schedule(wav: np.ndarray, sr: int, markers, g_start, g_len):
    """dict(gstart, glen, rate, out_off[rows+1], next, tail_zeros) for Engine.grain_render."""
    wav = np.ascontiguousarray(wav, np.float32)
    g_start = np.ascontiguousarray(g_start, np.int32)
    g_len = np.ascontiguousarray(g_len, np.int32)
    arr, nm = _markers(markers)
    cap = g_start.size * 6 + 16
    while True:
        sg, sl = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        srate, snext = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        soff = np.zeros(cap + 1, np.int64)
        tail = C.c_int()
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rows = lib().mlxh_export_schedule(p(wav), C.c_int64(wav.size), C.c_int(sr), arr, C.c_int(nm), p(g_start),
                                          p(g_len), C.c_int(g_start.size), p(sg), p(sl), p(srate), p(soff),
                                          p(snext), C.c_int(cap), C.byref(tail))
        if rows >= 0:
            break
        cap = -rows + 16
    return dict(gstart=sg[:rows].copy(), glen=sl[:rows].copy(), rate=srate[:rows].copy(),
                out_off=soff[:rows + 1].copy(), next=snext[:rows].copy(), tail_zeros=tail.value)


def export_wav(engine, track: int, wav: np.ndarray, sr: int, markers):
    """GPU mirror of App::exportWav (reference app.cpp:1194-1215): returns (pcm float32, pcm int16)."""
    gs, gl = grain_segment(wav)
    s = export_schedule(wav, sr, markers, gs, gl)
    return engine.grain_render(track, s["gstart"], s["glen"], s["rate"], s["out_off"], s["next"],
                               tail_zeros=s["tail_zeros"])


def picks_layout(n: int) -> np.ndarray:
    """level_off[levels + 1] of the min/max pyramid of an n-sample track (reference app.cpp:352-369)."""
    L = lib().mlxh_picks_levels(C.c_int64(n))
    off = np.zeros(L + 1, np.int64)
    lib().mlxh_picks_layout(C.c_int64(n), off.ctypes.data_as(C.c_void_p))
    return off


def minmax_ranges(wav: np.ndarray, pairs: np.ndarray, ranges) -> np.ndarray:
    """App::getMinMaxFromRange (app.cpp:380-426) on the host, from a pyramid built by Engine.picks_build."""
    wav = np.ascontiguousarray(wav, np.float32)
    pairs = np.ascontiguousarray(pairs, np.float32)
    r = np.ascontiguousarray(ranges, np.int32).reshape(-1, 2)
    out = np.zeros((r.shape[0], 2), np.float32)
    lib().mlxh_minmax_ranges(wav.ctypes.data_as(C.c_void_p), C.c_int64(wav.size), pairs.ctypes.data_as(C.c_void_p),
                             r.ctypes.data_as(C.c_void_p), C.c_int(r.shape[0]), out.ctypes.data_as(C.c_void_p))
    return out


def colour_ramp(mag: np.ndarray, k: float) -> np.ndarray:
    """SpecCache::populateTex's colour ramp (reference spec-cache.cpp:77-96) on the host: [..., 3] uint8.
    The drop-in Spec recolours cached float columns with it when the brightness gain changes."""
    mag = np.ascontiguousarray(mag, np.float32)
    out = np.zeros(mag.shape + (3,), np.uint8)
    lib().mlxh_colour_ramp(mag.ctypes.data_as(C.c_void_p), C.c_int(mag.size), C.c_float(k),
                           out.ctypes.data_as(C.c_void_p))
    return out
