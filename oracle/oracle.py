"""ctypes view of the CPU oracle (oracle/_build/liboracle.so) and of the reference's own Spec
compiled unmodified (oracle/_ref/libspec_ref.so).  TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "liboracle.so"
_REF = _HERE / "_ref" / "libspec_ref.so"
_REF_APP = _HERE / "_ref" / "libapp_ref.so"
_DROPIN_APP = _HERE / "_ref" / "libapp_dropin.so"

REF_SPECTR_SIZE = 32768  # reference spec.cpp:8


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    del force  # make is cheap when everything is up to date, and a stale library must never be used
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


_lib = None
_ref = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class PvParams(C.Structure):
    _fields_ = [("N", C.c_int), ("hop", C.c_int), ("fs", C.c_double), ("rate", C.c_float),
                ("rate_per_frame", C.c_void_p)]


class Marker(C.Structure):
    _fields_ = [("sample", C.c_int), ("note", C.c_double), ("dTime", C.c_double),
                ("pitchBend", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        if not _LIB.exists():
            build()
        _lib = C.CDLL(str(_LIB))
        _lib.mlxo_pv_num_frames.restype = C.c_int64
        _lib.mlxo_pv_num_frames.argtypes = [C.c_int64, C.c_int]
        _lib.mlxo_grain_export.restype = C.c_int64
        _lib.mlxo_sample2time.restype = C.c_double
        _lib.mlxo_duration.restype = C.c_double
        _lib.mlxo_time2pitchbend.restype = C.c_float
    return _lib


def have_ref() -> bool:
    return _REF.exists()


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(str(_REF))
    return _ref


def num_threads() -> int:
    return int(lib().mlxo_num_threads())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- Spec
def spec_batch(wav: np.ndarray, N: int, start_end: np.ndarray, nthreads: int = 0) -> np.ndarray:
    wav = np.ascontiguousarray(wav, np.float32)
    se = np.ascontiguousarray(start_end, np.int32).reshape(-1, 2)
    out = np.empty((se.shape[0], N // 2), np.float32)
    rc = lib().mlxo_spec_batch(_ptr(wav), C.c_int64(wav.size), C.c_int(N), _ptr(se),
                               C.c_int(se.shape[0]), _ptr(out), C.c_int(nthreads))
    if rc:
        raise RuntimeError("mlxo_spec_batch failed")
    return out


def ref_spec_run(wav: np.ndarray, start_end: np.ndarray) -> np.ndarray:
    """The reference's own Spec (N = 32768) driven through getSpec()."""
    wav = np.ascontiguousarray(wav, np.float32)
    se = np.ascontiguousarray(start_end, np.int32).reshape(-1, 2)
    out = np.zeros((se.shape[0], REF_SPECTR_SIZE // 2), np.float32)
    ln = ref().mlxo_ref_spec_run(_ptr(wav), C.c_longlong(wav.size), _ptr(se), C.c_int(se.shape[0]),
                                 _ptr(out))
    if ln != REF_SPECTR_SIZE // 2:
        raise RuntimeError(f"reference Spec returned length {ln}")
    return out


def colormap(spec: np.ndarray, k: float) -> np.ndarray:
    spec = np.ascontiguousarray(spec, np.float32)
    out = np.empty(spec.shape + (3,), np.uint8)
    lib().mlxo_colormap(_ptr(spec), C.c_int(spec.size), C.c_float(k), _ptr(out))
    return out


# ---------------------------------------------------------------- PV
def pv_num_frames(n: int, hop: int) -> int:
    return int(lib().mlxo_pv_num_frames(n, hop))


def pv_run(x: np.ndarray, N: int, hop: int, rate: float, fs: float = 48000.0,
           rate_per_frame: np.ndarray | None = None, want_debug: bool = False,
           want_audio: bool = True):
    """Returns dict(y, peak, f0, margin[, inc, smag])."""
    x = np.ascontiguousarray(x, np.float32)
    n = x.size
    F = pv_num_frames(n, hop)
    nb = N // 2 + 1
    y = np.zeros(n, np.float32) if want_audio else None
    peak = np.zeros(F, np.int32)
    f0 = np.zeros(F, np.float32)
    margin = np.zeros(F, np.float64)
    inc = np.zeros((F, nb), np.uint32) if want_debug else None
    smag = np.zeros((F, nb), np.float32) if want_debug else None
    rpf = None
    if rate_per_frame is not None:
        rpf = np.ascontiguousarray(rate_per_frame, np.float32)
        assert rpf.size == F
    p = PvParams(N, hop, fs, np.float32(rate), _ptr(rpf))
    rc = lib().mlxo_pv_run(_ptr(x), C.c_int64(n), C.byref(p), _ptr(y), _ptr(peak), _ptr(f0),
                           _ptr(margin), _ptr(inc), _ptr(smag))
    if rc:
        raise RuntimeError("mlxo_pv_run failed (N must be a power of two, hop = N/4)")
    out = dict(y=y, peak=peak, f0=f0, margin=margin)
    if want_debug:
        out.update(inc=inc, smag=smag)
    return out


def pv_run_batch(x: np.ndarray, N: int, hop: int, rate: float, fs: float = 48000.0,
                 nthreads: int = 0):
    x = np.ascontiguousarray(x, np.float32)
    ntracks, n = x.shape
    F = pv_num_frames(n, hop)
    y = np.zeros_like(x)
    peak = np.zeros((ntracks, F), np.int32)
    f0 = np.zeros((ntracks, F), np.float32)
    p = PvParams(N, hop, fs, np.float32(rate), None)
    rc = lib().mlxo_pv_run_batch(_ptr(x), C.c_int64(n), C.c_int(ntracks), C.byref(p), _ptr(y),
                                 _ptr(peak), _ptr(f0), C.c_int(nthreads))
    if rc:
        raise RuntimeError("mlxo_pv_run_batch failed")
    return dict(y=y, peak=peak, f0=f0)


def pv_gather_range(j: int, r: float, nbins: int):
    lo, hi = C.c_int(), C.c_int()
    lib().mlxo_pv_gather_range(C.c_int(j), C.c_float(r), C.c_int(nbins), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


# ---------------------------------------------------------------- grains
def _markers(markers):
    arr = (Marker * max(1, len(markers)))()
    for i, m in enumerate(markers):
        arr[i] = Marker(int(m[0]), float(m[1]), float(m[2]), float(m[3]))
    return arr, len(markers)


def grain_segment(wav: np.ndarray):
    wav = np.ascontiguousarray(wav, np.float32)
    cap = max(16, wav.size // 700 + 16)
    gs = np.zeros(cap, np.int32)
    gl = np.zeros(cap, np.int32)
    ng = lib().mlxo_grain_segment(_ptr(wav), C.c_int64(wav.size), _ptr(gs), _ptr(gl), C.c_int(cap))
    assert ng <= cap
    return gs[:ng].copy(), gl[:ng].copy()


def picks_build(wav: np.ndarray):
    """calcPicks (app.cpp:347-378): returns (pairs [total, 2] float32, level_off [levels + 1] int64)."""
    wav = np.ascontiguousarray(wav, np.float32)
    L = lib().mlxo_picks_levels(C.c_int64(wav.size))
    off = np.zeros(L + 1, np.int64)
    lib().mlxo_picks_layout.restype = C.c_int64
    total = lib().mlxo_picks_layout(C.c_int64(wav.size), off.ctypes.data_as(C.c_void_p))
    pairs = np.zeros((max(total, 1), 2), np.float32)
    lib().mlxo_picks_build(_ptr(wav), C.c_int64(wav.size), _ptr(pairs), off.ctypes.data_as(C.c_void_p))
    return pairs[:total], off


def minmax_ranges(wav: np.ndarray, pairs: np.ndarray, level_off: np.ndarray, ranges: np.ndarray) -> np.ndarray:
    """getMinMaxFromRange (app.cpp:380-426) for every (start, end) row of `ranges`."""
    wav = np.ascontiguousarray(wav, np.float32)
    pairs = np.ascontiguousarray(pairs, np.float32)
    ranges = np.ascontiguousarray(ranges, np.int32)
    out = np.zeros((ranges.shape[0], 2), np.float32)
    pp = pairs if pairs.size else np.zeros((1, 2), np.float32)
    lib().mlxo_minmax_ranges(_ptr(wav) if wav.size else None, C.c_int64(wav.size), _ptr(pp),
                             np.ascontiguousarray(level_off, np.int64).ctypes.data_as(C.c_void_p), _ptr(ranges),
                             C.c_int(ranges.shape[0]), _ptr(out))
    return out


def time2sample(markers, sr: int, val: float) -> int:
    arr, nm = _markers(markers)
    return int(lib().mlxo_time2sample(arr, C.c_int(nm), C.c_int(sr), C.c_double(val)))


def sample2time(markers, sr: int, val: int) -> float:
    arr, nm = _markers(markers)
    return float(lib().mlxo_sample2time(arr, C.c_int(nm), C.c_int(sr), C.c_int(val)))


def time2pitchbend(markers, sr: int, n: int, val: float) -> float:
    arr, nm = _markers(markers)
    return float(lib().mlxo_time2pitchbend(arr, C.c_int(nm), C.c_int(sr), C.c_int64(n),
                                           C.c_double(val)))


def grain_export(wav: np.ndarray, sr: int, markers, g_start=None, g_len=None):
    """Returns dict(pcm, pcm16, schedule=dict(gstart, glen, rate, out_off, next))."""
    wav = np.ascontiguousarray(wav, np.float32)
    if g_start is None:
        g_start, g_len = grain_segment(wav)
    g_start = np.ascontiguousarray(g_start, np.int32)
    g_len = np.ascontiguousarray(g_len, np.int32)
    arr, nm = _markers(markers)
    cap = int(wav.size * 4.5 + 4096)
    pcm = np.zeros(cap, np.float32)
    pcm16 = np.zeros(cap, np.int16)
    caps = g_start.size * 6 + 16
    sg = np.zeros(caps, np.int32)
    sl = np.zeros(caps, np.int32)
    srate = np.zeros(caps, np.float32)
    soff = np.zeros(caps, np.int64)
    snext = np.zeros(caps, np.float32)
    ns = C.c_int()
    ln = lib().mlxo_grain_export(_ptr(wav), C.c_int64(wav.size), C.c_int(sr), arr, C.c_int(nm),
                                 _ptr(g_start), _ptr(g_len), C.c_int(g_start.size), _ptr(pcm),
                                 _ptr(pcm16), C.c_int64(cap), _ptr(sg), _ptr(sl), _ptr(srate),
                                 _ptr(soff), _ptr(snext), C.byref(ns), C.c_int(caps))
    assert ln <= cap and ns.value <= caps, (ln, cap, ns.value, caps)
    k = ns.value
    return dict(pcm=pcm[:ln].copy(), pcm16=pcm16[:ln].copy(),
                schedule=dict(gstart=sg[:k].copy(), glen=sl[:k].copy(), rate=srate[:k].copy(),
                              out_off=soff[:k].copy(), next=snext[:k].copy()))


# ---------------------------------------------------------------- the reference's own App (oracle/_ref)
def have_ref_app() -> bool:
    """oracle/_ref/libapp_ref.so: /root/reference/{app,spec-cache,save-wav,spec}.cpp compiled unmodified
    against the no-op UI / audio / codec headers of oracle/shim_app (driver: oracle/ref_app.cpp)."""
    return _REF_APP.exists()


def have_dropin_app() -> bool:
    """oracle/_ref/libapp_dropin.so: the reference's front-end (app.cpp, save-wav.cpp, unmodified) built on
    the PRODUCT's Spec / SpecCache (melonix_b200/host) and libmelonix_b200.so -- INTEGRATION.md section 1."""
    return _DROPIN_APP.exists()


_ref_app = {}


def _ref_app_lib(which: str = "ref"):
    if which not in _ref_app:
        L = C.CDLL(str(_REF_APP if which == "ref" else _DROPIN_APP))
        vp = C.c_void_p
        L.mlxo_ref_app_create.restype = vp
        L.mlxo_ref_app_create.argtypes = [vp, C.c_longlong, C.c_int, vp, C.c_int]
        L.mlxo_ref_app_last_error.restype = C.c_char_p
        L.mlxo_ref_app_destroy.argtypes = [vp]
        L.mlxo_ref_app_grains.argtypes = [vp, vp, vp, C.c_int]
        L.mlxo_ref_app_picks.argtypes = [vp, vp, C.c_longlong, vp]
        L.mlxo_ref_app_minmax.argtypes = [vp, vp, C.c_int, vp]
        L.mlxo_ref_app_sample2time.argtypes = [vp, C.c_int]
        L.mlxo_ref_app_sample2time.restype = C.c_double
        L.mlxo_ref_app_time2sample.argtypes = [vp, C.c_double]
        L.mlxo_ref_app_duration.argtypes = [vp]
        L.mlxo_ref_app_duration.restype = C.c_double
        L.mlxo_ref_app_time2pitchbend.argtypes = [vp, C.c_double]
        L.mlxo_ref_app_time2pitchbend.restype = C.c_float
        L.mlxo_ref_app_export.argtypes = [vp, C.c_char_p]
        L.mlxo_ref_app_render.argtypes = [vp, vp, C.c_longlong]
        L.mlxo_ref_app_render.restype = C.c_longlong
        L.mlxo_ref_app_speccache_column.argtypes = [vp, C.c_float, C.c_int, C.c_double, C.c_double, vp, C.c_int]
        _ref_app[which] = L
    return _ref_app[which]


class RefApp:
    """The reference's `App` after `preproc()` on a synthetic track (no UI, no audio device, no decoder)."""

    def __init__(self, wav: np.ndarray, sr: int, markers=(), build: str = "ref"):
        """build = "ref": everything is the reference's code (CPU).  build = "dropin": the reference's
        front-end on the product's Spec / SpecCache (needs a B200)."""
        self._L = _ref_app_lib(build)
        self.wav = np.ascontiguousarray(wav, np.float32)
        arr, nm = _markers(list(markers))
        self._h = self._L.mlxo_ref_app_create(_ptr(self.wav), self.wav.size, int(sr), C.cast(arr, C.c_void_p), nm)
        if not self._h:
            raise RuntimeError(self._L.mlxo_ref_app_last_error().decode(errors="replace"))

    def close(self):
        if self._h:
            self._L.mlxo_ref_app_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def grains(self):
        cap = self.wav.size // 700 + 16
        gs, gl = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n = self._L.mlxo_ref_app_grains(self._h, _ptr(gs), _ptr(gl), cap)
        assert n <= cap
        return gs[:n].copy(), gl[:n].copy()

    def picks(self):
        cap = self.wav.size + 8
        pairs = np.zeros((cap, 2), np.float32)
        off = np.zeros(40, np.int64)
        L = self._L.mlxo_ref_app_picks(self._h, _ptr(pairs), cap, off.ctypes.data_as(C.c_void_p))
        return pairs[:int(off[L])].copy(), off[:L + 1].copy()

    def minmax_ranges(self, ranges):
        r = np.ascontiguousarray(ranges, np.int32).reshape(-1, 2)
        out = np.zeros((r.shape[0], 2), np.float32)
        self._L.mlxo_ref_app_minmax(self._h, _ptr(r), r.shape[0], _ptr(out))
        return out

    def sample2time(self, s):
        return self._L.mlxo_ref_app_sample2time(self._h, int(s))

    def time2sample(self, t):
        return self._L.mlxo_ref_app_time2sample(self._h, float(t))

    def duration(self):
        return self._L.mlxo_ref_app_duration(self._h)

    def time2pitchbend(self, t):
        return float(self._L.mlxo_ref_app_time2pitchbend(self._h, float(t)))

    def render(self):
        """float samples of exportWav's process() loop (app.cpp:1201-1207, 294-345)"""
        pcm = np.zeros(self.wav.size * 5 + 4096, np.float32)
        n = self._L.mlxo_ref_app_render(self._h, _ptr(pcm), pcm.size)
        assert n <= pcm.size
        return pcm[:n].copy()

    def export_wav(self, path) -> np.ndarray:
        """App::exportWav -> saveWav; returns the int16 samples found in the file.  NOTE: the reference's
        saveWav patches the data-chunk size with an 8-byte write (save-wav.cpp:42-43), which zeroes the
        first two samples of the file; callers compare from sample 2 on."""
        self._L.mlxo_ref_app_export(self._h, str(path).encode())
        raw = Path(path).read_bytes()
        assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
        return np.frombuffer(raw[44:], np.int16).copy()

    def speccache_column(self, k: float, width: int, range_time: float, t: float) -> np.ndarray:
        """RGB texels SpecCache::populateTex hands to glTexImage1D for the column at time t."""
        rgb = np.zeros((REF_SPECTR_SIZE // 2, 3), np.uint8)
        n = self._L.mlxo_ref_app_speccache_column(self._h, float(k), int(width), float(range_time), float(t),
                                                  _ptr(rgb), rgb.shape[0])
        assert n == rgb.shape[0], n
        return rgb
