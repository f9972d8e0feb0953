"""ctypes binding of include/melonix_gpu.h (libmelonix_b200.so, built in-tree by melonix_b200/csrc/Makefile).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present,
loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import re
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os

# MELONIX_B200_LIB: alternative build of the same library (kernel tuning experiments only)
LIB_PATH = Path(os.environ.get("MELONIX_B200_LIB", _PKG / "libmelonix_b200.so"))
HEADER_PATH = _PKG.parent / "include" / "melonix_gpu.h"

MLX_OK = 0


class MlxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"melonix_b200 error {code}: {msg}")
        self.code = code


class PvParams(C.Structure):
    _fields_ = [
        ("fftN", C.c_int),
        ("hop", C.c_int),
        ("rate", C.c_float),
        ("sample_rate", C.c_double),
        ("rate_per_frame_dev", C.POINTER(C.c_void_p)),
        ("frame_begin", C.c_int64),
        ("frame_end", C.c_int64),
        ("phase_in_dev", C.POINTER(C.c_void_p)),
        ("wave_mib", C.c_int),
    ]


class TimeShardC(C.Structure):
    _fields_ = [("frame_begin", C.c_int64), ("frame_end", C.c_int64), ("own_lo", C.c_int64), ("own_hi", C.c_int64),
                ("need_lo", C.c_int64), ("need_hi", C.c_int64), ("frame_offset", C.c_int64)]


def build(force: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not LIB_PATH.exists():
        subprocess.run(["make", "-C", str(_PKG / "csrc"), "-j8"], check=True, capture_output=True)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function include/melonix_gpu.h declares."""
    txt = HEADER_PATH.read_text()
    return sorted(set(re.findall(r"MLX_API[^;(]*?\b(mlx_\w+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(melonix_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    L.mlx_create.argtypes = [C.POINTER(vp), i32]
    L.mlx_destroy.argtypes = [vp]
    L.mlx_destroy.restype = None
    L.mlx_last_error.restype = C.c_char_p
    L.mlx_set_stream.argtypes = [vp, vp]
    L.mlx_sync.argtypes = [vp]
    L.mlx_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t)]
    L.mlx_launch_count.argtypes = [vp]
    L.mlx_launch_count.restype = i64
    L.mlx_profile_enable.argtypes = [vp, i32]
    L.mlx_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), i32]
    L.mlx_upload_tracks.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32]
    L.mlx_upload_tracks_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32]
    L.mlx_num_tracks.argtypes = [vp]
    L.mlx_track_len.argtypes = [vp, i32]
    L.mlx_track_len.restype = i64
    L.mlx_spec_batch.argtypes = [vp, i32, i32, vp, i32, vp]
    L.mlx_spec_batch_dev.argtypes = [vp, i32, i32, vp, i32, vp]
    L.mlx_spec_frames_dev.argtypes = [vp, i32, i32, i32, i64, i64, vp]
    L.mlx_spec_frames_all_dev.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.mlx_spec_batch_rgb.argtypes = [vp, i32, i32, vp, i32, f32, vp]
    L.mlx_pv_run.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_run_dev.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_phase_totals_dev.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_analyze_dev.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_synth_dev.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp)]
    L.mlx_pv_stage_export_dev.argtypes = [vp, i32, i64, i64, vp, vp]
    L.mlx_shard_frames.argtypes = [i64, i32, i32, i32, i32, C.POINTER(TimeShardC)]
    L.mlx_comm_unique_id.argtypes = [vp]
    L.mlx_comm_create.argtypes = [C.POINTER(vp), vp, vp, i32, i32]
    L.mlx_comm_destroy.argtypes = [vp]
    L.mlx_comm_destroy.restype = None
    L.mlx_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.mlx_pv_run_sharded_dev.argtypes = [vp, vp, C.POINTER(PvParams), C.POINTER(vp), i32, i64,
                                         C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_run_sharded.argtypes = L.mlx_pv_run_sharded_dev.argtypes
    L.mlx_pv_process_host.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), C.POINTER(i64), i32,
                                      C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mlx_pv_process_host_fmt.argtypes = [vp, C.POINTER(PvParams), C.POINTER(vp), i32, C.POINTER(i64), i32,
                                          C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(vp)]
    L.mlx_grain_render.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.mlx_picks_levels.argtypes = [i64]
    L.mlx_picks_layout.argtypes = [i64, vp]
    L.mlx_picks_layout.restype = i64
    L.mlx_picks_build.argtypes = [vp, i32, vp]
    L.mlx_picks_build_dev.argtypes = [vp, i32, vp]
    L.mlx_picks_build_all_dev.argtypes = [vp, C.POINTER(vp)]
    L.mlx_minmax_ranges.argtypes = [vp, i32, vp, i32, vp]
    L.mlx_grain_segment.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32, vp]
    L.mlx_grain_segment_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != MLX_OK:
        raise MlxError(rc, lib().mlx_last_error().decode(errors="replace"))


def ptr_array(ptrs) -> C.Array:
    """list of int addresses (or None) -> void*[]"""
    arr = (C.c_void_p * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr
