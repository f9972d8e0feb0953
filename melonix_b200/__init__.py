"""melonix_b200 -- B200-native (sm_100a) implementation of melonix's Spec STFT / pitch-shift hot path.

Everything numeric runs in hand-written CUDA kernels behind the C ABI of include/melonix_gpu.h
(libmelonix_b200.so).  This package is the Python-side mirror used by the tests, bench.py and
multi-GPU drivers; the C++ mirror of the reference's `Spec` / `SpecCache` classes lives in
melonix_b200/host/.  There is no CPU fallback anywhere in this package.
"""
from .capi import MlxError, build, declared_symbols  # noqa: F401
from .engine import Engine, semitone_ratio  # noqa: F401

__all__ = ["Engine", "MlxError", "build", "declared_symbols", "semitone_ratio"]
