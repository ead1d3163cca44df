"""spectral — SURVEY §8(f) "next" rows built on the device FFT:

* `spectrogram`: the compute core of `util::plot::waterfall` / `spectrum` (src/util/plot.rs:46-68,
  :109-130) — per fft_len chunk `vec_rfft(Scale::SN).vec_mirror()`, then `c.norm()` and optionally
  `DB::from(..).db()` (src/util/mod.rs:26-34) — fused into one kernel (8 B in, 4 B out per sample).
  Plotting itself (gnuplot) stays out of scope.
* `correlate`: the frequency-domain correlator the crate benchmarks (benches/benches.rs:410-416),
  `input.vec_rfft(&mut fft, s).vec_mul(&sig).vec_rifft(&mut fft, s)`, one kernel per batch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import call, lib
from .fft import Cfft, Scale
from .vecops import DeviceVec


class DeviceF32:
    """Device `Vec<f32>`."""

    def __init__(self, n: int = 0):
        h = C.c_void_p()
        call("ae_f32_alloc", n, C.byref(h))
        self._h = h

    def __len__(self) -> int:
        return int(lib().ae_f32_len(self._h))

    def device_ptr(self) -> int:
        p = C.c_void_p()
        call("ae_f32_device_ptr", self._h, C.byref(p))
        return p.value or 0

    def to_numpy(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.float32)
        call("ae_f32_download", self._h, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    def vec_stats(self):
        """VecStats of the f32 vector (min/max by value); see `stats.VecStats`."""
        from . import _lib
        from .stats import VecStats

        raw = _lib.VecStatsRaw()
        call("ae_f32_stats", self._h, C.byref(raw))
        return VecStats(raw, False)

    def __del__(self):
        try:
            if self._h:
                lib().ae_f32_free(self._h)
                self._h = None
        except Exception:
            pass


def spectrogram(symbols: DeviceVec, fft: Cfft, use_db: bool = True, levels: DeviceF32 | None = None) -> DeviceF32:
    """rows = ceil(len / fft.len()) chunks of fft.len() levels, fft-shifted (DC in the middle)."""
    if levels is None:
        levels = DeviceF32(0)
    call("ae_spectrogram", fft._h, symbols._h, levels._h, int(use_db))
    return levels


def correlate(inout: DeviceVec, sig: DeviceVec, fft: Cfft, scale: Scale = Scale.None_, howmany: int = 1) -> DeviceVec:
    call("ae_correlate", fft._h, inout._h, sig._h, scale.kind, C.c_float(scale.x), howmany)
    return inout
