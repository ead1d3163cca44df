"""fir — the filter src/fir.rs:1-22 only sketches (constructor, no filter method).

y[n] = sum_k taps[k] * x[n-k]; output length = input length.  `Fir.new(taps, input_len)` keeps
the reference constructor's signature; `filter` is the method the crate never wrote.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, lib
from .vecops import DeviceVec

AUTO, DIRECT, OVERLAP_SAVE = _lib.FIR_AUTO, _lib.FIR_DIRECT, _lib.FIR_OVERLAP_SAVE


class Fir:
    def __init__(self, taps, mode: int = AUTO):
        t = np.ascontiguousarray(taps, dtype=np.complex64)
        h = C.c_void_p()
        call("ae_fir_create", t.ctypes.data_as(C.c_void_p), t.size, mode, C.byref(h))
        self._h = h

    @classmethod
    def new(cls, taps, input_len: int = 0, mode: int = AUTO) -> "Fir":  # Fir::new (src/fir.rs:14)
        return cls(taps, mode)

    def ntaps(self) -> int:
        return int(lib().ae_fir_ntaps(self._h))

    def block_hop(self) -> int:
        """outputs per overlap-save segment (1 for the direct form): shard starts that are multiples of it
        reproduce the unsharded stream bit for bit (see sharding.ShardedFir)"""
        return int(lib().ae_fir_block_hop(self._h))

    def reset(self) -> None:
        call("ae_fir_reset", self._h)

    def filter(self, input: DeviceVec, output: DeviceVec, frame_len: int = 0) -> None:
        """frame_len = 0: streaming (history of ntaps-1 samples carried across calls);
        frame_len > 0: zero state at the start of every frame."""
        call("ae_fir_exec", self._h, input._h, output._h, frame_len)

    def __del__(self):
        try:
            if self._h:
                lib().ae_fir_destroy(self._h)
                self._h = None
        except Exception:
            pass
