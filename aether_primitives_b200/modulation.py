"""modulation — device mirror of src/modulation.rs (Modulation trait, bpsk(), qpsk())."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, lib
from .vecops import DeviceBits, DeviceVec

GENERIC_BPSK_TABLE = np.array([1 + 1j, -1 - 1j], dtype=np.complex64)            # src/modulation.rs:77
GENERIC_QPSK_TABLE = np.array([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j], dtype=np.complex64)  # :87-92


class Modulation:
    """`impl Modulation for [cf32; 2]` / `[cf32; 4]` (src/modulation.rs:5-57, :94-149)."""

    def __init__(self, table):
        t = np.ascontiguousarray(table, dtype=np.complex64)
        h = C.c_void_p()
        call("ae_mod_create", t.ctypes.data_as(C.c_void_p), t.size, C.byref(h))
        self._h = h
        self.table = t.copy()

    @property
    def BITS_PER_SYMBOL(self) -> int:
        return int(lib().ae_mod_bits_per_symbol(self._h))

    def bits_per_symbol(self) -> int:  # :146-148
        return self.BITS_PER_SYMBOL

    def symbol(self, idx: int):  # :12-15 / :27-29 (slice indexing panics when out of range)
        if not 0 <= idx < self.table.size:
            raise _lib.AeError(_lib.AE_EIDX, "index out of bounds: the len is %d but the index is %d" % (self.table.size, idx))
        return self.table[idx]

    def modulate(self, input: DeviceBits) -> DeviceVec:  # :115-121 (collect into a new Vec)
        out = DeviceVec.with_capacity(max(1, len(input) // self.BITS_PER_SYMBOL))
        call("ae_mod_modulate", self._h, input._h, out._h)
        return out

    def modulate_into(self, input: DeviceBits, output: DeviceVec) -> None:  # :123-131
        call("ae_mod_modulate_into", self._h, input._h, output._h)

    def demod_naive(self, symbols: DeviceVec, output: DeviceBits, compat: int = _lib.COMPAT_REFERENCE) -> None:
        """:133-144 (BPSK) / :33-56 (QPSK override).  APPENDS to output.  compat=reference emits
        the QPSK second bit as idx & 2 in {0, 2}."""
        call("ae_mod_demod", self._h, symbols._h, output._h, compat)

    def __del__(self):
        try:
            if self._h:
                lib().ae_mod_destroy(self._h)
                self._h = None
        except Exception:
            pass


def bpsk() -> Modulation:  # :61-63
    return Modulation(GENERIC_BPSK_TABLE)


def qpsk() -> Modulation:  # :66-68
    return Modulation(GENERIC_QPSK_TABLE)
