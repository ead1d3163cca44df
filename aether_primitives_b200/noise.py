"""noise — device mirror of src/noise.rs (Awgn, generator(), new()).

The RNG is Philox4x32-10 + Box-Muller keyed by (seed, stream id, global sample index) instead
of the reference's ChaCha20 + ziggurat, so the sample stream differs by construction; it is
validated statistically (tests/test_noise.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, lib
from .vecops import DeviceVec

DEFAULT_RNG_SEED = 815  # src/noise.rs:6


class Awgn:
    def __init__(self, power: float, seed: int):  # Awgn::new (src/noise.rs:29-37)
        h = C.c_void_p()
        call("ae_awgn_create", C.c_float(power), C.c_uint64(seed), C.byref(h))
        self._h = h
        self.power = float(np.float32(power))

    def set_power(self, power: float) -> None:  # :47-50
        call("ae_awgn_set_power", self._h, C.c_float(power))
        self.power = float(np.float32(power))

    def set_stream_id(self, stream_id: int) -> None:
        call("ae_awgn_set_stream_id", self._h, C.c_uint64(stream_id))

    def seek(self, sample_offset: int) -> None:
        call("ae_awgn_seek", self._h, C.c_uint64(sample_offset))

    def tell(self) -> int:
        return int(lib().ae_awgn_tell(self._h))

    def apply(self, signal: DeviceVec, compat: int = _lib.COMPAT_REFERENCE) -> None:
        """:53-59.  compat=reference scales the noise twice like the crate (sigma = power)."""
        call("ae_awgn_apply", self._h, signal._h, compat)

    def fill(self, target: DeviceVec) -> None:
        """:62-66: push noise until len == capacity."""
        call("ae_awgn_fill", self._h, target._h)

    def iter(self, block: int = 4096):
        """Awgn::iter (:68-70): infinite iterator over the stream (host side, pulled in blocks)."""
        while True:
            buf = np.empty(block, dtype=np.complex64)
            call("ae_awgn_next_host", self._h, buf.ctypes.data_as(C.c_void_p), block)
            for z in buf:
                yield z

    def next_host(self, n: int) -> np.ndarray:
        buf = np.empty(n, dtype=np.complex64)
        call("ae_awgn_next_host", self._h, buf.ctypes.data_as(C.c_void_p), n)
        return buf

    def __del__(self):
        try:
            if self._h:
                lib().ae_awgn_destroy(self._h)
                self._h = None
        except Exception:
            pass


def generator() -> Awgn:  # src/noise.rs:9-11
    return Awgn(1.0, DEFAULT_RNG_SEED)


def new(power: float, seed: int) -> Awgn:  # src/noise.rs:14-16
    return Awgn(power, seed)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    call("ae_philox4x32_10", c, k, o)
    return list(o)
