"""fft — device mirror of `fft::{Scale, Fft, Cfft}` (src/fft.rs:6-235)."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import call, lib
from .vecops import DeviceVec


class Scale:
    """fft::Scale (src/fft.rs:6-18): None_, SN, N, X(f32)."""

    def __init__(self, kind: int, x: float = 1.0):
        self.kind = kind
        self.x = float(x)

    @staticmethod
    def X(x: float) -> "Scale":
        return Scale(_lib.SCALE_X, x)

    def factor(self, n: int) -> float:
        s = C.c_float(0)
        call("ae_scale_factor", self.kind, n, C.c_float(self.x), C.byref(s))
        return s.value

    def scale(self, data: DeviceVec) -> None:  # Scale::scale (src/fft.rs:22-37)
        call("ae_vec_scale_kind", data._h, self.kind, C.c_float(self.x))

    def __repr__(self):
        return "Scale(%s)" % {0: "None", 1: "SN", 2: "N"}.get(self.kind, "X(%r)" % self.x)


Scale.None_ = Scale(_lib.SCALE_NONE)
Scale.SN = Scale(_lib.SCALE_SN)
Scale.N = Scale(_lib.SCALE_N)


class Cfft:
    """fft::Cfft (src/fft.rs:134-235) implementing the `Fft` trait (src/fft.rs:48-77).

    compat=COMPAT_REFERENCE keeps the crate's behaviour that `fwd` runs rustfft's *inverse*
    kernel, i.e. exp(+2 pi i nk/N) (src/fft.rs:148); COMPAT_CORRECTED swaps it.
    `howmany` > 1 transforms consecutive frames of len() samples in one launch (new surface).
    """

    def __init__(self, length: int, compat: int = _lib.COMPAT_REFERENCE):
        h = C.c_void_p()
        call("ae_fft_create", length, C.byref(h))
        self._h = h
        call("ae_fft_set_compat", self._h, compat)

    @classmethod
    def with_len(cls, length: int, compat: int = _lib.COMPAT_REFERENCE) -> "Cfft":
        return cls(length, compat)

    def len(self) -> int:
        return int(lib().ae_fft_len(self._h))

    __len__ = len

    def _howmany(self, v: DeviceVec, howmany):
        return 1 if howmany is None else howmany

    def fwd(self, input: DeviceVec, output: DeviceVec, s: Scale, howmany=None) -> None:
        call("ae_fft_exec", self._h, _lib.FFT_FWD, input._h, output._h, s.kind, C.c_float(s.x), self._howmany(input, howmany))

    def bwd(self, input: DeviceVec, output: DeviceVec, s: Scale, howmany=None) -> None:
        call("ae_fft_exec", self._h, _lib.FFT_BWD, input._h, output._h, s.kind, C.c_float(s.x), self._howmany(input, howmany))

    def ifwd(self, input: DeviceVec, s: Scale, howmany=None) -> None:
        call("ae_fft_exec", self._h, _lib.FFT_FWD, input._h, None, s.kind, C.c_float(s.x), self._howmany(input, howmany))

    def ibwd(self, input: DeviceVec, s: Scale, howmany=None) -> None:
        call("ae_fft_exec", self._h, _lib.FFT_BWD, input._h, None, s.kind, C.c_float(s.x), self._howmany(input, howmany))

    def _tmp(self, direction, input, s, howmany):
        view = C.c_void_p()
        call("ae_fft_exec_tmp", self._h, direction, input._h, s.kind, C.c_float(s.x), self._howmany(input, howmany), C.byref(view))
        return DeviceVec(view.value, owner=self)  # plan-owned: ae_vec_free on it is a no-op

    def tfwd(self, input: DeviceVec, s: Scale, howmany=None) -> DeviceVec:
        return self._tmp(_lib.FFT_FWD, input, s, howmany)

    def tbwd(self, input: DeviceVec, s: Scale, howmany=None) -> DeviceVec:
        return self._tmp(_lib.FFT_BWD, input, s, howmany)

    def __del__(self):
        try:
            if self._h:
                lib().ae_fft_destroy(self._h)
                self._h = None
        except Exception:
            pass
