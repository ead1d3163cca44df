"""chain — fused kernels for the chains the reference composes call by call:
examples/modem.rs:15-32 (modem loop-back), the headline FFT -> FIR -> QPSK demod chain and the
OFDM-like chain of BASELINE config 5."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call
from .fft import Scale
from .modulation import Modulation
from .noise import Awgn
from .stats import DeviceStats
from .vecops import DeviceBits, DeviceVec


def modem_fused(m: Modulation, awgn: Awgn, bits_in: DeviceBits, bits_out: DeviceBits, stats: DeviceStats | None = None,
                compat: int = _lib.COMPAT_REFERENCE) -> None:
    call("ae_modem_fused", m._h, awgn._h, bits_in._h, bits_out._h, stats._h if stats is not None else None, compat)


class FftFirDemod:
    """Per frame: Cfft::fwd(scale) -> FIR (zero state at each frame start) -> QPSK demod_naive."""

    def __init__(self, fft_len: int, taps, scale: Scale = Scale.SN, compat: int = _lib.COMPAT_REFERENCE):
        t = np.ascontiguousarray(taps, dtype=np.complex64)
        h = C.c_void_p()
        call("ae_chain_create", fft_len, t.ctypes.data_as(C.c_void_p), t.size, scale.kind, C.c_float(scale.x), compat, C.byref(h))
        self._h = h
        self.fft_len = fft_len

    def run(self, input: DeviceVec, bits_out: DeviceBits) -> None:
        call("ae_chain_exec", self._h, input._h, bits_out._h)

    def run_unfused(self, input: DeviceVec, bits_out: DeviceBits, symbols_out: DeviceVec | None = None) -> None:
        call("ae_chain_exec_unfused", self._h, input._h, bits_out._h, symbols_out._h if symbols_out is not None else None)

    def run_host(self, host_in_ptr: int, n_samples: int, host_bits_ptr: int) -> None:
        """HOST buffers (ideally pinned): chunked H2D -> kernel -> D2H pipeline inside the call."""
        call("ae_chain_exec_host", self._h, C.c_void_p(host_in_ptr), n_samples, C.c_void_p(host_bits_ptr))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ae_chain_destroy(self._h)
                self._h = None
        except Exception:
            pass


def ofdm_chain(fft_len: int, frames: int, first_frame_id: int, noise_power: float, noise_seed: int,
               stats: DeviceStats | None, tx_bits: DeviceBits | None = None, rx_bits: DeviceBits | None = None,
               compat: int = _lib.COMPAT_REFERENCE) -> None:
    call("ae_ofdm_chain", fft_len, frames, C.c_uint64(first_frame_id), C.c_float(noise_power), C.c_uint64(noise_seed), compat,
         tx_bits._h if tx_bits is not None else None, rx_bits._h if rx_bits is not None else None, stats._h if stats is not None else None)


class ChainPipeline:
    """Streaming form of `FftFirDemod.run_host` — the analogue of the reference's thread pipeline and
    buffer pool (src/pipeline.rs:26-137, src/pool.rs:43-130) for this path: `send` queues one block of
    `block_frames` frames (H2D copy -> fused kernel -> D2H copy on one of `depth` buffer slots) and returns
    at once; `recv` returns the address of the next finished bit buffer, in order; `report` gives what
    each reference stage prints once a second (processed, active time, rate, utilisation)."""

    def __init__(self, chain: FftFirDemod, block_frames: int, depth: int = 3):
        h = C.c_void_p()
        call("ae_pipe_create", chain._h, block_frames, depth, C.byref(h))
        self._h = h
        self._chain = chain              # the pipe borrows the chain's window / taps
        self.block_frames = block_frames

    def send(self, host_in_ptr: int, host_bits_ptr: int) -> None:
        call("ae_pipe_send", self._h, C.c_void_p(host_in_ptr), C.c_void_p(host_bits_ptr))

    def recv(self) -> int:
        p = C.c_void_p()
        call("ae_pipe_recv", self._h, C.byref(p))
        return p.value or 0

    def in_flight(self) -> int:
        return int(_lib.lib().ae_pipe_in_flight(self._h))

    def report(self, reset: bool = False) -> list[dict]:
        st = (_lib.PipeStage * 3)()
        call("ae_pipe_report", self._h, st, int(reset))
        return [dict(name=s.name.decode(), processed=int(s.processed), active_ms=s.active_ms, elapsed_ms=s.elapsed_ms,
                     per_second=s.per_second, utilisation_pct=s.utilisation_pct) for s in st]

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ae_pipe_destroy(self._h)
                self._h = None
        except Exception:
            pass
