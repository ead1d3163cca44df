"""sequence — device mirror of src/sequence.rs (expand, generate)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import call
from .vecops import DeviceBits


def expand(seed: int, length: int) -> DeviceBits:
    """src/sequence.rs:18-21: bit i of `seed` -> byte i."""
    out = DeviceBits.with_capacity(max(1, length))
    call("ae_mseq_expand", C.c_uint64(seed), length, out._h)
    return out


def generate(init, back_offsets, length: int) -> DeviceBits:
    """src/sequence.rs:47-53 with the generator closure given as its tap list:
    x[n] = (sum_t x[n - back_offsets[t]]) % 2, e.g. LTE x1 = [28, 31] (:42), test = [1, 2] (:62).
    An arbitrary closure cannot cross the device boundary."""
    init = np.ascontiguousarray(init, dtype=np.uint8)
    back = np.ascontiguousarray(back_offsets, dtype=np.uint32)
    out = DeviceBits.with_capacity(max(1, length, init.size))
    call("ae_mseq_generate", init.ctypes.data_as(C.c_void_p), init.size, back.ctypes.data_as(C.c_void_p), back.size, length, out._h)
    return out
