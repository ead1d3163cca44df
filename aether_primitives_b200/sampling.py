"""sampling — device mirror of src/sampling.rs (interpolate, downsample, downsample_sb)."""
from __future__ import annotations

from . import _lib
from ._lib import call
from .vecops import DeviceBits, DeviceVec


def interpolate(src: DeviceVec, dst: DeviceVec, n_between: int, compat: int = _lib.COMPAT_REFERENCE) -> None:
    """src/sampling.rs:7-24.  APPENDS (len-1)*(n_between+1)+1 samples to dst.
    compat=reference reproduces `im: x1.re + i*rate.1` (:19)."""
    call("ae_interpolate", src._h, dst._h, n_between, compat)


def downsample(src, dst, strict: bool = True) -> None:
    """src/sampling.rs:28-42: dst[i] = src[i * (len(src)/len(dst))]; generic over T: Copy.
    strict=True enforces the debug_assert on divisibility (what `cargo test` runs)."""
    if isinstance(src, DeviceBits):
        call("ae_downsample_bits", src._h, dst._h, int(strict))
    else:
        call("ae_downsample", src._h, dst._h, int(strict))


def downsample_sb(src, dst, strict: bool = True) -> None:
    """src/sampling.rs:49-62 (step_by variant; identical result)."""
    if isinstance(src, DeviceBits):
        call("ae_downsample_bits", src._h, dst._h, int(strict))
    else:
        call("ae_downsample_sb", src._h, dst._h, int(strict))
