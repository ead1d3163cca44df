"""aether_primitives_b200 — B200-native (sm_100a) implementation of the cf32 hot path of
razorheadfx/aether_primitives, behind the crate's own API surface.

Module names mirror the reference crate (src/lib.rs:51-79): vecops, fft, fir, sampling,
modulation, noise, sequence.  Every data operation runs as a hand-written CUDA kernel through
the C ABI in include/aether_b200.h; there is no CPU fallback.
"""
from . import _lib
from ._lib import AeError, COMPAT_REFERENCE, COMPAT_CORRECTED
from .runtime import init, sync, set_stream, device_count, launch_count, sm_count, use_torch_stream, Graph, PinnedBuf
from .vecops import DeviceVec, DeviceBits
from .fft import Scale, Cfft
from .fir import Fir
from . import sampling, modulation, noise, sequence, chain, stats, spectral, util, pipeline, pool

cf32 = "complex64"  # numpy dtype of the reference's cf32 (src/lib.rs:12)

__all__ = [
    "AeError", "COMPAT_REFERENCE", "COMPAT_CORRECTED", "init", "sync", "set_stream", "device_count",
    "launch_count", "sm_count", "use_torch_stream", "Graph", "PinnedBuf", "DeviceVec", "DeviceBits", "Scale", "Cfft", "Fir",
    "sampling", "modulation", "noise", "sequence", "chain", "stats", "spectral", "util", "pipeline", "pool", "cf32",
]
