"""Device/stream plumbing of the C ABI (ae_init, ae_set_stream, ae_sync)."""
from __future__ import annotations

import ctypes as C

from ._lib import call, lib


def device_count() -> int:
    n = C.c_int(0)
    try:
        call("ae_device_count", C.byref(n))
    except Exception:
        return 0
    return n.value


def init(device: int = 0) -> None:
    call("ae_init", device)


def set_stream(cuda_stream_ptr: int | None) -> None:
    call("ae_set_stream", C.c_void_p(cuda_stream_ptr or 0))


def use_torch_stream() -> None:
    """Issue all library work on torch's current CUDA stream (so torch.cuda.Event times it)."""
    import torch

    ptr = torch.cuda.current_stream().cuda_stream
    # torch's default stream is the legacy NULL stream: name it explicitly (cudaStreamLegacy = 0x1),
    # because ae_set_stream(NULL) means "the library's own stream"
    set_stream(ptr if ptr else 0x1)


def sync() -> None:
    call("ae_sync")


def launch_count() -> int:
    return int(lib().ae_launch_count())


def sm_count() -> int:
    n = C.c_int(0)
    call("ae_sm_count", C.byref(n))
    return n.value
