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


class Graph:
    """CUDA graph over the context stream (`ae_graph_*`): record the launches of a launch-bound step once, replay
    them with one driver call.  Outputs must be sized before recording; nothing inside the bracket may synchronise.

        with Graph() as g:          # records, does not run
            for _ in range(16):
                chain.modem_fused(...)
        g.launch()                  # replays the 16 launches
    """

    def __init__(self):
        self._h = None

    def __enter__(self) -> "Graph":
        call("ae_graph_begin")
        return self

    def __exit__(self, exc_type, exc, tb) -> bool:
        h = C.c_void_p()
        call("ae_graph_end", C.byref(h))
        self._h = h
        return False

    def launch(self) -> None:
        call("ae_graph_launch", self._h)

    def close(self) -> None:
        if self._h:
            call("ae_graph_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuf:
    """Page-locked host memory (`ae_host_alloc`) viewed as a numpy array: the staging buffers of the pipeline stages
    and of the *_host entry points (pageable memory makes every "async" copy synchronous)."""

    def __init__(self, n: int, dtype="complex64"):
        import numpy as np

        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        p = C.c_void_p()
        call("ae_host_alloc", max(self.nbytes, 1), C.byref(p))
        self._p = p
        raw = (C.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=int(n))

    @property
    def ptr(self) -> int:
        return self._p.value or 0

    def close(self) -> None:
        if self._p:
            self.array = None
            lib().ae_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
