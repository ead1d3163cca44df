"""vecops — device mirror of the reference's `VecOps` trait (src/vecops.rs:39-89).

`DeviceVec` plays the role of `Vec<cf32>` / `&mut [cf32]`: the same chainable methods with
the same names, argument meaning and failure behaviour (length mismatch -> the reference's
"Vectors must have same length" panic becomes AeError(AE_ELEN)).  Calls are recorded on the
handle's op tape and executed as ONE fused CUDA kernel when the data is needed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, lib


class DeviceBits:
    """Device `Vec<u8>` holding one bit per byte (src/modulation.rs:102-103)."""

    def __init__(self, handle: int, owner=None):
        self._h = C.c_void_p(handle)
        self._owner = owner

    @classmethod
    def with_capacity(cls, capacity: int) -> "DeviceBits":
        h = C.c_void_p()
        call("ae_bits_alloc", 0, capacity, C.byref(h))
        return cls(h.value)

    @classmethod
    def zeros(cls, n: int) -> "DeviceBits":
        h = C.c_void_p()
        call("ae_bits_alloc", n, n, C.byref(h))
        return cls(h.value)

    @classmethod
    def from_numpy(cls, a) -> "DeviceBits":
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = cls.zeros(a.size)
        call("ae_bits_upload", b._h, a.ctypes.data_as(C.c_void_p), a.size)
        return b

    @classmethod
    def wrap(cls, device_ptr: int, n: int, owner=None) -> "DeviceBits":
        h = C.c_void_p()
        call("ae_bits_wrap", C.c_void_p(device_ptr), n, C.byref(h))
        return cls(h.value, owner)

    def __len__(self) -> int:
        return int(lib().ae_bits_len(self._h))

    def capacity(self) -> int:
        return int(lib().ae_bits_capacity(self._h))

    def clear(self) -> None:
        call("ae_bits_set_len", self._h, 0)

    def device_ptr(self) -> int:
        p = C.c_void_p()
        call("ae_bits_device_ptr", self._h, C.byref(p))
        return p.value or 0

    def upload_async(self, host_ptr: int, n: int) -> "DeviceBits":
        call("ae_bits_upload_async", self._h, C.c_void_p(host_ptr), n)
        return self

    def download_async(self, host_ptr: int, n: int) -> "DeviceBits":
        """stream-ordered D2H copy into (pinned) host memory; valid once the pipeline item has been received"""
        call("ae_bits_download_async", self._h, C.c_void_p(host_ptr), n)
        return self

    def to_numpy(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.uint8)
        call("ae_bits_download", self._h, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    def __del__(self):
        try:
            if self._h:
                lib().ae_bits_free(self._h)
                self._h = None
        except Exception:
            pass


_MUTATE_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float * 2), C.c_void_p)


class DeviceVec:
    """Device `Vec<cf32>` with the reference's VecOps methods (each returns self for chaining)."""

    def __init__(self, handle: int, owner=None):
        self._h = C.c_void_p(handle)
        self._owner = owner  # keeps a parent / foreign tensor alive

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def zeros(cls, n: int) -> "DeviceVec":  # vec![cf32::default(); n]
        h = C.c_void_p()
        call("ae_vec_alloc", n, n, C.byref(h))
        return cls(h.value)

    @classmethod
    def with_capacity(cls, capacity: int) -> "DeviceVec":  # Vec::with_capacity
        h = C.c_void_p()
        call("ae_vec_alloc", 0, capacity, C.byref(h))
        return cls(h.value)

    @classmethod
    def from_numpy(cls, a) -> "DeviceVec":
        a = np.ascontiguousarray(a, dtype=np.complex64)
        v = cls.zeros(a.size)
        call("ae_vec_upload", v._h, a.ctypes.data_as(C.c_void_p), a.size)
        return v

    @classmethod
    def wrap(cls, device_ptr: int, n: int, owner=None) -> "DeviceVec":
        """Borrow foreign device memory (e.g. a torch complex64 tensor's data_ptr())."""
        h = C.c_void_p()
        call("ae_vec_wrap", C.c_void_p(device_ptr), n, C.byref(h))
        return cls(h.value, owner)

    @classmethod
    def from_torch(cls, t) -> "DeviceVec":
        assert t.is_cuda and t.is_contiguous() and str(t.dtype) == "torch.complex64"
        return cls.wrap(t.data_ptr(), t.numel(), owner=t)

    @classmethod
    def from_file(cls, path: str) -> "DeviceVec":
        """util::file::binary_reader::<cf32> (src/util/file.rs:29-70): raw native-endian cf32, no header."""
        h = C.c_void_p()
        call("ae_vec_read_raw", path.encode(), C.byref(h))
        return cls(h.value)

    def to_file(self, path: str) -> None:
        """util::file::binary_writer::<cf32> (src/util/file.rs:72-107)."""
        call("ae_vec_write_raw", self._h, path.encode())

    @classmethod
    def from_csv(cls, path: str) -> "DeviceVec":
        """util::file::csv_reader for cf32 (src/util/file.rs:117-124): one `re,im` record per line, no header."""
        a = np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2)
        if a.size and a.shape[1] != 2:
            raise _lib.AeError(_lib.AE_EARG, "CSV deserialize error: expected 2 fields per record")
        return cls.from_numpy((a[:, 0] + 1j * a[:, 1]).astype(np.complex64) if a.size else np.zeros(0, np.complex64))

    def to_csv(self, path: str) -> None:
        """util::file::csv_writer for cf32 (src/util/file.rs:109-116): `re,im` per line, no header."""
        x = self.to_numpy()
        with open(path, "w") as f:
            for v in x:
                f.write("%s,%s\n" % (repr(float(np.float32(v.real))), repr(float(np.float32(v.imag)))))

    def vec_stats(self):
        """VecStats over the vector (pending VecOps are flushed first); see `stats.VecStats`."""
        from .stats import VecStats

        raw = _lib.VecStatsRaw()
        call("ae_vec_stats", self._h, C.byref(raw))
        return VecStats(raw, True)

    def view(self, start: int, stop: int) -> "DeviceVec":  # &mut v[start..stop]
        if stop < start:
            raise _lib.AeError(_lib.AE_EIDX, "slice index starts at %d but ends at %d" % (start, stop))
        h = C.c_void_p()
        call("ae_vec_view", self._h, start, stop - start, C.byref(h))
        return DeviceVec(h.value, owner=self)

    # ---- Vec-like accessors -------------------------------------------------------------------
    def __len__(self) -> int:
        return int(lib().ae_vec_len(self._h))

    def capacity(self) -> int:
        return int(lib().ae_vec_capacity(self._h))

    def clear(self) -> None:
        call("ae_vec_set_len", self._h, 0)

    def pending_ops(self) -> int:
        return int(lib().ae_vec_pending_ops(self._h))

    def flush(self) -> "DeviceVec":
        call("ae_vec_flush", self._h)
        return self

    def device_ptr(self) -> int:
        p = C.c_void_p()
        call("ae_vec_device_ptr", self._h, C.byref(p))
        return p.value or 0

    def upload(self, a) -> "DeviceVec":
        a = np.ascontiguousarray(a, dtype=np.complex64)
        call("ae_vec_upload", self._h, a.ctypes.data_as(C.c_void_p), a.size)
        return self

    def upload_async(self, host_ptr: int, n: int) -> "DeviceVec":
        """stream-ordered H2D copy of n cf32 from (pinned) host memory; no synchronisation (pipeline stages)"""
        call("ae_vec_upload_async", self._h, C.c_void_p(host_ptr), n)
        return self

    def download_async(self, host_ptr: int, n: int) -> "DeviceVec":
        call("ae_vec_download_async", self._h, C.c_void_p(host_ptr), n)
        return self

    def to_numpy(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.complex64)
        call("ae_vec_download", self._h, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    # ---- VecOps (src/vecops.rs:39-89) ----------------------------------------------------------
    def vec_scale(self, scale: float) -> "DeviceVec":
        call("ae_vec_scale", self._h, C.c_float(scale))
        return self

    def vec_mul(self, other: "DeviceVec") -> "DeviceVec":
        call("ae_vec_mul", self._h, other._h)
        return self

    def vec_div(self, other: "DeviceVec") -> "DeviceVec":
        call("ae_vec_div", self._h, other._h)
        return self

    def vec_conj(self) -> "DeviceVec":
        call("ae_vec_conj", self._h)
        return self

    def vec_mirror(self) -> "DeviceVec":
        call("ae_vec_mirror", self._h)
        return self

    def vec_clone(self, other: "DeviceVec") -> "DeviceVec":
        call("ae_vec_clone", self._h, other._h)
        return self

    def vec_zero(self) -> "DeviceVec":
        call("ae_vec_zero", self._h)
        return self

    def vec_mutate(self, f) -> "DeviceVec":
        """Arbitrary closure in element order: host round trip (documented slow path).
        `f(c: complex) -> complex` returns the new value (Python cannot mutate in place)."""

        def tramp(elem, _user):
            c = f(complex(elem.contents[0], elem.contents[1]))
            elem.contents[0] = np.float32(c.real)
            elem.contents[1] = np.float32(c.imag)

        cb = _MUTATE_CB(tramp)
        call("ae_vec_mutate", self._h, C.cast(cb, C.c_void_p), None)
        return self

    def vec_add(self, other: "DeviceVec") -> "DeviceVec":
        call("ae_vec_add", self._h, other._h)
        return self

    def vec_sub(self, other: "DeviceVec") -> "DeviceVec":
        call("ae_vec_sub", self._h, other._h)
        return self

    def vec_fft(self, scale, compat: int = _lib.COMPAT_REFERENCE) -> "DeviceVec":
        call("ae_vec_fft", self._h, scale.kind, C.c_float(scale.x), compat)
        return self

    def vec_ifft(self, scale, compat: int = _lib.COMPAT_REFERENCE) -> "DeviceVec":
        call("ae_vec_ifft", self._h, scale.kind, C.c_float(scale.x), compat)
        return self

    def vec_rfft(self, fft, scale) -> "DeviceVec":  # fft.ifwd(self, scale)  (:316-319)
        fft.ifwd(self, scale)
        return self

    def vec_rifft(self, fft, scale) -> "DeviceVec":  # fft.ibwd(self, scale)  (:321-324)
        fft.ibwd(self, scale)
        return self

    def __del__(self):
        try:
            if self._h:
                lib().ae_vec_free(self._h)
                self._h = None
        except Exception:
            pass
