"""stats — bit-error / EVM partial sums: the only quantities reduced across GPUs."""
from __future__ import annotations

import ctypes as C
import math

from ._lib import Stats, call, lib
from .vecops import DeviceBits, DeviceVec


class DeviceStats:
    def __init__(self):
        h = C.c_void_p()
        call("ae_stats_alloc", C.byref(h))
        self._h = h

    def zero(self) -> None:
        call("ae_stats_zero", self._h)

    def read(self) -> dict:
        s = Stats()
        call("ae_stats_read", self._h, C.byref(s))
        return {"bit_errors": int(s.bit_errors), "n_bits": int(s.n_bits), "err_pow": float(s.err_pow), "ref_pow": float(s.ref_pow)}

    def count_bit_errors(self, a: DeviceBits, b: DeviceBits) -> None:
        call("ae_count_bit_errors", a._h, b._h, self._h)

    def evm_accumulate(self, actual: DeviceVec, reference: DeviceVec) -> None:
        call("ae_evm_accumulate", actual._h, reference._h, self._h)

    def allreduce(self, comm: "Comm") -> None:
        """In place on the device: this <- sum over the communicator's ranks (ncclAllReduce through the C ABI,
        asynchronous on the context stream; `read()` synchronises)."""
        call("ae_stats_allreduce", self._h, comm._h)

    def __del__(self):
        try:
            if self._h:
                lib().ae_stats_free(self._h)
                self._h = None
        except Exception:
            pass


class Comm:
    """NCCL communicator owned by the C-ABI library (`ae_comm_*`): the one collective of the path is the sum of
    `ae_stats` over the GPUs.  One process per GPU: rank 0 creates the 128-byte id, every rank joins."""

    ID_BYTES = 128

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * Comm.ID_BYTES)()
        call("ae_comm_unique_id", buf)
        return bytes(buf)

    @classmethod
    def init_rank(cls, uid: bytes, nranks: int, rank: int) -> "Comm":
        assert len(uid) == cls.ID_BYTES
        buf = (C.c_uint8 * cls.ID_BYTES).from_buffer_copy(uid)
        h = C.c_void_p()
        call("ae_comm_init_rank", buf, nranks, rank, C.byref(h))
        return cls(h)

    @classmethod
    def from_torch_distributed(cls, group=None) -> "Comm":
        """Join the ranks of an initialised torch.distributed group: torch only carries the 128-byte id from rank 0
        to the others (the out-of-band channel); the communicator and the reduction are the library's."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(cls.ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.tensor(list(cls.unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls.init_rank(bytes(t.cpu().tolist()), world, rank)

    def info(self) -> dict:
        n, r, d = C.c_int(), C.c_int(), C.c_int()
        call("ae_comm_info", self._h, C.byref(n), C.byref(r), C.byref(d))
        return {"nranks": n.value, "rank": r.value, "device": d.value}

    def close(self) -> None:
        if self._h:
            call("ae_comm_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def evm_db(err_pow: float, ref_pow: float) -> float:
    """10*log10(P_error/P_ref) (src/lib.rs:21, util::DB src/util/mod.rs:26-34)."""
    if err_pow == 0:
        return -math.inf
    return 10.0 * math.log10(err_pow / ref_pow)


def allreduce(values: dict, group=None) -> dict:
    """Host-side sum of {bit_errors, n_bits, err_pow, ref_pow} over the ranks of a torch.distributed group.  The
    product path on GPUs is `DeviceStats.allreduce(Comm)` (NCCL through the C ABI, the counters never leave the
    device); this helper reduces already-read values and is what the CPU tests run over gloo."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(values)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ints = torch.tensor([values["bit_errors"], values["n_bits"]], dtype=torch.int64, device=dev)
    flts = torch.tensor([values["err_pow"], values["ref_pow"]], dtype=torch.float64, device=dev)
    dist.all_reduce(ints, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(flts, op=dist.ReduceOp.SUM, group=group)
    return {"bit_errors": int(ints[0]), "n_bits": int(ints[1]), "err_pow": float(flts[0]), "ref_pow": float(flts[1])}


class VecStats:
    """Result of `DeviceVec.vec_stats()` / `DeviceF32.vec_stats()` — the reference's README TODO
    "VecStats (f32,cf32): Min(index), Max(index), Mean(index), Power" (README.md:90-92).  cf32 elements
    are ranked by `norm_sqr`; `min`/`max` are `(value, index)` or None when nothing is comparable."""

    def __init__(self, raw, complex_input: bool):
        self.n = int(raw.n)
        self.min = (float(raw.min_val), int(raw.min_idx)) if raw.min_idx < raw.n else None
        self.max = (float(raw.max_val), int(raw.max_idx)) if raw.max_idx < raw.n else None
        self.sum = complex(raw.sum_re, raw.sum_im) if complex_input else float(raw.sum_re)
        self.mean = self.sum / self.n
        self.power = float(raw.sum_pow) / self.n

    def __repr__(self):
        return "VecStats(n=%d, min=%r, max=%r, mean=%r, power=%r)" % (self.n, self.min, self.max, self.mean, self.power)
