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

    def __del__(self):
        try:
            if self._h:
                lib().ae_stats_free(self._h)
                self._h = None
        except Exception:
            pass


def evm_db(err_pow: float, ref_pow: float) -> float:
    """10*log10(P_error/P_ref) (src/lib.rs:21, util::DB src/util/mod.rs:26-34)."""
    if err_pow == 0:
        return -math.inf
    return 10.0 * math.log10(err_pow / ref_pow)


def allreduce(values: dict, group=None) -> dict:
    """Sum {bit_errors, n_bits, err_pow, ref_pow} over ranks with torch.distributed (NCCL on GPUs,
    gloo in the CPU tests).  <= 32 bytes per rank: the only collective of the whole path."""
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
