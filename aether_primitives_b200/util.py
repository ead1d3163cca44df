"""util — the host-side helper of the reference's `util` module that the signal path touches:
`DB` (src/util/mod.rs:11-46), decibel <-> ratio in f64.  (`util::file` raw sample files are
`DeviceVec.from_file` / `to_file`; `util::plot` is out of scope.)"""
from __future__ import annotations

import math


class DB:
    """`DB(x)` stores a value in dB; `DB.from_ratio(r)` is `DB::from(r)` = 10·log10(r) (src/util/mod.rs:26-34)."""

    __slots__ = ("value",)

    def __init__(self, db: float):
        self.value = float(db)

    @classmethod
    def from_ratio(cls, ratio: float) -> "DB":
        r = float(ratio)
        if r > 0.0:
            return cls(10.0 * math.log10(r))
        return cls(-math.inf if r == 0.0 else math.nan)      # f64::log semantics: log(0) = -inf, log(<0) = NaN

    def db(self) -> float:                                   # :39-41
        return self.value

    def ratio(self) -> float:                                # :44-46
        return math.pow(10.0, self.value / 10.0)

    def __eq__(self, other):
        return isinstance(other, DB) and self.value == other.value

    def __repr__(self):
        return "DB(%r)" % self.value


def assert_evm(actual, ref, evm_limit_db: float = -80.0) -> None:
    """`assert_evm!(actual, ref[, evm_limit_db])` (src/lib.rs:26-49) on host data (numpy complex64 or anything
    with `to_numpy()`): per element, `(act - re).norm() > re.norm() * (10^(dB/10) as f32)` fails.  `norm` is
    `hypot` in f32; the limit factor is computed in f64 and cast to f32, exactly as the macro does.
    Raises AssertionError with the macro's messages."""
    import numpy as np

    a = np.asarray(actual.to_numpy() if hasattr(actual, "to_numpy") else actual, dtype=np.complex64).ravel()
    r = np.asarray(ref.to_numpy() if hasattr(ref, "to_numpy") else ref, dtype=np.complex64).ravel()
    if a.size != r.size:
        raise AssertionError("Input slices/vectors must be same length")
    if not float(evm_limit_db) < 0.0:
        raise AssertionError("The EVM threshold must be negative")
    d = (a - r).astype(np.complex64)
    evm = np.hypot(d.real.astype(np.float32), d.imag.astype(np.float32)).astype(np.float32)
    factor = np.float32(math.pow(10.0, float(evm_limit_db) / 10.0))
    limit = (np.hypot(r.real.astype(np.float32), r.imag.astype(np.float32)).astype(np.float32) * factor).astype(np.float32)
    bad = np.nonzero(evm > limit)[0]
    if bad.size:
        i = int(bad[0])
        with np.errstate(divide="ignore"):
            evm_db = float(np.log10(evm[i]) * np.float32(10.0))
        raise AssertionError("EVM limit exceeded:  %s(%sdB) > %s(%sdB) for element %d. Actual %s, Expected %s"
                             % (evm[i], evm_db, limit[i], evm_limit_db, i, a[i], r[i]))
