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
