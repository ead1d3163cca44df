"""pool::{make, Pool, Elem} (src/pool.rs) — the module name the reference uses; the implementation lives next to the
pipeline it feeds (pipeline.py)."""
from .pipeline import Elem, Pool


def make(initial_len: int, maker, resetter) -> Pool:
    return Pool.make(initial_len, maker, resetter)


__all__ = ["make", "Pool", "Elem"]
