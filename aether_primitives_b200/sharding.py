"""Frame sharding across GPUs: the path is embarrassingly parallel, so every rank
owns a contiguous block of frames and no data-path collective exists; only the <= 32-byte
BER/EVM counters are all-reduced (stats.allreduce)."""
from __future__ import annotations


def frame_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the frames rank `rank` of `world` owns: frames r*F/R .. (r+1)*F/R."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (total_frames * rank) // world, (total_frames * (rank + 1)) // world


def fir_halo(ntaps: int) -> int:
    """Samples of the PREVIOUS shard's input a streaming FIR shard must read (no result exchange)."""
    return max(0, ntaps - 1)
