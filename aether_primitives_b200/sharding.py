"""Sharding across GPUs.  The path is embarrassingly parallel: every rank owns a contiguous
block of frames (or of stream samples) and no data-path collective exists; only the <= 32-byte BER/EVM
counters are all-reduced (stats.DeviceStats.allreduce over a stats.Comm).

A streaming FIR is the one operator whose shards overlap: rank r needs the ntaps-1 input samples before its
first output (the halo) and discards the outputs computed there; nothing is exchanged between ranks
(BASELINE config 3, "frame-sharded across 8 B200")."""
from __future__ import annotations

from . import fir as _fir
from .vecops import DeviceVec


def frame_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the frames rank `rank` of `world` owns: frames r*F/R .. (r+1)*F/R."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (total_frames * rank) // world, (total_frames * (rank + 1)) // world


def fir_halo(ntaps: int) -> int:
    """Samples of the PREVIOUS shard's input a streaming FIR shard must read (no result exchange)."""
    return max(0, ntaps - 1)


class ShardedFir:
    """One rank's part of a streaming FIR over a global stream of `total` samples with zero initial state.

    The rank owns outputs [lo, hi) = frame_range(total, rank, world) and reads inputs [lo_in, hi_in), where
    lo_in = lo - halo and halo >= ntaps - 1 (0 on the rank that starts the stream).  Both ends are rounded outwards
    to multiples of the filter's block hop (hi_in is clipped to the stream's end) and the halo holds at least one
    whole hop: the overlap-save segments that produce kept outputs then ARE segments of the unsharded run, fed with
    the same samples, so the kept outputs are bit-identical to it (a segment that started from zero state or ended
    at hi would see zeros where the whole stream has samples, and an FFT-based convolution rounds differently then).
    For the direct form the hop is 1: the halo is exactly ntaps - 1 and nothing is read past hi.
    """

    def __init__(self, taps, total: int, rank: int, world: int, mode: int = _fir.AUTO):
        self.fir = _fir.Fir(taps, mode)
        self.total, self.rank, self.world = total, rank, world
        self.lo, self.hi = frame_range(total, rank, world)
        t1 = fir_halo(self.fir.ntaps())
        hop = max(1, self.fir.block_hop())
        # overlap-save: the shard's first segment starts from zero state, so its whole FFT block differs from the
        # unsharded run's; discard it entirely (one full hop of halo), the next segment then sees true samples only
        need = max(t1, hop if hop > 1 else 0)
        lo_in = max(0, self.lo - need)
        self.lo_in = (lo_in // hop) * hop
        self.halo = self.lo - self.lo_in
        self.hi_in = min(total, -(-self.hi // hop) * hop)
        self._work = None

    def input_range(self) -> tuple[int, int]:
        """[begin, end) of the global stream this rank must hold on its device (halo + shard + read-ahead)"""
        return self.lo_in, self.hi_in

    def filter(self, x_with_halo: DeviceVec) -> DeviceVec:
        """x_with_halo = global samples [lo_in, hi_in).  Returns a view of the rank's outputs [lo, hi) (no copy: the
        halo and read-ahead outputs are simply not part of the view)."""
        n = self.hi_in - self.lo_in
        if len(x_with_halo) != n:
            raise ValueError("shard input must hold %d samples (halo %d + shard %d + read-ahead %d)"
                             % (n, self.halo, self.hi - self.lo, self.hi_in - self.hi))
        if self._work is None or len(self._work) != n:
            self._work = DeviceVec.zeros(n)
        self.fir.reset()                       # zero state at lo_in: exact for rank 0, discarded with the halo elsewhere
        self.fir.filter(x_with_halo, self._work)
        return self._work.view(self.halo, self.halo + self.hi - self.lo)
