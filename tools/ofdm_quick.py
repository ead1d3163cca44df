#!/usr/bin/env python
"""Quick device-timed run of the OFDM-like chain (BASELINE config 5) and the fused modem loop-back."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
from aether_primitives_b200.stats import DeviceStats

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
frames = (1 << 29) // n
ae.init(0)
ae.use_torch_stream()
st = DeviceStats()


def timed(fn, k=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


ms = timed(lambda: ae.chain.ofdm_chain(n, frames, 0, 0.05, 5, st))
print("ofdm N=%d: %.3f ms  %.1f Gsymbols/s" % (n, ms, frames * n / ms / 1e6))
nsym = 1 << 28
bits = ae.DeviceBits.wrap(torch.randint(0, 2, (2 * nsym,), dtype=torch.uint8, device="cuda").data_ptr(), 2 * nsym)
out = ae.DeviceBits.zeros(2 * nsym)
m = ae.modulation.qpsk()
g = ae.noise.new(0.01, 815)
ms = timed(lambda: ae.chain.modem_fused(m, g, bits, out, st))
print("modem fused: %.3f ms  %.1f Gsymbols/s" % (ms, nsym / ms / 1e6))
