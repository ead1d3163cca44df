#!/usr/bin/env python
"""Quick device-timed run of the spectrogram core and the correlator (kernel iteration helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = (1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 29)) // n
ae.init(0)
ae.use_torch_stream()
x = torch.view_as_complex(torch.randn(frames * n, 2, device="cuda"))
dx = ae.DeviceVec.from_torch(x)
fft = ae.Cfft.with_len(n)
sig = ae.DeviceVec.from_torch(torch.view_as_complex(torch.randn(n, 2, device="cuda")))


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


lv = ae.spectral.spectrogram(dx, fft, True)
for name, fn, bps in (("spectrogram dB", lambda: ae.spectral.spectrogram(dx, fft, True, lv), 12.0),
                      ("spectrogram lin", lambda: ae.spectral.spectrogram(dx, fft, False, lv), 12.0),
                      ("fft fwd", lambda: fft.ifwd(dx, ae.Scale.SN, howmany=frames), 16.0),
                      ("correlator", lambda: ae.spectral.correlate(dx, sig, fft, ae.Scale.SN, howmany=frames), 16.0)):
    ms = timed(fn)
    print("N=%d %-16s %.3f ms  %.1f Gsamples/s  %.1f%% of 6534 GB/s" % (n, name, ms, frames * n / ms / 1e6, bps * frames * n / ms / 1e6 / 6534.1 * 100))
