#!/usr/bin/env python
"""Quick device-timed run of the batched FFT / spectrogram / correlator kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 29) // n
ae.init(0)
ae.use_torch_stream()
x = torch.view_as_complex(torch.randn(n * frames, 2, device="cuda"))
d = ae.DeviceVec.from_torch(x)
fft = ae.Cfft.with_len(n)
lv = ae.spectral.DeviceF32(n * frames)
sig = ae.DeviceVec.zeros(n)


def t(fn, k=10):
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


S = n * frames
for name, fn, b in (("fft fwd SN", lambda: fft.ifwd(d, ae.Scale.SN, howmany=frames), 16),
                    ("spectrogram dB", lambda: ae.spectral.spectrogram(d, fft, True, lv), 12),
                    ("correlator", lambda: ae.spectral.correlate(d, sig, fft, ae.Scale.SN, howmany=frames), 16)):
    ms = t(fn)
    print("%-16s N=%d  %.3f ms  %.1f Gsamples/s  %.1f%% of 6534 GB/s" % (name, n, ms, S / ms / 1e6, b * S / ms / 1e6 / 6534.1 * 100))
