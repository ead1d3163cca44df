#!/usr/bin/env python
"""Quick device-timed run of the FIR kernels (kernel iteration helper)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
from aether_primitives_b200 import fir as F
from bench import make_taps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
ae.init(0)
ae.use_torch_stream()
x = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
y = torch.empty_like(x)
dx, dy = ae.DeviceVec.from_torch(x), ae.DeviceVec.from_torch(y)
cases = [("direct64", 64, F.DIRECT), ("os64", 64, F.OVERLAP_SAVE), ("os1024", 1024, F.OVERLAP_SAVE)]
if len(sys.argv) > 2:   # threshold sweep: direct vs overlap-save for short filters
    cases = [(("direct%d" if m == F.DIRECT else "os%d") % t, t, m) for t in (2, 4, 6, 8, 12, 16, 24, 32) for m in (F.DIRECT, F.OVERLAP_SAVE)]
for name, t, mode in cases:
    filt = F.Fir(make_taps(t), mode)
    for _ in range(3):
        filt.filter(dx, dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        filt.filter(dx, dy)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-9s %.3f ms  %.1f Gsamples/s  %.1f%% of HBM  %.1f TFLOP/s(8T/sample)" % (name, ms, n / ms / 1e6, 16 * n / ms / 1e6 / 6534.1 * 100, 8 * t * n / ms / 1e9))
