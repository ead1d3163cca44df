#!/usr/bin/env python
"""Quick device-timed run of the fused chain (for kernel iteration; bench.py is the judged run)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
from aether_primitives_b200.chain import FftFirDemod
from bench import make_taps

frames = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 29) // int(os.environ.get("NFFT", "1024"))
n = int(os.environ.get("NFFT", "1024"))
ae.init(0)
ae.use_torch_stream()
x = torch.view_as_complex(torch.randn(frames * n, 2, device="cuda"))
bits = torch.empty(2 * frames * n, dtype=torch.uint8, device="cuda")
d_in = ae.DeviceVec.from_torch(x)
d_bits = ae.DeviceBits.wrap(bits.data_ptr(), bits.numel(), owner=bits)
ntaps = int(os.environ.get('NTAPS', '64'))
ch = FftFirDemod(n, make_taps(ntaps), ae.Scale.SN)
for _ in range(3):
    ch.run(d_in, d_bits)
torch.cuda.synchronize()
# SM clock while the kernel runs (NVML, 5 ms period)
import threading
clk, stop = [], threading.Event()
def _sample():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.005)
    except Exception:
        pass
th = threading.Thread(target=_sample, daemon=True)
th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = int(os.environ.get("REPS", "10"))
e0.record()
for _ in range(K):
    ch.run(d_in, d_bits)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
stop.set()
th.join(timeout=1)
print("ntaps=%d" % ntaps, end="  ")
if clk:
    print("sm_mhz median %d min %d" % (sorted(clk)[len(clk) // 2], min(clk)), end="  ")
print("chain: %.3f ms  %.1f Gsamples/s  %.1f%% of 6534 GB/s" % (ms, frames * n / ms / 1e6, 10 * frames * n / ms / 1e6 / 6534.1 * 100))
# correctness is the job of tests/ (this helper only times the kernel)
