#!/usr/bin/env python
"""Quick device-timed run of interpolate x4 / downsample by 4 on random data."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
ae.init(0); ae.use_torch_stream()
m = 1 << 28
x = torch.view_as_complex(torch.randn(m, 2, device="cuda"))
a = ae.DeviceVec.from_torch(x)
def timed(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
src = a.view(0, m // 4)
dst = ae.DeviceVec.with_capacity(m)
def interp():
    dst.clear(); ae.sampling.interpolate(src, dst, 3)
ms = timed(interp); print("interpolate x4: %.3f ms %.1f Ginputs/s %.1f%% of HBM" % (ms, (m // 4) / ms / 1e6, 40 * (m // 4) / ms / 1e6 / 6534.1 * 100))
ds = ae.DeviceVec.zeros(m // 4)
ms = timed(lambda: ae.sampling.downsample(a, ds)); print("downsample /4: %.3f ms %.1f Goutputs/s %.1f%% of HBM" % (ms, (m // 4) / ms / 1e6, 16 * (m // 4) / ms / 1e6 / 6534.1 * 100))
