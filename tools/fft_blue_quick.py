import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
ae.init(0); ae.use_torch_stream()
for n in (1009, 4099, 10000, 12288, 12289, 100000):
    frames = max(1, (1 << 25) // n)
    x = torch.view_as_complex(torch.randn(n * frames, 2, device="cuda"))
    d = ae.DeviceVec.from_torch(x)
    f = ae.Cfft.with_len(n)
    for _ in range(2):
        f.ifwd(d, ae.Scale.SN, howmany=frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f.ifwd(d, ae.Scale.SN, howmany=frames)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("fft N=%d x %d frames: %.3f ms  %.1f Gsamples/s" % (n, frames, ms, n * frames / ms / 1e6))
