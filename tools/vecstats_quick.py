import os, sys, torch
sys.path.insert(0, os.getcwd())
import aether_primitives_b200 as ae
ae.init(0); ae.use_torch_stream()
n = 1 << 28
x = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
d = ae.DeviceVec.from_torch(x)
for _ in range(3): d.vec_stats()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.vec_stats()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("vecstats %.3f ms %.1f Gs/s %.1f%%" % (ms, n / ms / 1e6, 8 * n / ms / 1e6 / 6534.1 * 100))
