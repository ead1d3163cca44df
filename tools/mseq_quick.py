import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
ae.init(0); ae.use_torch_stream()
init = np.array([1] + [0] * 30, dtype=np.uint8)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
for _ in range(2):
    s = ae.sequence.generate(init, [28, 31], n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    s = ae.sequence.generate(init, [28, 31], n)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("mseq %d bits:" % n + " %.3f ms  %.1f Gbit/s (1 B/bit: %.1f GB/s)" % (ms, n / ms / 1e6, n / ms / 1e6))
