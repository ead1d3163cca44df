import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
ae.init(0); ae.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
frames = (1 << 26) // n
x = torch.view_as_complex(torch.randn(n * frames, 2, device="cuda"))
d = ae.DeviceVec.from_torch(x)
f = ae.Cfft.with_len(n)
for _ in range(4):
    f.ifwd(d, ae.Scale.SN, howmany=frames)
torch.cuda.synchronize()
print("ok")
