#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fir.py tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
if [ $v = 1 ]; then export AE_FIR_NO_CTAPS=1; echo "taps in shared memory"; else echo "taps as kernel parameters"; fi
timeout 300 python - <<'PY'
import torch, aether_primitives_b200 as ae
from aether_primitives_b200 import fir as F
from bench import make_taps
ae.init(0); ae.use_torch_stream()
for t, n in ((64, 1 << 26), (128, 1 << 26), (200, 1 << 25), (1024, 1 << 23)):
    x = torch.view_as_complex(torch.randn(n, 2, device="cuda")); y = torch.empty_like(x)
    dx, dy = ae.DeviceVec.from_torch(x), ae.DeviceVec.from_torch(y)
    f = F.Fir(make_taps(t), F.DIRECT)
    for _ in range(2): f.filter(dx, dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f.filter(dx, dy)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("direct %4d taps: %.3f ms  %.1f Gsamples/s  %.1f TFLOP/s" % (t, ms, n / ms / 1e6, 8 * t * n / ms / 1e9))
PY
done
