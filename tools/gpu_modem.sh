#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_noise.py tests/test_gpu_chain.py tests/test_gpu_fullsize.py tests/test_gpu_elementwise.py tests/test_cpp_host_mirror.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, aether_primitives_b200 as ae
from aether_primitives_b200.stats import DeviceStats
ae.init(0); ae.use_torch_stream()
n = 1 << 28
bits = ae.DeviceBits.wrap((t := torch.randint(0, 2, (2 * n,), dtype=torch.uint8, device="cuda")).data_ptr(), 2 * n, owner=t)
out = ae.DeviceBits.zeros(2 * n)
m = ae.modulation.qpsk(); g = ae.noise.new(0.5, 7); st = DeviceStats()
for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
    for _ in range(2): ae.chain.modem_fused(m, g, bits, out, st, compat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ae.chain.modem_fused(m, g, bits, out, st, compat)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("modem_fused compat=%d: %.3f ms  %.1f Gsymbols/s" % (compat, ms, n / ms / 1e6))
PY
