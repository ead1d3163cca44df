#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fir.py tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -3
echo "taps in the constant bank"; timeout 300 python tools/fir_quick.py 2>&1 | tail -3
echo "taps in shared memory"; AE_FIR_NO_CTAPS=1 timeout 300 python tools/fir_quick.py 2>&1 | head -1
echo "sweep"; timeout 300 python tools/fir_quick.py 67108864 sweep 2>&1 | grep direct
