#!/bin/bash
mkdir -p gpurun_out
{ echo "CT=8 (default)"; timeout 300 python tools/fft_big_quick.py; echo "CT=16 for 2^20"; AE_COL_CT16=1 timeout 300 python tools/fft_big_quick.py; } > gpurun_out/fft_big.log 2>&1
cat gpurun_out/fft_big.log
AE_COL_CT16=1 timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -x -q -k "four_step" 2>&1 | tail -2
