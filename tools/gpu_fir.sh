#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fir.py tests/test_gpu_sharding.py tests/test_gpu_handles.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/fir_tests.log 2>&1; rc=$?
tail -4 gpurun_out/fir_tests.log
[ $rc -ne 0 ] && exit $rc
{
echo "K4b (x2)"; timeout 200 python tools/fir_quick.py
echo "K4 (v1)"; AE_FIR_OS_V1=1 timeout 200 python tools/fir_quick.py
echo "1024 taps NF=8192"; AE_FIR_NFFT=8192 timeout 200 python tools/fir_quick.py
echo "1024 taps NF=16384"; AE_FIR_NFFT=16384 timeout 200 python tools/fir_quick.py
} > gpurun_out/fir_quick.log 2>&1
cat gpurun_out/fir_quick.log
