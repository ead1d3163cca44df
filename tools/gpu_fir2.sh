#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fir.py tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/fir_tests.log 2>&1; rc=$?
tail -3 gpurun_out/fir_tests.log
[ $rc -ne 0 ] && exit $rc
{
for w in 8 10 12 16; do echo "K4b warps=$w"; AE_FIR_WARPS=$w timeout 200 python tools/fir_quick.py 268435456 | grep os64; done
} > gpurun_out/fir_quick.log 2>&1
cat gpurun_out/fir_quick.log
