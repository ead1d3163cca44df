#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_handles.py -m gpu -x -q > gpurun_out/spec_tests.log 2>&1; rc=$?
tail -3 gpurun_out/spec_tests.log
[ $rc -ne 0 ] && exit $rc
{ echo "x2"; timeout 200 python tools/spectral_quick.py; echo "v1"; AE_CORR_V1=1 timeout 200 python tools/spectral_quick.py; } > gpurun_out/spec_quick.log 2>&1
cat gpurun_out/spec_quick.log
