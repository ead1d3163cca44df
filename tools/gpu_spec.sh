#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/spec_tests.log 2>&1; rc=$?
tail -3 gpurun_out/spec_tests.log
[ $rc -ne 0 ] && exit $rc
{ for w in 12 16; do echo "x2 warps=$w"; AE_SPEC_WARPS=$w timeout 200 python tools/spectral_quick.py | grep spectrogram; done; echo "v1"; AE_SPEC_V1=1 timeout 200 python tools/spectral_quick.py | grep spectrogram; } > gpurun_out/spec_quick.log 2>&1
cat gpurun_out/spec_quick.log
