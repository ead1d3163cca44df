#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_cpp_host_mirror.py tests/test_gpu_fft.py -m gpu -x -q > gpurun_out/pipeline_tests.log 2>&1; rc=$?
tail -15 gpurun_out/pipeline_tests.log
exit $rc
