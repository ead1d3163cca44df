#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_host_mirror.py tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/mirror_tests.log 2>&1; rc=$?
tail -5 gpurun_out/mirror_tests.log
timeout 200 build/tc_operand_prep > gpurun_out/tc_operand_prep.txt 2>&1; cat gpurun_out/tc_operand_prep.txt
exit $rc
