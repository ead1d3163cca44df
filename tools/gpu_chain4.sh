#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -x -q -k chain > gpurun_out/x2_tests.log 2>&1; rc=$?
tail -2 gpurun_out/x2_tests.log
[ $rc -ne 0 ] && exit $rc
AE_CHAIN_NO_TMA=1 timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -x -q -k chain > gpurun_out/x2_tests_plain.log 2>&1; rc=$?
tail -2 gpurun_out/x2_tests_plain.log
[ $rc -ne 0 ] && exit $rc
{
echo "staged warps=8"; AE_CHAIN_WARPS=8 timeout 120 python tools/chain_quick.py
for w in 8 12 16; do echo "plain shared-x warps=$w"; AE_CHAIN_NO_TMA=1 AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done
} > gpurun_out/x2_quick.log 2>&1
cat gpurun_out/x2_quick.log
