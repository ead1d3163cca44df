#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -k "chain" > gpurun_out/x2_tests.log 2>&1; rc=$?
tail -3 gpurun_out/x2_tests.log
[ $rc -ne 0 ] && exit $rc
{
for w in 4 8 10; do echo "staged warps=$w"; AE_CHAIN_STAGGER=0 AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done
for w in 12 16; do echo "plain warps=$w"; AE_CHAIN_NO_TMA=1 AE_CHAIN_STAGGER=0 AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done
NTAPS=32 AE_CHAIN_STAGGER=0 AE_CHAIN_WARPS=8 timeout 120 python tools/chain_quick.py
} > gpurun_out/x2_quick.log 2>&1
cat gpurun_out/x2_quick.log
