#!/bin/bash
# chain kernel sweep on the GPU: parity tests first, then warps x staging, tap counts, K14 for comparison
mkdir -p gpurun_out
python -m pytest tests/test_gpu_chain.py -x -q -k "chain" > gpurun_out/x2_tests.log 2>&1; rc=$?
tail -3 gpurun_out/x2_tests.log
[ $rc -ne 0 ] && exit $rc
{
for w in 8 11; do echo "staged warps=$w"; AE_CHAIN_WARPS=$w python tools/chain_quick.py; done
for w in 8 12 14 16 20; do echo "plain warps=$w"; AE_CHAIN_NO_TMA=1 AE_CHAIN_WARPS=$w python tools/chain_quick.py; done
echo "K14 (v1)"; AE_CHAIN_V1=1 python tools/chain_quick.py
for t in 1 32; do NTAPS=$t python tools/chain_quick.py; done
} > gpurun_out/x2_quick.log 2>&1
cat gpurun_out/x2_quick.log
