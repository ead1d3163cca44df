#!/bin/bash
# first GPU run of K14b: parity tests, warp-count sweep, comparison with K14, then one ncu capture
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_chain.py -x -q -k "chain" > gpurun_out/x2_tests.log 2>&1; rc=$?
tail -5 gpurun_out/x2_tests.log
[ $rc -ne 0 ] && exit $rc
{
for w in 8 10 11; do echo "staged warps=$w"; AE_CHAIN_WARPS=$w python tools/chain_quick.py; done
for w in 8 12 16; do echo "plain warps=$w"; AE_CHAIN_NO_TMA=1 AE_CHAIN_WARPS=$w python tools/chain_quick.py; done
echo "K14 (v1)"; AE_CHAIN_V1=1 python tools/chain_quick.py
for t in 1 16 32; do NTAPS=$t python tools/chain_quick.py; done
} > gpurun_out/x2_quick.log 2>&1
cat gpurun_out/x2_quick.log
python tools/chain_quick.py 262144 > gpurun_out/x2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_x2 -s 3 -c 1 -o gpurun_out/x2_prof python tools/chain_quick.py 262144 > gpurun_out/x2_ncu.log 2>&1
tail -3 gpurun_out/x2_ncu.log
