#!/bin/bash
mkdir -p gpurun_out
{
for w in 8 10; do echo "staged warps=$w"; AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done
for w in 8 12 14 16; do echo "plain+prefetch warps=$w"; AE_CHAIN_NO_TMA=1 AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done
} > gpurun_out/x2_quick.log 2>&1
cat gpurun_out/x2_quick.log
