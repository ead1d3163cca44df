#!/bin/bash
mkdir -p gpurun_out
{
for w in 8 11; do for d in 0 1 2 3; do echo "staged warps=$w debug=$d"; AE_CHAIN_DEBUG=$d AE_CHAIN_WARPS=$w python tools/chain_quick.py; done; done
for w in 16; do for d in 0 1 2; do echo "plain warps=$w debug=$d"; AE_CHAIN_NO_TMA=1 AE_CHAIN_DEBUG=$d AE_CHAIN_WARPS=$w python tools/chain_quick.py; done; done
} > gpurun_out/x2_dbg.log 2>&1
cat gpurun_out/x2_dbg.log
