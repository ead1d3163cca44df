#!/bin/bash
mkdir -p gpurun_out
{
for w in 8 11; do for sg in 0 1200 2400 3600 5000; do echo "staged warps=$w stagger=$sg"; AE_CHAIN_STAGGER=$sg AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done; done
for w in 12 16; do for sg in 0 1500 2400 4000; do echo "plain warps=$w stagger=$sg"; AE_CHAIN_NO_TMA=1 AE_CHAIN_STAGGER=$sg AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done; done
} > gpurun_out/x2_stagger.log 2>&1
cat gpurun_out/x2_stagger.log
