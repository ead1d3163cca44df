#!/bin/bash
mkdir -p gpurun_out
timeout 120 build/issue_mix > gpurun_out/issue_mix.log 2>&1
grep -v "16 LOP  \|^16 LOP" gpurun_out/issue_mix.log | grep "warps/SMSP=[14]"
{
for w in 4 8; do for d in 0 1; do echo "staged warps=$w debug=$d"; AE_CHAIN_STAGGER=0 AE_CHAIN_DEBUG=$d AE_CHAIN_WARPS=$w timeout 120 python tools/chain_quick.py; done; done
} > gpurun_out/x2_w4.log 2>&1
cat gpurun_out/x2_w4.log
