#!/bin/bash
mkdir -p gpurun_out
timeout 120 build/issue_mix > gpurun_out/issue_mix.log 2>&1
cat gpurun_out/issue_mix.log
{
echo "staged warps=8 reps=100"; REPS=100 AE_CHAIN_WARPS=8 timeout 120 python tools/chain_quick.py
echo "staged warps=8 debug=1 reps=100"; REPS=100 AE_CHAIN_DEBUG=1 AE_CHAIN_WARPS=8 timeout 120 python tools/chain_quick.py
echo "K14 reps=100"; REPS=100 AE_CHAIN_V1=1 timeout 120 python tools/chain_quick.py
} > gpurun_out/x2_clk.log 2>&1
cat gpurun_out/x2_clk.log
