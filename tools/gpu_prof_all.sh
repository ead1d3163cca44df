#!/bin/bash
# one ncu pass per call: --set full over every kernel of tools/profile_all.py (after the plain run exited 0); summarised on the box
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 300 python tools/profile_all.py > gpurun_out/profile_all_plain.log 2>&1 || { tail -5 gpurun_out/profile_all_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none -o gpurun_out/all_kernels python tools/profile_all.py > gpurun_out/ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/all_kernels.ncu-rep > gpurun_out/all_kernels_ncu_summary.txt 2>&1
rm -f gpurun_out/all_kernels.ncu-rep
grep -c "== kernel" gpurun_out/all_kernels_ncu_summary.txt; du -sh gpurun_out
