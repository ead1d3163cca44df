#!/bin/bash
mkdir -p gpurun_out
timeout 200 build/regbw > gpurun_out/regbw.txt 2>&1; cat gpurun_out/regbw.txt
