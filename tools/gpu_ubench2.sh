#!/bin/bash
mkdir -p gpurun_out
timeout 200 build/issue_cost > gpurun_out/issue_cost.txt 2>&1; cat gpurun_out/issue_cost.txt
