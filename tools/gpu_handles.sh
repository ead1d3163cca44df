#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_handles.py tests/test_cpp_host_mirror.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -4
