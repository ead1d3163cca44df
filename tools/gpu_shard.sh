#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -12
