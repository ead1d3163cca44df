#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullsize.py tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/ofdm_quick.py 2>&1 | tail -6
