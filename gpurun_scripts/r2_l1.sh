#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
timeout 300 python gpurun_scripts/gpu_compress.py 1 16384 2>&1 | grep "L1:"
